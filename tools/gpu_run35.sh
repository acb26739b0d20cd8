#!/bin/bash
echo "== tests conv"; timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "conv" 2>&1 | tail -3
echo "== tests net"; timeout 600 python -m pytest tests/test_network_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== profile"; timeout 300 python tools/conv_profile.py 2>&1 | head -24
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample-steps 0 2>&1 | tail -1 | cut -c1-220
echo "== bench PREV"; MFN_LIB_PATH=$PWD/tools/lib_prev.so timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample-steps 0 2>&1 | tail -1 | cut -c1-220
