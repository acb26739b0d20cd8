#!/bin/bash
echo "== tests warp"; timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "warp or deform" 2>&1 | tail -5
echo "== tests net"; timeout 600 python -m pytest tests/test_network_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== profile"; timeout 300 python tools/conv_profile.py 2>&1 | head -14
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample-steps 0 2>&1 | tail -1 | cut -c1-220
