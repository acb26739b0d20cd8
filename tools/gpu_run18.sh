#!/bin/bash
mkdir -p gpurun_out
echo "== tests conv"; timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "conv" 2>&1 | tail -8
echo "== tests net"; timeout 600 python -m pytest tests/test_network_gpu.py -x -q -m gpu 2>&1 | tail -5
echo "== try"; timeout 300 python tools/try_umma.py 2>&1 | tail -24
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample-steps 0 2>&1 | tail -1 | cut -c1-220
