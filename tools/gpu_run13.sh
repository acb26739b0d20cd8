#!/bin/bash
mkdir -p gpurun_out
echo "== tests corr"; timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "corr" 2>&1 | tail -5
echo "== kbench th4"; timeout 300 python tools/kbench.py --what corr --iters 20 --levels 2 --algos mma_bf16x3 --ring-th 4 2>&1 | tail -2
echo "== kbench th8"; timeout 300 python tools/kbench.py --what corr --iters 20 --levels 2 --algos mma_bf16x3 --ring-th 8 2>&1 | tail -2
echo "== dbg th4"; timeout 300 python tools/dbg_ring.py 4 2>&1 | tail -22
