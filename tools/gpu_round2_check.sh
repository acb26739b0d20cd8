#!/bin/bash
mkdir -p gpurun_out
echo "== pytest warp"; timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_network_gpu.py -x -q -m gpu -k "linearity or band or warp or product or cascade_matches" 2>&1 | tail -4
echo "== kbench warp"; timeout 300 python tools/kbench.py --what warp --levels 2,3,4,5 --iters 10 2>&1 | grep -v "^$" | cut -c1-230
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --cpu-sample-steps 0 2>&1 | tail -1 | cut -c1-800
