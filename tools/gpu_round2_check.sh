#!/bin/bash
# one GPU call: new-kernel parity (TMA correlation, exact-linearity warp, cascade), then timings
mkdir -p gpurun_out
echo "== dev_tma"; timeout 300 python tools/dev_tma.py 2>&1 | tail -45
echo "== pytest new"; timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_network_gpu.py -x -q -m gpu -k "linearity or band or cascade or correlation_parity or fused_leaky or product_graph" 2>&1 | tail -15
echo "== kbench warp"; timeout 300 python tools/kbench.py --what warp --iters 10 2>&1 | grep -v "^$" | tail -8
