#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-600
echo "== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_ours.json | cut -c1-3000
