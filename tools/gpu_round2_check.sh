#!/bin/bash
# round-2 final GPU batch: full GPU suite, configs[2] and configs[1] bench lines, smoke
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 250 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest_full.log 2>&1; echo "rc=$?"; tail -14 gpurun_out/pytest_full.log | cut -c1-300
echo "== bench fwdbwd"; timeout 120 python bench.py --config fwdbwd --steps 10 --warmup 3 --cpu-sample-steps 0 2>&1 | tail -1 > gpurun_out/bench_fwdbwd_final.json; cut -c1-330 gpurun_out/bench_fwdbwd_final.json
echo "== bench ours"; timeout 200 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_ours.json; cut -c1-300 gpurun_out/bench_ours.json
echo "== smoke"; timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
