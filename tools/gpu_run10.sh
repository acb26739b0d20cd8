#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['roofline']['hot_path_ms_per_step'])
"
echo "== profile"; timeout 600 python tools/exp_convs.py 2>&1 | cut -c1-62,118-200 | tail -16
