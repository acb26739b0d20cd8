#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== bench ours"; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_ours.json; cut -c1-1500 gpurun_out/bench_ours.json
echo "== launch list"; timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --cpu-sample-steps 0 > gpurun_out/launches_bench.log 2>&1; echo "rc=$?"; wc -l gpurun_out/launches.csv; tail -3 gpurun_out/launches_bench.log
