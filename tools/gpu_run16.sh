#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== bench umma all"; timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample-steps 0 2>&1 | tail -1 > gpurun_out/bench_umma.json; cut -c1-250 gpurun_out/bench_umma.json; python -c "
import json; d=json.load(open('gpurun_out/bench_umma.json')); print(d['roofline']['per_kernel_ms']); print(d['e2e'])"
echo "== bench umma min_w 64"; MFN_TUNING="conv_umma_min_w=64" timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample-steps 0 2>&1 | tail -1 | cut -c1-200
echo "== bench sync"; MFN_TUNING="conv_umma=0" timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample-steps 0 2>&1 | tail -1 | cut -c1-200
echo "== launch list (timed region only)"; timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --cpu-sample-steps 0 > gpurun_out/launches_bench.log 2>&1; echo "rc=$?"; wc -l gpurun_out/launches.csv
