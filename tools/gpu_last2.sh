#!/bin/bash
mkdir -p gpurun_out
timeout 105 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --cpu-sample-steps 0 > gpurun_out/launches_bench.log 2>&1; echo "rc=$?"; wc -l gpurun_out/launches.csv
