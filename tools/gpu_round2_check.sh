#!/bin/bash
# round-2 last GPU batch: launch list of one configs[2] step (who owns the 52 ms), configs[4] step on one GPU
mkdir -p gpurun_out
echo "== launch list fwdbwd"; timeout 100 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fwdbwd.csv python bench.py --config fwdbwd --steps 1 --warmup 3 --cpu-sample-steps 0 --sustain-seconds 0 > gpurun_out/launches_fwdbwd.log 2>&1; echo "rc=$?"; wc -l gpurun_out/launches_fwdbwd.csv
echo "== train step, one rank"; timeout 80 python bench.py --config train8 --steps 5 --warmup 3 --cpu-sample-steps 0 --sustain-seconds 0 2>&1 | tail -1 > gpurun_out/bench_train_1rank.json; cut -c1-250 gpurun_out/bench_train_1rank.json
