#!/bin/bash
mkdir -p gpurun_out
timeout 150 python bench.py --steps 10 --warmup 3 --cpu-sample-steps 0 2>&1 | tail -1 > gpurun_out/bench_last.json; echo "rc=$?"; cut -c1-2500 gpurun_out/bench_last.json
