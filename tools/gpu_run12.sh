#!/bin/bash
mkdir -p gpurun_out
echo "== kbench cold"; rm -f gpurun_out/kbench.jsonl; timeout 600 python tools/kbench.py --what corr,warp,bwd --iters 20 > gpurun_out/kbench_cold.log 2>&1; echo "rc=$?"
echo "== launch list (ncu, 1 step)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --cpu-sample-steps 0 > gpurun_out/launches_bench.log 2>&1; echo "rc=$?"; wc -l gpurun_out/launches.csv
echo "== ncu full: corr ring kernel"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_mma_ring -s 3 -c 1 -f -o gpurun_out/prof_ring python tools/kbench.py --what corr --iters 1 --levels 2 --algos mma_bf16x3 > gpurun_out/ncu_ring.log 2>&1; echo "rc=$?"
echo "== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_ours.json; cut -c1-400 gpurun_out/bench_ours.json
echo "== bench ref"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref.json; cut -c1-300 gpurun_out/bench_ref.json
