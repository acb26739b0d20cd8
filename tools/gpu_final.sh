#!/bin/bash
# round-end artefacts: full GPU suite, smoke, bench (both arms), launch list, per-op profile, kernel micro-benchmarks
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench ours"; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_ours.json; cut -c1-300 gpurun_out/bench_ours.json
echo "== bench ref"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref.json; cut -c1-300 gpurun_out/bench_ref.json
echo "== per-op profile"; timeout 300 python tools/conv_profile.py > gpurun_out/conv_profile.txt 2>&1; head -3 gpurun_out/conv_profile.txt
echo "== kbench"; rm -f gpurun_out/kbench.jsonl; timeout 600 python tools/kbench.py --what corr,warp,bwd --iters 20 > gpurun_out/kbench_cold.log 2>&1; echo "rc=$?"
echo "== launch list"; timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --cpu-sample-steps 0 > gpurun_out/launches_bench.log 2>&1; echo "rc=$?"; wc -l gpurun_out/launches.csv
echo "== ncu umma 579->128"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_umma -s 2 -c 1 -f -o gpurun_out/prof_umma_n128 python tools/prof_umma.py 579 128 > gpurun_out/ncu_umma128.log 2>&1; echo "rc=$?"
echo "== ncu umma 547->32"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_umma -s 2 -c 1 -f -o gpurun_out/prof_umma_n32 python tools/prof_umma.py 547 32 > gpurun_out/ncu_umma32.log 2>&1; echo "rc=$?"
