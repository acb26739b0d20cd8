#!/bin/bash
mkdir -p gpurun_out
echo "== corr tests"; timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "correlation or launch_counter" > gpurun_out/pytest_gpu5.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_gpu5.log
echo "== ablation"; timeout 300 python tools/dbg_ring.py 2>&1 | grep -E "dbg=|grid cap"
echo "== kbench cold"; rm -f gpurun_out/kbench.jsonl; timeout 600 python tools/kbench.py --what corr --iters 30 --algos mma_bf16x3 --levels 2,3 > gpurun_out/kbench_cold.log 2>&1; echo "rc=$?"; python -c "
import sys, json
for l in open('gpurun_out/kbench_cold.log'):
    d = json.loads(l); print(d['level'], d['kernel'], d.get('algo'), d.get('launched'), 'ms', d['ms_avg'], 'best', d['ms_best'], 'GB/s', d['gbs'], 'frac', d['frac_of_peak'])
"
echo "== ncu ring kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_mma_ring -s 3 -c 1 -f -o gpurun_out/prof_ring python tools/kbench.py --what corr --iters 1 --levels 2 --algos mma_bf16x3 > gpurun_out/ncu_ring.log 2>&1; echo "rc=$?"
