#!/bin/bash
mkdir -p gpurun_out
echo "== kbench cold (device-side timing)"; rm -f gpurun_out/kbench.jsonl; timeout 600 python tools/kbench.py --what corr,warp,bwd --iters 30 > gpurun_out/kbench_cold.log 2>&1; echo "rc=$?"; python -c "
import sys, json
for l in open('gpurun_out/kbench_cold.log'):
    d = json.loads(l); print(d['level'], d['kernel'], d.get('algo'), d.get('launched'), 'ms', d['ms_avg'], 'best', d['ms_best'], 'GB/s', d['gbs'], 'frac', d['frac_of_peak'])
"
echo "== ncu ring kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_mma_ring -s 3 -c 1 -f -o gpurun_out/prof_ring python tools/kbench.py --what corr --iters 1 --levels 2 --algos mma_bf16x3 > gpurun_out/ncu_ring.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_ring.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_simt -s 3 -c 1 -f -o gpurun_out/prof_simt python tools/kbench.py --what corr --iters 1 --levels 2 --algos simt > gpurun_out/ncu_simt.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
