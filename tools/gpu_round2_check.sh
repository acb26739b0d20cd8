#!/bin/bash
mkdir -p gpurun_out
echo "== dev_tma"; timeout 300 python tools/dev_tma.py 2>&1 | grep -v '"ok": true' | tail -12
echo "== ncu corr L2"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:corr_tma_kernel -s 1 -c 1 -o gpurun_out/r02_corr_tma_L2_v2 -f python tools/prof_corr.py > gpurun_out/ncu_corr.log 2>&1; tail -2 gpurun_out/ncu_corr.log
echo "== ncu corr L3 rb"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:corr_rb_kernel -s 1 -c 1 -o gpurun_out/r02_corr_rb_L3 -f python tools/prof_corr.py --level 3 > gpurun_out/ncu_corr3.log 2>&1; tail -2 gpurun_out/ncu_corr3.log
echo "== ncu corr bwd"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:corr_bwd_kernel -s 2 -c 1 -o gpurun_out/r02_corr_bwd_L2 -f python tools/kbench.py --what bwd --levels 2 --iters 3 > gpurun_out/ncu_bwd.log 2>&1; tail -2 gpurun_out/ncu_bwd.log
