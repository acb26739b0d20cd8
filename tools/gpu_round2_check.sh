#!/bin/bash
mkdir -p gpurun_out
echo "== dev_tma (sigma-permuted B)"; timeout 300 python tools/dev_tma.py 2>&1 | grep -v '"ok": true' | tail -10 | cut -c1-200
echo "== dev_tma v1"; MFN_LIB_PATH=tools/ab/lib_v1.so timeout 300 python tools/dev_tma.py 2>&1 | grep -v '"ok": true' | tail -10 | cut -c1-200
echo "== ncu corr L3 tile kernel"; MFN_TUNING=corr_rb=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:corr_mma_kernel -s 1 -c 1 -o gpurun_out/r02_corr_mma_tile_L3 -f python tools/prof_corr.py --level 3 > gpurun_out/ncu_corr3t.log 2>&1; tail -2 gpurun_out/ncu_corr3t.log
