#!/bin/bash
mkdir -p gpurun_out
echo "== dev_tma base (sigma)"; timeout 300 python tools/dev_tma.py 2>&1 | grep -v '"ok": true' | grep -v "corr_mma_ring" | tail -8 | cut -c1-200
echo "== dev_tma pair-interleaved"; MFN_LIB_PATH=tools/ab/lib_pair.so timeout 300 python tools/dev_tma.py 2>&1 | grep -v '"ok": true' | grep -v "corr_mma_ring" | tail -8 | cut -c1-200
