#!/bin/bash
mkdir -p gpurun_out
echo "== dev_tma (compact + rolled producer/storer loops)"; timeout 300 python tools/dev_tma.py 2>&1 | grep -v '"ok": true' | grep -v "corr_mma_ring" | tail -8 | cut -c1-200
