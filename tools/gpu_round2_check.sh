#!/bin/bash
# round-2 GPU batch (edited per batch)
mkdir -p gpurun_out
echo "== dev_rb sweep"; timeout 600 python tools/dev_rb.py 2>&1 | tail -40 | cut -c1-300
