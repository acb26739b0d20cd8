#!/bin/bash
mkdir -p gpurun_out
echo "== ncu umma 579->128"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_umma -s 2 -c 1 -f -o gpurun_out/prof_umma_n128 python tools/prof_umma.py 579 128 > gpurun_out/ncu_umma128.log 2>&1; echo "rc=$?"
echo "== ncu umma 547->32"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_umma -s 2 -c 1 -f -o gpurun_out/prof_umma_n32 python tools/prof_umma.py 547 32 > gpurun_out/ncu_umma32.log 2>&1; echo "rc=$?"
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample-steps 0 2>&1 | tail -1 | cut -c1-220
ls -la gpurun_out/*.ncu-rep
