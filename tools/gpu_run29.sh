#!/bin/bash
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv
echo "== profile NEW"; timeout 300 python tools/conv_profile.py 2>&1 | head -12
echo "== profile PREV"; MFN_LIB_PATH=$PWD/tools/lib_prev.so timeout 300 python tools/conv_profile.py 2>&1 | head -12
echo "== profile NEW again"; timeout 300 python tools/conv_profile.py 2>&1 | head -4
echo "== bench NEW"; timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample-steps 0 2>&1 | tail -1 | cut -c1-220
echo "== bench PREV"; MFN_LIB_PATH=$PWD/tools/lib_prev.so timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample-steps 0 2>&1 | tail -1 | cut -c1-220
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv
