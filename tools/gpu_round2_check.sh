#!/bin/bash
mkdir -p gpurun_out
echo "== full pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6
echo "== kbench bwd"; timeout 300 python tools/kbench.py --what bwd --levels 2,3 --iters 10 2>&1 | grep -v "^$" | cut -c1-260
echo "== conv profile"; timeout 300 python tools/conv_profile.py 2>&1 | head -48
