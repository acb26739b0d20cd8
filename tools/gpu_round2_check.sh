#!/bin/bash
mkdir -p gpurun_out
echo "== full pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== conv profile"; timeout 300 python tools/conv_profile.py 2>&1 | head -40
echo "== bench cascade"; timeout 600 python bench.py --config cascade --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r02_bench_cascade_1gpu.json | cut -c1-700
echo "== bench fwdbwd"; timeout 600 python bench.py --config fwdbwd --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r02_bench_fwdbwd_1gpu.json | cut -c1-700
