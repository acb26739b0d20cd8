#!/bin/bash
mkdir -p gpurun_out
echo "== conv + network tests"; timeout 900 python -m pytest tests -m gpu -q -x -k "conv3x3 or network or product or predict or cascade or training or shim or reference_file" 2>&1 | tail -8
echo "== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['roofline']['hot_path_ms_per_step'], d['roofline']['per_kernel_ms'], d.get('cpu_baseline'))
"
