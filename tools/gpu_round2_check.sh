#!/bin/bash
# round-2 GPU batch (edited per batch)
mkdir -p gpurun_out
echo "== conv + network tests"; timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_network_gpu.py -q -m gpu -x -k "conv or heads or network or cascade or linear or predict" 2>&1 | tail -3
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample-steps 0 --sustain-seconds 0 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'ms_per_step_eager')}, d['e2e']['value'])"
timeout 600 python tools/conv_profile.py > gpurun_out/conv_profile.txt 2>&1; head -22 gpurun_out/conv_profile.txt | cut -c1-150
