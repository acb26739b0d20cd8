#!/bin/bash
# round-2 GPU batch (edited per batch)
mkdir -p gpurun_out
echo "== full gpu tests"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample-steps 0 --sustain-seconds 0 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'ms_per_step_eager')}, d['e2e']['value'])"
