#!/bin/bash
# round-2 GPU batch (edited per batch)
mkdir -p gpurun_out
echo "== conv tests"; timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_network_gpu.py -q -m gpu -x -k "conv or heads or network or cascade or linear" 2>&1 | tail -5
echo "== narrow convs"; timeout 300 python tools/dev_conv_narrow.py 2>&1 | tail -12 | cut -c1-300
echo "== narrow convs, conv_nacc=2"; MFN_TUNING=conv_nacc=2 timeout 300 python tools/dev_conv_narrow.py 2>&1 | tail -12 | cut -c1-120
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample-steps 0 --sustain-seconds 0 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'ms_per_step_eager')}, d['e2e']['value'])"
