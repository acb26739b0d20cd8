#!/bin/bash
# round-2 GPU batch (edited per batch)
mkdir -p gpurun_out
echo "== conv tests (conv_as=2)"; MFN_TUNING=conv_as=2 timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "conv" 2>&1 | tail -3
for t in "" "conv_as=2"; do
echo "== bench [$t]"; MFN_TUNING=$t timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample-steps 0 --sustain-seconds 0 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'ms_per_step_eager')}, d['e2e']['value'])"
done
echo "== wide convs"; for t in "" "conv_as=2"; do MFN_TUNING=$t timeout 300 python tools/dev_conv_narrow.py 2>&1 | tail -2 | cut -c1-130; done
