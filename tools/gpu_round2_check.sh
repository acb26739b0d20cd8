#!/bin/bash
# round-2 GPU batch (edited per batch)
mkdir -p gpurun_out
echo "== warp tests"; timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_network_gpu.py -q -m gpu -x -k "warp or linear or network or cascade" 2>&1 | tail -3
for t in "warp_lin_fch=16" ""; do echo "== kbench warp [$t]"; MFN_TUNING=$t timeout 300 python tools/kbench.py --what warp --levels 2,3,4,5 --iters 20 2>&1 | grep warp_mask | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['level'], d['launched'], d['ms_avg'])"; done
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample-steps 0 --sustain-seconds 0 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'ms_per_step_eager')}, d['e2e']['value'], {k: v['ms'] for k, v in d['roofline']['k3_warp_levels'].items()})"
