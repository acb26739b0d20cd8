#!/bin/bash
mkdir -p gpurun_out
echo "== warp tests"; timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "warp or deform or real_checkpoint" 2>&1 | tail -3
echo "== ablation"; timeout 300 python tools/dbg_ring.py 2>&1 | grep -E "dbg=|grid cap|copy"
echo "== kbench warp"; rm -f gpurun_out/kbench.jsonl; timeout 600 python tools/kbench.py --what warp --iters 20 > gpurun_out/kbench_warp.log 2>&1; python -c "
import sys, json
for l in open('gpurun_out/kbench_warp.log'):
    d = json.loads(l); print(d['level'], d['kernel'], 'ms', d['ms_avg'], 'TFLOP/s', d.get('tflops'))
"
echo "== train bench cfg3"; timeout 600 python tools/train_bench.py --hw 384x512 --batch 8 --steps 5 2>&1 | tail -1
