#!/bin/bash
mkdir -p gpurun_out
echo "== long-run tests"; timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "long_tile_runs or parity or full_size or leaky" > gpurun_out/pytest_gpu2.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu2.log
echo "== kbench cold"; rm -f gpurun_out/kbench.jsonl; timeout 600 python tools/kbench.py --what corr --iters 30 > gpurun_out/kbench_cold.log 2>&1; echo "rc=$?"; grep -v generic gpurun_out/kbench_cold.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['level'], d['algo'], d.get('launched'), 'ms', d['ms_avg'], 'best', d['ms_best'], 'GB/s', d['gbs'], 'frac', d['frac_of_peak'], 'err', d.get('max_abs_diff_vs_first'))
"
