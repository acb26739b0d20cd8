#!/bin/bash
# round-2 GPU batch (edited per batch): the training-side rows (N4 augmentation, fused MultiscaleEpe) + the last warp_lin change
mkdir -p gpurun_out
echo "== new + touched tests"; timeout 420 python -m pytest tests/test_train_side.py tests/test_network_gpu.py tests/test_ops_gpu.py -q -m gpu -k "augment or epe or warp or linear or abi" 2>&1 | tail -15 | tee gpurun_out/check_tests.log
echo "== train-side bench"; timeout 120 python tools/train_side_bench.py > gpurun_out/train_side_bench.jsonl 2> gpurun_out/train_side_bench.err; cat gpurun_out/train_side_bench.jsonl; tail -3 gpurun_out/train_side_bench.err
echo "== bench fwdbwd"; timeout 240 python bench.py --config fwdbwd --steps 5 --warmup 3 --cpu-sample-steps 0 --sustain-seconds 0 2>&1 | tail -1 > gpurun_out/bench_fwdbwd.json; cut -c1-400 gpurun_out/bench_fwdbwd.json
echo "== bench"; timeout 300 python bench.py --steps 10 --warmup 3 --cpu-sample-steps 0 --sustain-seconds 0 2>&1 | tail -1 > gpurun_out/bench_ours_short.json; python -c "
import sys, json
d = json.loads(open('gpurun_out/bench_ours_short.json').read())
print({k: d[k] for k in ('value', 'ms_per_step', 'ms_per_step_eager')}, d['e2e']['value'], {k: v['ms'] for k, v in d['roofline']['k3_warp_levels'].items()})"
