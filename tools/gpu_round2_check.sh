#!/bin/bash
# round-2 GPU batch (edited per batch): final measurements
mkdir -p gpurun_out
echo "== conv tests"; timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "conv" 2>&1 | tail -3
echo "== default bench"; ( time timeout 900 python bench.py ) 2>&1 | tail -6 | tee gpurun_out/r02_bench_default.txt | cut -c1-300
echo "== reference arm"; ( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) 2>&1 | tail -5 | tee gpurun_out/r02_bench_reference.txt | cut -c1-200
echo "== cascade"; timeout 600 python bench.py --config cascade --cpu-sample-steps 0 2>&1 | tail -1 | tee gpurun_out/r02_bench_cascade_1gpu.json | cut -c1-300
echo "== fwdbwd"; timeout 600 python bench.py --config fwdbwd --cpu-sample-steps 0 2>&1 | tail -1 | tee gpurun_out/r02_bench_fwdbwd_1gpu.json | cut -c1-300
