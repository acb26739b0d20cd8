#!/bin/bash
# First GPU session: micro-benchmarks, isolation smoke of the correlation kernels, GPU test-suite, kernel benchmark.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> gpurun_out/nproc.txt
echo "== ubench"; timeout 120 ./tools/ubench > gpurun_out/ubench.jsonl 2>&1; tail -20 gpurun_out/ubench.jsonl
echo "== smoke corr"; timeout 180 python tools/smoke_corr.py > gpurun_out/smoke_corr.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/smoke_corr.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_ops_gpu.py::test_correlation_full_size_properties > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_gpu.log
echo "== pytest gpu (rest, no -x)"; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/pytest_gpu_all.log
echo "== kbench cold"; rm -f gpurun_out/kbench.jsonl; timeout 600 python tools/kbench.py --what corr,warp,bwd --iters 20 > gpurun_out/kbench_cold.log 2>&1; echo "rc=$?"; cat gpurun_out/kbench_cold.log | cut -c1-400
echo "== kbench warm"; timeout 600 python tools/kbench.py --what corr --iters 20 --warm > gpurun_out/kbench_warm.log 2>&1; echo "rc=$?"; cat gpurun_out/kbench_warm.log | cut -c1-400
