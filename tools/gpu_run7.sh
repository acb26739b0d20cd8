#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== bench reference (2 ranks launch: rank1 must exit 0)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | tail -2 | cut -c1-400
echo "== bench ours 2 gpus"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_ours_2gpu.json | cut -c1-700
