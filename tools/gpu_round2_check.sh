#!/bin/bash
# round-2 GPU batch (edited per batch): hybrid training convolution, geometry kernel with the v/255 table, ncu of the new kernels
mkdir -p gpurun_out
echo "== train-side tests"; timeout 300 python -m pytest tests/test_train_side.py -q -m gpu 2>&1 | tail -8 | tee gpurun_out/check_tests2.log
echo "== train-side bench"; timeout 120 python tools/train_side_bench.py > gpurun_out/train_side_bench2.jsonl 2> gpurun_out/train_side_bench2.err; cut -c1-200 gpurun_out/train_side_bench2.jsonl; tail -3 gpurun_out/train_side_bench2.err
echo "== bench fwdbwd tc forward"; timeout 240 python bench.py --config fwdbwd --train-tc-forward 1 --steps 5 --warmup 3 --cpu-sample-steps 0 --sustain-seconds 0 2>&1 | tail -1 > gpurun_out/bench_fwdbwd_tc.json; cut -c1-330 gpurun_out/bench_fwdbwd_tc.json
echo "== ncu"; timeout 240 ncu --set full --clock-control none --import-source on -k regex:"geometry_augment|color_|epe_" -c 14 -f -o gpurun_out/prof_train_side python tools/prof_train_side.py > gpurun_out/ncu_train_side.log 2>&1; echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
