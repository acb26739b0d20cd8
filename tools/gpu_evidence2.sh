#!/bin/bash
# Final evidence run of round 2 (one GPU): launch list of the bench's timed region + `--set full` captures of the kernels
# that changed in the second half of the round.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() {  # name kernel-regex skip args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 $NCU -k regex:$rx -s $skip -c 1 -o gpurun_out/r02_$name -f python tools/prof_ops.py "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
}
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_bench_timed_region.csv python bench.py --steps 2 --warmup 3 --cpu-sample-steps 0 --sustain-seconds 0 > gpurun_out/r02_launches_bench.log 2>&1; tail -1 gpurun_out/r02_launches_bench.log | cut -c1-200
cap corr_tma_L2 corr_tma_kernel 1 corr --level 2
cap warp_lin_L2 warp_lin_kernel 1 warp --level 2
cap conv_narrow_16_16 conv3x3_umma_kernel 1 conv --conv 16,16,224,512,16,1
cap conv_wide_579_128 conv3x3_umma_kernel 1 conv --conv 579,128,112,256,8,1
cap conv_splitk_675_64 conv3x3_umma_kernel 1 conv --conv 675,64,14,32,8,1
cap conv_reduce_675_64 conv3x3_umma_reduce_kernel 1 conv --conv 675,64,14,32,8,1
ls -la gpurun_out/r02_*.ncu-rep | wc -l
