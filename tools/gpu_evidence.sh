#!/bin/bash
# Evidence run (one GPU): ncu launch list of the bench's timed region + one `--set full` capture per hot-path kernel.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() {  # name kernel-regex skip args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 $NCU -k regex:$rx -s $skip -c 1 -o gpurun_out/r02_$name -f python tools/prof_ops.py "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
}
echo "== kbench bwd"; timeout 300 python tools/kbench.py --what bwd --levels 2,3 --iters 10 2>&1 | grep -v "^$" | cut -c1-200
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_bench_timed_region.csv python bench.py --steps 2 --warmup 3 --cpu-sample-steps 0 --sustain-seconds 0 > gpurun_out/r02_launches_bench.log 2>&1; tail -1 gpurun_out/r02_launches_bench.log | cut -c1-200
cap corr_tma_L2 corr_tma_kernel 1 corr --level 2
cap corr_tile_L3 corr_mma_kernel 1 corr --level 3
cap corr_tile_L4 corr_mma_kernel 1 corr --level 4
cap corr_rb_L5 corr_rb_kernel 1 corr --level 5
cap corr_rb_L6 corr_rb_kernel 1 corr --level 6
cap corr_tma_md2_L2 corr_tma_kernel 1 cascade_corr --level 2
cap warp_lin_L2 warp_lin_kernel 1 warp --level 2
cap warp_lin_L3 warp_lin_kernel 1 warp --level 3
cap warp_conv_ext_L3 conv3x3_umma_kernel 1 warp --level 3
cap warp_lin_L5 warp_lin_kernel 1 warp --level 5
cap k5_image_warp image_warp_concat_kernel 1 k5
cap corr_bwd_L2 corr_bwd_kernel 2 corr_bwd --level 2
cap warp_bwd_input_L3 deform_bwd_input_kernel 1 warp_bwd --level 3
cap warp_bwd_weight_L3 deform_bwd_weight 1 warp_bwd --level 3
ls -la gpurun_out/r02_*.ncu-rep | wc -l
