#!/bin/bash
mkdir -p gpurun_out
# conv launches per forward: 27 (+27 packs in the first).  Skip the first forward (54 launches of mfn kernels named conv3x3*), take
# the 5th conv of the second forward = conv6_4? order: level 6 (5 convs), level 5 ...; we want a level-2 conv: index 20..24 -> skip 27+21
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_mma_kernel --launch-skip 48 --launch-count 1 -f -o gpurun_out/prof_conv python tools/prof_fwd.py 2 > gpurun_out/ncu_conv.log 2>&1; echo "rc=$?"; grep -E "conv3x3|==PROF" gpurun_out/ncu_conv.log | tail -3
