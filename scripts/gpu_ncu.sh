#!/bin/bash
# ncu --set full (with source-level stall sampling) of the hot kernels inside one timed bench step.
mkdir -p gpurun_out
GSN_CUDA_GRAPH=0 GSN_NCU_RANGE=1 timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on \
  --kernel-name regex:"cab_pass_a_tc|cab_dense|cab_pass_b|shift_conv1" --launch-skip 40 --launch-count 14 -o gpurun_out/prof_r1e -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r1e.log 2>&1
tail -5 gpurun_out/ncu_r1e.log; ls -la gpurun_out/prof_r1e.ncu-rep
