#!/bin/bash
# Round artefacts: (1) ncu launch list (gpu__time_duration) of exactly one timed bench step, (2) ncu --set full (with
# source-level sampling) of the hot kernels inside a timed step.  Numbers printed by bench.py under ncu are NOT bench values.
TAG=${1:-r1f}
mkdir -p gpurun_out
GSN_CUDA_GRAPH=0 GSN_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
  --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${TAG}_list.log 2>&1
GSN_CUDA_GRAPH=0 GSN_NCU_RANGE=1 timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on \
  --kernel-name regex:"cab_pass_a_pre|cab_pass_b|shift_conv1|cab_dense" --launch-skip 46 --launch-count 16 -o gpurun_out/prof_${TAG} -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_${TAG}.log; ls -la gpurun_out/prof_${TAG}.ncu-rep gpurun_out/${TAG}_launches.csv
