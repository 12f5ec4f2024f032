#!/bin/bash
# Round artefacts (1 GPU): (1) the headline bench line with baselines, (2) ncu launch list of exactly one timed bench step,
# (3) ncu --set full of the block's kernels (tile pass A default, then the streaming pass A), (4) the K5 one_len sweep, K3.
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; cut -c1-200 gpurun_out/${TAG}_bench_n1.json
GSN_PASS_A_STREAM=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_stream.json 2> gpurun_out/${TAG}_bench_stream.err; cut -c1-200 gpurun_out/${TAG}_bench_stream.json
GSN_CUDA_GRAPH=0 GSN_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
  --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${TAG}_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cab_pass_a_pre|cab_pass_b_tc|shift_conv1_ln|cab_fold" -s 10 -c 10 \
  -o gpurun_out/prof_${TAG} -f python scripts/ncu_block.py > gpurun_out/ncu_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_${TAG}.log
GSN_PASS_A_STREAM=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cab_pass_a_stream" -s 2 -c 2 \
  -o gpurun_out/prof_${TAG}_stream -f python scripts/ncu_block.py > gpurun_out/ncu_${TAG}_stream.log 2>&1; tail -1 gpurun_out/ncu_${TAG}_stream.log
timeout 600 python bench.py --arch gshift_deblur1 --frames 52 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_k3.json 2> gpurun_out/${TAG}_bench_k3.err; cut -c1-160 gpurun_out/${TAG}_bench_k3.json
for ol in 8 12 16 24 32 48 96; do
  timeout 900 python bench.py --arch gshift_deblur1 --frames $((ol + 4)) --height 1080 --width 1920 --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_bench_k5_ol${ol}.json 2> gpurun_out/${TAG}_bench_k5_ol${ol}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_k5_ol${ol}.json"))
    print("K5 one_len=${ol}: %.2f frames/s  %.1f ms/clip  e2e %.2f  peak %.1f GiB" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["hbm_peak_gib"]))
except Exception as e: print("K5 one_len=${ol} ERR", e)
PY
done
timeout 300 python bench.py --arch gshift_denoise2 --frames 68 --height 272 --width 448 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_k4tile.json 2> gpurun_out/${TAG}_bench_k4tile.err; cut -c1-160 gpurun_out/${TAG}_bench_k4tile.json
