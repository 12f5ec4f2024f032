#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "gated_cab or shift_block or stage1 or full_forward or tfr_unet" > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_quick.log
tail -3 gpurun_out/pytest_quick.log
timeout 900 python bench.py --arch gshift_deblur1 --frames 100 --height 1080 --width 1920 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1f_bench_k5.json 2> gpurun_out/r1f_bench_k5.err; cut -c1-300 gpurun_out/r1f_bench_k5.json; grep -o '"hbm_peak_gib": [0-9.]*' gpurun_out/r1f_bench_k5.json; tail -2 gpurun_out/r1f_bench_k5.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cut -c1-200 gpurun_out/bench_a.json; grep -o '"hbm_peak_gib": [0-9.]*' gpurun_out/bench_a.json
