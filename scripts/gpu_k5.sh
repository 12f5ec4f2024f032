#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --arch gshift_deblur1 --frames 100 --height 1080 --width 1920 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r1f_bench_k5.json 2> gpurun_out/r1f_bench_k5.err; cut -c1-300 gpurun_out/r1f_bench_k5.json; grep -o '"hbm_peak_gib": [0-9.]*' gpurun_out/r1f_bench_k5.json; tail -2 gpurun_out/r1f_bench_k5.err
