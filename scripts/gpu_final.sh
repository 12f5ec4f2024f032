#!/bin/bash
# Round-end measurement set (1 GPU): full parity suite, the headline bench line (with cpu_baseline), the reference arm, and the
# other BASELINE.json configs that fit one GPU (K3 Ours+ 720p one_len=48; K5's per-GPU share: Ours+ 1080p one_len=96).
TAG=${1:-r1f}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -c 1200 gpurun_out/${TAG}_bench_n1.json
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; tail -c 600 gpurun_out/${TAG}_bench_ref.json
timeout 600 python bench.py --arch gshift_deblur1 --frames 52 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_k3.json 2> gpurun_out/${TAG}_bench_k3.err; cut -c1-420 gpurun_out/${TAG}_bench_k3.json; tail -3 gpurun_out/${TAG}_bench_k3.err
timeout 900 python bench.py --arch gshift_deblur1 --frames 100 --height 1080 --width 1920 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_k5.json 2> gpurun_out/${TAG}_bench_k5.err; cut -c1-420 gpurun_out/${TAG}_bench_k5.json; tail -3 gpurun_out/${TAG}_bench_k5.err
timeout 600 python bench.py --arch gshift_denoise2 --frames 68 --height 272 --width 448 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_k4tile.json 2> gpurun_out/${TAG}_bench_k4tile.err; cut -c1-420 gpurun_out/${TAG}_bench_k4tile.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.txt
