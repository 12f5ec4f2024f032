#!/bin/bash
# 2-GPU one-shot: T-sharded forward as CUDA graphs cut at the halo exchanges -- bit-exactness (Ours-s at the bench size, denoise1 small)
# and the strong-scaling bench line
TAG=${1:-r2m}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 80 $TR --master-port 29571 scripts/tshard_check.py 20 720 1280 gshift_deblur2 > gpurun_out/${TAG}_tshard_k2.log 2>&1; echo "exit $?" >> gpurun_out/${TAG}_tshard_k2.log; grep -E "tshard|exit|Error|error" gpurun_out/${TAG}_tshard_k2.log | tail -5
timeout 60 $TR --master-port 29572 scripts/tshard_check.py 9 96 128 gshift_denoise1 > gpurun_out/${TAG}_tshard_dn1.log 2>&1; echo "exit $?" >> gpurun_out/${TAG}_tshard_dn1.log; grep -E "tshard\]|exit|Error" gpurun_out/${TAG}_tshard_dn1.log | tail -3
timeout 100 $TR --master-port 29573 bench.py --gpus 2 --steps 5 --warmup 3 --scaling strong > gpurun_out/${TAG}_bench_strong_n2.json 2> gpurun_out/${TAG}_bench_strong_n2.err; cut -c1-200 gpurun_out/${TAG}_bench_strong_n2.json; grep -o '"sharding": "[^"]*"' gpurun_out/${TAG}_bench_strong_n2.json; grep -o '"halo": {[^}]*}' gpurun_out/${TAG}_bench_strong_n2.json; tail -2 gpurun_out/${TAG}_bench_strong_n2.err
