#!/bin/bash
# N-GPU pass (N = number of visible GPUs): weak-scaling bench line under torchrun with NCCL_DEBUG=INFO, strong scaling (one long clip
# T-sharded over all ranks) and the T-shard bit-exactness check.
TAG=${1:-r2c}
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
NCCL_DEBUG=INFO timeout 600 $TR --master-port 29561 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
cut -c1-260 gpurun_out/${TAG}_bench_n${N}.json; grep -E "Init COMPLETE|nranks" gpurun_out/${TAG}_bench_n${N}.err | head -3
timeout 600 $TR --master-port 29562 scripts/tshard_check.py 52 720 1280 > gpurun_out/${TAG}_tshard_n${N}.log 2>&1; grep "tshard" gpurun_out/${TAG}_tshard_n${N}.log | tail -$((N + 1)) | cut -c1-230
timeout 600 $TR --master-port 29563 bench.py --gpus $N --steps 3 --warmup 3 --scaling strong --frames 52 > gpurun_out/${TAG}_bench_strong52_n${N}.json 2> gpurun_out/${TAG}_bench_strong52_n${N}.err; cat gpurun_out/${TAG}_bench_strong52_n${N}.json | cut -c1-1800
timeout 600 $TR --master-port 29564 bench.py --gpus $N --steps 3 --warmup 3 --scaling strong > gpurun_out/${TAG}_bench_strong20_n${N}.json 2> gpurun_out/${TAG}_bench_strong20_n${N}.err; cut -c1-300 gpurun_out/${TAG}_bench_strong20_n${N}.json
