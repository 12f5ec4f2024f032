#!/bin/bash
# 2-GPU pass: the two-rank tests (inference entry point with NCCL all_gather, T-sharded clip bit-exact), the T-shard check at the
# benchmark size with halo statistics, and bench.py in weak and strong scaling under torchrun with NCCL_DEBUG=INFO.
TAG=${1:-r2b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_n2_smi.txt
timeout 900 python -m pytest tests -m gpu -q -s -k "two_ranks" > gpurun_out/${TAG}_pytest_n2.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_n2.log
tail -4 gpurun_out/${TAG}_pytest_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29551 scripts/tshard_check.py 20 720 1280 > gpurun_out/${TAG}_tshard_n2.log 2>&1; grep "tshard" gpurun_out/${TAG}_tshard_n2.log
NCCL_DEBUG=INFO timeout 600 $TR --master-port 29552 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; cut -c1-300 gpurun_out/${TAG}_bench_n2.json; grep -c "NCCL INFO" gpurun_out/${TAG}_bench_n2.err; grep -E "nranks|NVLS|P2P" gpurun_out/${TAG}_bench_n2.err | head -5
timeout 600 $TR --master-port 29553 bench.py --gpus 2 --steps 5 --warmup 3 --scaling strong > gpurun_out/${TAG}_bench_strong_n2.json 2> gpurun_out/${TAG}_bench_strong_n2.err; cat gpurun_out/${TAG}_bench_strong_n2.json; tail -3 gpurun_out/${TAG}_bench_strong_n2.err
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --scaling strong > gpurun_out/${TAG}_bench_strong_n1.json 2> gpurun_out/${TAG}_bench_strong_n1.err; cut -c1-420 gpurun_out/${TAG}_bench_strong_n1.json
