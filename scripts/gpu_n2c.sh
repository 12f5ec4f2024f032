#!/bin/bash
# 2-GPU check of gshift_denoise1 in T-sharded mode (bounded: NCCL timeout 90 s inside the script, hard timeouts here)
TAG=${1:-r2j}
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -q -s -k "tshard_two_ranks and denoise1" > gpurun_out/${TAG}_pytest_tshard_denoise1.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_tshard_denoise1.log
grep -E "passed|failed|FAILED|exit|MISMATCH" gpurun_out/${TAG}_pytest_tshard_denoise1.log | tail -4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29561 scripts/tshard_check.py 20 256 448 gshift_denoise1 > gpurun_out/${TAG}_tshard_denoise1_n2.log 2>&1; grep "tshard" gpurun_out/${TAG}_tshard_denoise1_n2.log
