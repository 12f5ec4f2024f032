#!/bin/bash
# 2-GPU pass: T-shard tests (all four nets), the UMMA issue-cost micro-benchmark
TAG=${1:-r2i}
mkdir -p gpurun_out
timeout 120 scripts/_bin/ubench_umma > gpurun_out/${TAG}_ubench_umma.txt 2>&1; cat gpurun_out/${TAG}_ubench_umma.txt
timeout 1200 python -m pytest tests -m gpu -q -s -k "tshard" > gpurun_out/${TAG}_pytest_tshard.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_tshard.log
grep -E "passed|failed|FAILED|Error|exit|MISMATCH" gpurun_out/${TAG}_pytest_tshard.log | tail -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29561 scripts/tshard_check.py 20 256 448 gshift_denoise1 > gpurun_out/${TAG}_tshard_denoise1_n2.log 2>&1; grep "tshard" gpurun_out/${TAG}_tshard_denoise1_n2.log
