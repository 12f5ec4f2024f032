#!/bin/bash
# 2-GPU pass after GSN_ROLL_HALO: the two-rank tests (T-shard bit-exact x4 nets, inference entry point) and the strong-scaling bench line
TAG=${1:-r2l}
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -s -k "two_ranks" > gpurun_out/${TAG}_pytest_n2.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_n2.log
grep -E "passed|failed|FAILED|exit|MISMATCH" gpurun_out/${TAG}_pytest_n2.log | tail -6
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29553 bench.py --gpus 2 --steps 5 --warmup 3 --scaling strong > gpurun_out/${TAG}_bench_strong_n2.json 2> gpurun_out/${TAG}_bench_strong_n2.err; cat gpurun_out/${TAG}_bench_strong_n2.json; tail -2 gpurun_out/${TAG}_bench_strong_n2.err
