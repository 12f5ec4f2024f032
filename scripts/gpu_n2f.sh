#!/bin/bash
TAG=${1:-r2n}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 35 $TR --master-port 29581 scripts/tshard_check.py 9 96 128 gshift_denoise2 > gpurun_out/${TAG}_tshard_dn2.log 2>&1; echo "exit $?" >> gpurun_out/${TAG}_tshard_dn2.log; grep -E "tshard\] gshift|exit|Error" gpurun_out/${TAG}_tshard_dn2.log | tail -3
timeout 35 $TR --master-port 29582 scripts/tshard_check.py 9 96 128 gshift_deblur1 > gpurun_out/${TAG}_tshard_db1.log 2>&1; echo "exit $?" >> gpurun_out/${TAG}_tshard_db1.log; grep -E "tshard\] gshift|exit|Error" gpurun_out/${TAG}_tshard_db1.log | tail -3
