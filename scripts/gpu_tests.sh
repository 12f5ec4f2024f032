#!/bin/bash
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -E "parity-at-size|passed|failed|FAILED|Error|exit" gpurun_out/${TAG}_pytest_gpu.log | tail -25
