#!/bin/bash
# Validation of HEAD on one B200 (shorter than gpu_final.sh): full GPU suite, smoke(),
# the headline bench line with baselines, the T-sharded code path on one rank.
TAG=${1:-r2h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -E "passed|failed|FAILED|exit" gpurun_out/${TAG}_pytest_gpu.log | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; cut -c1-220 gpurun_out/${TAG}_bench_n1.json
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --scaling strong > gpurun_out/${TAG}_bench_strong_n1.json 2> gpurun_out/${TAG}_bench_strong_n1.err; cut -c1-200 gpurun_out/${TAG}_bench_strong_n1.json
