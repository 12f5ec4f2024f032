#!/bin/bash
# Validation of HEAD on one B200 (shorter than gpu_final.sh): full GPU suite, smoke(), racecheck over the Ours+ tcgen05 kernels,
# the headline bench line with baselines, K3.
TAG=${1:-r2h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -E "passed|failed|FAILED|exit" gpurun_out/${TAG}_pytest_gpu.log | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 30 python scripts/sanitize_block.py gshift_deblur1 > gpurun_out/${TAG}_sanitizer_plus_racecheck.log 2>&1
echo "== racecheck (Ours+) exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|ok" gpurun_out/${TAG}_sanitizer_plus_racecheck.log | head -6
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; cut -c1-220 gpurun_out/${TAG}_bench_n1.json
timeout 600 python bench.py --arch gshift_deblur1 --frames 52 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_k3.json 2> gpurun_out/${TAG}_bench_k3.err; cut -c1-160 gpurun_out/${TAG}_bench_k3.json
