#!/bin/bash
# final artefacts of the round: parity lines of the whole GPU suite, the headline bench line, pass-A stage clocks
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_s.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_s.log
grep "\[parity\]" gpurun_out/pytest_gpu_s.log > gpurun_out/r1f_parity.txt; tail -3 gpurun_out/pytest_gpu_s.log; grep -c parity gpurun_out/r1f_parity.txt; grep "full forward\|K1\|k1" gpurun_out/r1f_parity.txt | head -12
timeout 900 python bench.py > gpurun_out/r1f_bench_n1.json 2> gpurun_out/r1f_bench_n1.err; cut -c1-260 gpurun_out/r1f_bench_n1.json
timeout 300 python scripts/stage_clocks.py > gpurun_out/r1f_stage_clocks.txt 2>&1; head -18 gpurun_out/r1f_stage_clocks.txt
