#!/bin/bash
# final artefacts of the round: parity lines of the whole GPU suite, the headline bench line, K3, pass-A stage clocks
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_s.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_s.log
grep "\[parity\]" gpurun_out/pytest_gpu_s.log > gpurun_out/r1f_parity.txt; tail -3 gpurun_out/pytest_gpu_s.log; grep -c parity gpurun_out/r1f_parity.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r1f_bench_n1.json 2> gpurun_out/r1f_bench_n1.err; cut -c1-260 gpurun_out/r1f_bench_n1.json
timeout 600 python bench.py --arch gshift_deblur1 --frames 52 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1f_bench_k3.json 2> gpurun_out/r1f_bench_k3.err; cut -c1-260 gpurun_out/r1f_bench_k3.json; tail -2 gpurun_out/r1f_bench_k3.err
