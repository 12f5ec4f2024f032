#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
GSN_PASS_A_PRE=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nopre.json 2> gpurun_out/bench_nopre.err
GSN_TIMELINE_DETAIL=1 timeout 300 python scripts/timeline_detail.py > gpurun_out/timeline_detail.txt 2>&1
timeout 300 python scripts/stage_clocks.py > gpurun_out/stage_clocks.txt 2>&1
tail -5 gpurun_out/pytest_gpu.log; python - <<PY
import json
for f in ("bench_a","bench_nopre"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_share_of_step"], d["roofline"]["instrumented_step_ms"])
    except Exception as e: print("ERR", e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
head -16 gpurun_out/timeline_detail.txt
tail -30 gpurun_out/stage_clocks.txt
