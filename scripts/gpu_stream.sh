#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stream.py -x -q -s > gpurun_out/stream_test.log 2>&1; echo "pytest exit $?" >> gpurun_out/stream_test.log
grep -E "\[stream\]|passed|failed|Error|error|exit|trap|illegal" gpurun_out/stream_test.log | head -60
if grep -q "pytest exit 0" gpurun_out/stream_test.log; then
  GSN_PASS_A_STREAM=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/stream_bench.json 2> gpurun_out/stream_bench.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/stream_bench.json")); r=d["roofline"]
    print("bench", d["value"], d["ms_per_step"], "frac", r["frac"], "avg_launch_ms", r["avg_launch_ms"], r["kernel_share_of_step"], "block", r.get("block",{}).get("frac"))
except Exception as e: print("ERR", e, open("gpurun_out/stream_bench.err").read()[-1500:])
PY
fi
