#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s -k "conv3x3 or (shift_block and denoise1) or pw_gate or ln_pw or conv3x3 or (gated_cab and deblur1) or (gated_cab and denoise1) or (golden and denoise1) or (tfr and deblur1) or (tfr and denoise1) or K3 or (golden and deblur1) or K5shape" > gpurun_out/c3_test.log 2>&1; echo "pytest exit $?" >> gpurun_out/c3_test.log
grep -E "conv3x3|parity-at-size|TFR|passed|failed|Error|error|exit" gpurun_out/c3_test.log | head -30
if grep -q "pytest exit 0" gpurun_out/c3_test.log; then
  for tc in 1; do
    GSN_CONV_TC=$tc timeout 600 python bench.py --arch gshift_deblur1 --frames 52 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c3_k3_tc$tc.json 2> gpurun_out/c3_k3_tc$tc.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c3_k3_tc$tc.json")); r=d["roofline"]
    print("K3 GSN_CONV_TC=$tc:", round(d["value"],2), "fps", round(d["ms_per_step"],1), "ms", r["kernel_share_of_step"])
except Exception as e: print("ERR", e, open("gpurun_out/c3_k3_tc$tc.err").read()[-1500:])
PY
  done
fi
