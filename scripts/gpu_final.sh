#!/bin/bash
# Round-end validation of HEAD on one B200: full GPU suite, smoke(), sanitizer passes over the tcgen05 kernels (incl. the streaming
# pass A and the Ours+ kernels), the headline bench line with baselines, K3, the K5 end points.
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -E "passed|failed|FAILED|exit" gpurun_out/${TAG}_pytest_gpu.log | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
for tool in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python scripts/sanitize_block.py gshift_deblur1 > gpurun_out/${TAG}_sanitizer_plus_${tool}.log 2>&1
  echo "== $tool (Ours+) exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|ok" gpurun_out/${TAG}_sanitizer_plus_${tool}.log | head -6
done
GSN_PASS_A_STREAM=1 timeout 900 compute-sanitizer --tool racecheck --print-limit 30 python scripts/sanitize_block.py gshift_deblur2 > gpurun_out/${TAG}_sanitizer_stream_racecheck.log 2>&1
echo "== racecheck (streaming pass A) exit $?"; grep -E "RACECHECK SUMMARY|hazard|ok" gpurun_out/${TAG}_sanitizer_stream_racecheck.log | head -4
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; cut -c1-220 gpurun_out/${TAG}_bench_n1.json
timeout 600 python bench.py --arch gshift_deblur1 --frames 52 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_k3.json 2> gpurun_out/${TAG}_bench_k3.err; cut -c1-160 gpurun_out/${TAG}_bench_k3.json
for ol in 8 96; do
  timeout 900 python bench.py --arch gshift_deblur1 --frames $((ol + 4)) --height 1080 --width 1920 --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_bench_k5_ol${ol}.json 2> gpurun_out/${TAG}_bench_k5_ol${ol}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_k5_ol${ol}.json"))
    print("K5 one_len=${ol}: %.2f frames/s  %.1f ms/clip  e2e %.2f  peak %.1f GiB" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["hbm_peak_gib"]))
except Exception as e: print("K5 one_len=${ol} ERR", e)
PY
done
