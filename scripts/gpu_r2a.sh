#!/bin/bash
# Round-2 first GPU pass: full parity suite (incl. parity at benchmark sizes), bench + baselines, reference arm on K2,
# sanitizer runs over the tcgen05 kernels, and the launch list the driver would see for smoke().
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep -E "parity-at-size|passed|failed|exit" gpurun_out/${TAG}_pytest_gpu.log | tail -15
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -c 2500 gpurun_out/${TAG}_bench_n1.json; tail -5 gpurun_out/${TAG}_bench_n1.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; tail -c 900 gpurun_out/${TAG}_bench_ref.json; tail -4 gpurun_out/${TAG}_bench_ref.err
for tool in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python scripts/sanitize_block.py > gpurun_out/${TAG}_sanitizer_${tool}.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|shift block ok" gpurun_out/${TAG}_sanitizer_${tool}.log | head -12
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/${TAG}_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_smoke_launches.csv")) if len(r) > 5 and r[0].isdigit()]
c = collections.Counter(r[4].split("(")[0][:60] for r in rows)
print("smoke launches seen by ncu (first 1000):", len(rows))
for k, v in c.most_common(25): print(f"  {v:5d}  {k}")
PY
