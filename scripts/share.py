"""Print the per-kernel shares of a bench.py JSON line (stdin)."""
import json
import sys

d = json.loads(sys.stdin.read())
r = d["roofline"]
print(d["config"]["workload"][:60], round(d["value"], 2), "fps", round(d["ms_per_step"], 1), "ms; instrumented", round(r["instrumented_step_ms"], 1))
for k, v in sorted(r["kernel_share_of_step"].items(), key=lambda kv: -kv[1])[:24]:
    print(f"  {k:48s} {100 * v:5.1f}%  {v * r['instrumented_step_ms']:7.1f} ms")
