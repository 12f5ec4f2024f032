"""Summarise an `ncu --set full` report: per kernel launch the duration, DRAM bytes, pipe utilisation, issue rate, registers.
Also (re)writes profiles/traffic.json -- the per-launch DRAM traffic of the dominant kernel that bench.py reports as
roofline.traffic -- from the level-1 pass-A launches found in the report, so the number is generated, not hand-edited.

    python scripts/ncu_summary.py gpurun_out/prof_r2.ncu-rep profiles/r2_ncu_summary.txt [--traffic]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [("gpu__time_duration.sum", "ms", 1.0), ("dram__bytes_read.sum", "GB_rd", None), ("dram__bytes_write.sum", "GB_wr", None),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1.0),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1.0), ("launch__registers_per_thread", "regs", 1.0)]
UNIT = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, rows = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    lines = ["# " + os.path.basename(rep) + " (ncu --set full --clock-control none; per-launch values, cold cache, serialised)",
             " | ".join(["kernel"] + [c[1] for c in COLS])]
    launches = []
    for r in rows:
        vals = {}
        for name, short, _ in COLS:
            if name not in hdr:
                vals[short] = None
                continue
            i = hdr.index(name)
            v = float(r[i].replace(",", "")) if r[i] else 0.0
            v *= UNIT.get(units[i], 1.0)
            vals[short] = v
        name = r[ki].replace("void ", "").replace("gsn::", "").split("(")[0]
        launches.append((name, vals))
        lines.append(" | ".join([name] + [("%.4g" % vals[c[1]]) if vals[c[1]] is not None else "-" for c in COLS]))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if "--traffic" in sys.argv:
        big = {}
        for name, v in launches:              # the level-1 launches are the long ones of each pass-A instance
            if name.startswith("cab_pass_a_pre_kernel"):
                key = "cab2" if "<12" in name else "cab1"
                if v["ms"] > big.get(key, (0, None))[0]:
                    big[key] = (v["ms"], (v["GB_rd"] + v["GB_wr"]) * 1e9)
        if len(big) == 2:
            mean_l1 = 0.5 * (big["cab2"][1] + big["cab1"][1])
            tj = {"source": f"scripts/ncu_summary.py over {os.path.basename(rep)} (ncu --set full; level-1 launches T=20, 360x640, C=64, cab_pass_a_pre_kernel)",
                  "cab_pass_a_level1_dram_bytes": {"cab2_pre_normalised": big["cab2"][1], "cab1_pre_normalised": big["cab1"][1]},
                  "note": "bench.py averages over the 48+48 pass-A launches of a step (24+24 at level 1, 24+24 at level 2 = 1/4 of the pixels each): per-launch mean = 0.625 * mean(level-1 values)",
                  "cab_pass_a_bytes_per_launch": 0.625 * mean_l1}
            json.dump(tj, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
            print("traffic.json updated:", tj["cab_pass_a_bytes_per_launch"])


if __name__ == "__main__":
    main()
