"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total ms, share."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"^(void )?gsn::", "", name)
        agg[name][0] += 1
        agg[name][1] += float(r[vi].replace(",", "")) / 1e6
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot:.2f} ms total (cold-cache, serialised: compare SHARES)")
    print(f"{'kernel':60s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>9s} {'share':>7s}")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:60s} {n:8d} {ms:10.3f} {1000 * ms / n:9.1f} {ms / tot:7.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
