"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: device time per kernel, shares, per-launch mean.
Usage: python tools/summarize_launches.py gpurun_out/launches.csv [--exclude k_scalar_mul_u64] > profiles/launches_<round>_summary.txt"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    exclude = sys.argv[3].split(",") if len(sys.argv) > 3 and sys.argv[2] == "--exclude" else ["k_scalar_mul_u64"]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    header = next(rd)
    ki, mi, vi, ui = header.index("Kernel Name"), header.index("Metric Name"), header.index("Metric Value"), header.index("Metric Unit")
    for r in rd:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)   # -> microseconds
        rows.append((r[ki], v * scale))
    agg, cnt = defaultdict(float), defaultdict(int)
    for name, us in rows:
        short = re.sub(r"\(.*$", "", name)
        if any(e in short for e in exclude):
            continue
        agg[short] += us
        cnt[short] += 1
    total = sum(agg.values())
    print("per-kernel device time summed over the captured launches (cold-cache, serialised: compare SHARES); excluded:", ",".join(exclude))
    for name, us in sorted(agg.items(), key=lambda kv: -kv[1]):
        print(f"{us / 1e3:10.3f} ms {cnt[name]:5d} launches {100 * us / total:6.1f}% {us / cnt[name]:10.1f} us/launch  {name[:110]}")
    print(f"{total / 1e3:10.3f} ms total")


if __name__ == "__main__":
    main()
