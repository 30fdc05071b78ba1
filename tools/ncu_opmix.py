#!/usr/bin/env python3
"""Executed-instruction mix by SASS opcode from `ncu -i X.ncu-rep --page source --csv` output (one or more kernels).
   python tools/ncu_opmix.py source.csv [top]"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    hdr = None
    cnt = collections.Counter()
    samp = collections.Counter()
    name = None

    def flush():
        tot = sum(cnt.values())
        if not tot:
            return
        print("kernel:", name, "| warp-instructions executed:", tot)
        for k, v in cnt.most_common(top):
            print(f"  {k:28s} {v:12d} {100 * v / tot:6.2f}%  stall samples {samp[k]}")

    for r in rows:
        if r and r[0] == "Kernel Name":
            flush()
            cnt.clear(); samp.clear()
            name = r[1][:110]
            continue
        if r and r[0] == "Address":
            hdr = r
            ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) <= ie:
            continue
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ia].strip())
        if not m:
            continue
        try:
            n = int(float(r[ie] or 0))
        except ValueError:
            continue
        cnt[m.group(2)] += n
        samp[m.group(2)] += int(float(r[isamp] or 0))
    flush()


if __name__ == "__main__":
    main()
