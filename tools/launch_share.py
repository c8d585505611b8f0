#!/usr/bin/env python
"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/launch_share.py gpurun_out/launches.csv [first_launch last_launch]

The launch list of `bench.py --steps 2 --warmup 1` holds warm-up, timed and profiled steps back to
back; shares are taken over whatever range is given (default: everything)."""
import csv
import sys
from collections import defaultdict


def main(path, lo=0, hi=None):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    i_name, i_val, i_id = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        k = int(r[i_id])
        if k < lo or (hi is not None and k > hi):
            continue
        name = r[i_name].split("(")[0]
        tot[name] += float(r[i_val].replace(",", "")) / 1e3
        cnt[name] += 1
    s = sum(tot.values())
    print(f"# {path}: launches {lo}..{hi if hi is not None else 'end'}, total {s:.1f} us (cold-cache, serialised under ncu)")
    print(f"{'kernel':40s} {'launches':>8s} {'total_us':>10s} {'share':>7s}")
    for name, t in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{name:40s} {cnt[name]:8d} {t:10.1f} {100 * t / s:6.1f}%")


if __name__ == "__main__":
    a = sys.argv
    main(a[1], int(a[2]) if len(a) > 2 else 0, int(a[3]) if len(a) > 3 else None)
