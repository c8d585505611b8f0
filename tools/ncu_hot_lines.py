#!/usr/bin/env python
"""Instructions of a kernel that collect the most warp-stall samples, from the source page of an
Nsight Compute report taken with --import-source on (the SASS view; lines carry their sample share
and execution count).  This is how the exposed-latency spots of the guided filter were found
(a look-up depending on a load issued one instruction earlier).

    python tools/ncu_hot_lines.py gpurun_out/r02f_k_guided_ab.ncu-rep [min_share_percent] [context_lines]
"""
import csv
import subprocess
import sys


def main(path, min_share=1.5, ctx=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    name = rows[0][1] if rows and len(rows[0]) > 1 else "?"
    hdr = rows[1]
    i_src, i_smp, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    body = rows[2:]
    tot = sum(int(r[i_smp]) for r in body) or 1
    print(f"# {name}: {len(body)} SASS instructions, {tot} stall samples")
    hot = [i for i, r in enumerate(body) if 100.0 * int(r[i_smp]) / tot >= min_share]
    shown = set()
    for i in hot:
        for j in range(max(0, i - ctx), min(len(body), i + ctx + 1)):
            if j in shown:
                continue
            shown.add(j)
            r = body[j]
            print(f"{j:5d} {100.0 * int(r[i_smp]) / tot:5.1f}%  executed {int(r[i_ex]):>9d}  {r[i_src].strip()[:100]}")
        if ctx:
            print()


if __name__ == "__main__":
    a = sys.argv
    main(a[1], float(a[2]) if len(a) > 2 else 1.5, int(a[3]) if len(a) > 3 else 0)
