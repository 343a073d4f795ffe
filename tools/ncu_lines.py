#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export per CUDA source line.

usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > x.csv ; ncu_lines.py x.csv [top]
Prints, per (file, line): warp instructions executed, share, avg active threads, stall samples.
"""
import csv, sys, collections

def main(path, top=60):
    rows = list(csv.reader(open(path)))
    cur_file = None
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            iI = hdr.index("Instructions Executed")
            iT = hdr.index("Thread Instructions Executed")
            iS = hdr.index("# Samples")
            continue
        if hdr is None or r[0] == "":
            continue  # SASS rows
        try:
            line = int(r[0])
            inst = int(r[iI]); thr = int(r[iT]); smp = int(r[iS])
        except ValueError:
            continue
        key = (cur_file, line)
        a = agg.setdefault(key, [0, 0, 0, r[1].strip()])
        a[0] += inst; a[1] += thr; a[2] += smp
    tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
    print(f"total warp-instructions {tot:,}  samples {tots:,}")
    items = sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]
    for (f, l), (inst, thr, smp, src) in items:
        if inst == 0: continue
        print(f"{f}:{l:<5d} {inst/tot*100:6.2f}%  lanes {thr/max(inst,1):5.1f}  stall {smp/max(tots,1)*100:5.1f}%  {src[:90]}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60)
