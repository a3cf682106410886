#!/usr/bin/env python
"""Per-function attribution of one ncu capture of a lane = cell kernel: executed warp instructions, active lanes per
instruction, stall samples.   python tools/ncu_lane_regions.py REP.ncu-rep [ncells]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys


def function_map(path):
    out, cur = {}, os.path.basename(path)
    try:
        lines = open(path, errors="replace").read().splitlines()
    except OSError:
        return {}
    pend = False
    for i, ln in enumerate(lines, 1):
        if re.search(r"PLM_LHD|PLM_LNOINL|__device__|__global__", ln) or pend:
            names = [x for x in re.findall(r"([A-Za-z_][A-Za-z_0-9]*)\s*\(", ln)
                     if x not in ("__launch_bounds__", "if", "for", "while", "defined", "asm", "operator")]
            if names:
                cur, pend = names[0], False
            else:
                pend = True
        out[i] = cur
    return out


def _i(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rep = sys.argv[1]
    ncells = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    fmaps = {}
    agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    cur_file, hdr, cur_fn = "?", None, "?"
    for r in csv.reader(io.StringIO(txt)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1]
            if cur_file not in fmaps:
                fmaps[cur_file] = function_map(cur_file)
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = {c: i for i, c in enumerate(r)}
            stall = [(c, i) for i, c in enumerate(r) if c.startswith("stall_") and "Not Issued" not in c]
            continue
        if hdr is None:
            continue
        if r[0]:  # a CUDA source line: the SASS rows that follow belong to it
            cur_fn = os.path.basename(cur_file).replace(".h", "") + "::" + fmaps[cur_file].get(int(r[0]), "?")
            continue
        a = agg[cur_fn]
        a[0] += _i(r[hdr["Instructions Executed"]])
        a[1] += _i(r[hdr["Thread Instructions Executed"]])
        a[2] += _i(r[hdr["# Samples"]])
        for c, i in stall:
            v = _i(r[i])
            if v:
                a[3][c] += v
    tot_i = sum(a[0] for a in agg.values())
    tot_s = sum(a[2] for a in agg.values())
    tot_t = sum(a[1] for a in agg.values())
    print("total warp instructions %d, thread instructions %d (%.1f lanes / instruction), samples %d" % (tot_i, tot_t, tot_t / max(tot_i, 1), tot_s))
    print("%-44s %7s %9s %6s %7s  %s" % ("function", "inst %", "inst/cell", "lanes", "smpl %", "top stalls"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        if a[0] == 0 and a[2] == 0:
            continue
        top = ", ".join("%s %.0f%%" % (c.replace("stall_", ""), 100.0 * v / max(a[2], 1)) for c, v in a[3].most_common(3))
        print("%-44s %6.2f%% %9.0f %6.1f %6.2f%%  %s" % (k[:44], 100.0 * a[0] / tot_i, a[0] / ncells if ncells else 0, a[1] / max(a[0], 1), 100.0 * a[2] / max(tot_s, 1), top))


if __name__ == "__main__":
    main()
