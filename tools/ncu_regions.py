#!/usr/bin/env python
"""Attribute the executed instructions / stall samples of one ncu capture to source functions and opcodes.

    python tools/ncu_regions.py gpurun_out/prof_react.ncu-rep [ncells] > profiles/rNN_react_kernel_by_region.txt

Reads `ncu -i REP --page source --csv --print-source cuda,sass` (needs -lineinfo and --import-source on), maps every
SASS instruction to the source line it was compiled from and that line to the enclosing function of our own sources
(library headers are reported under their file name), and prints per function and per opcode: executed warp
instructions (share, and per cell when ncells is given), share of stall samples, and the dominant stall reasons.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys


def function_map(path):
    """line number -> enclosing function (a definition starts on a line mentioning __device__ / __global__)."""
    out, cur, pend = {}, os.path.basename(path), False
    try:
        lines = open(path, errors="replace").read().splitlines()
    except OSError:
        return None
    for i, ln in enumerate(lines, 1):
        if "__device__" in ln or "__global__" in ln or pend:
            names = [x for x in re.findall(r"([A-Za-z_][A-Za-z_0-9]*)\s*\(", ln)
                     if x not in ("__launch_bounds__", "if", "for", "while", "defined", "asm")]
            if names:
                cur, pend = names[0], False
            else:
                pend = True
        out[i] = cur
    return out


def main():
    rep = sys.argv[1]
    ncells = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    fmaps = {}
    fn = collections.defaultdict(lambda: collections.Counter())
    op = collections.defaultdict(lambda: collections.Counter())
    lines_hot = collections.Counter()
    cur_file, cur_fn, hdr, stall_cols = "?", "?", None, []
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
            c_inst, c_samp, c_addr = hdr["Instructions Executed"], hdr["Warp Stall Sampling (All Samples)"], hdr["Address"]
            stall_cols = [(c, i) for c, i in hdr.items() if c.startswith("stall_") and "Not Issued" not in c]
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0] != "":  # a CUDA source line: sets the attribution of the SASS rows that follow
            fm = fmaps.get(cur_file)
            base = os.path.basename(cur_file)
            cur_fn = fm.get(int(r[0]), base) if (fm and "pelelm_b200" in cur_file) else base
            cur_line = "%s:%s" % (base, r[0])
            continue
        try:
            ninst, nsamp = int(r[c_inst]), int(r[c_samp])
        except ValueError:
            continue
        sass = r[hdr["Source"] + 2] if False else r[3]
        toks = sass.split()
        if toks and toks[0].startswith("@"):
            toks = toks[1:]
        opc = toks[0].split(".")[0] if toks else "?"
        for tgt, key in ((fn, cur_fn), (op, opc)):
            tgt[key]["inst"] += ninst
            tgt[key]["samp"] += nsamp
            tgt[key]["sass"] += 1
            for c, i in stall_cols:
                v = int(r[i] or 0)
                if v:
                    tgt[key][c] += v
        lines_hot[cur_line] += nsamp
    ti = sum(v["inst"] for v in fn.values()) or 1
    ts = sum(v["samp"] for v in fn.values()) or 1
    print("# %s: %d warp instructions executed, %d stall samples%s" % (
        os.path.basename(rep), ti, ts, (", %.0f instructions per cell" % (ti / ncells)) if ncells else ""))
    for title, tab in (("function", fn), ("opcode", op)):
        print("\n%-28s %7s %7s %6s%s  top stall reasons" % (title, "inst%", "samp%", "sass", "  inst/cell" if ncells else ""))
        for k, v in sorted(tab.items(), key=lambda kv: -kv[1]["samp"])[:40]:
            st = sorted(((c[6:], n) for c, n in v.items() if c.startswith("stall_")), key=lambda x: -x[1])[:3]
            print("%-28s %7.1f %7.1f %6d%s  %s" % (k[:28], 100.0 * v["inst"] / ti, 100.0 * v["samp"] / ts, v["sass"],
                                                   ("  %9.0f" % (v["inst"] / ncells)) if ncells else "",
                                                   " ".join("%s=%.0f%%" % (c, 100.0 * n / max(v["samp"], 1)) for c, n in st)))
    print("\nhottest source lines (stall samples %)")
    for k, n in lines_hot.most_common(40):
        print("%-32s %5.1f" % (k, 100.0 * n / ts))


if __name__ == "__main__":
    main()
