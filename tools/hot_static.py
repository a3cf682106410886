#!/usr/bin/env python
"""Estimate the instruction-cache footprint of the per-step code of a NEW build from an OLD ncu capture.

    python tools/hot_static.py OLD_src_cs.csv OLD_GIT_REV NEW_kernels.cu.o [kernel-substring] [ncells steps_per_cell]

The integration kernel is bound by instruction fetch (profiles/r01b_*: SM instruction cache hit rate 71 %), and the GPU
box is minutes away, so code-size work is steered with this estimate: every SASS instruction of the new build carries
its inline chain of source lines (nvdisasm -gi); each source line gets the execution frequency (per BDF step) it had in
the old capture (lines are matched through a diff of the old and new sources); an instruction is as hot as the
coldest line of its chain. Prints the number of instructions / 128-byte lines at several hotness thresholds.
"""
import collections
import csv
import difflib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def old_line_freq(csv_path, per_step):
    freq = collections.defaultdict(float)
    hdr, cur = None, None
    for r in csv.reader(open(csv_path)):
        if not r:
            continue
        if r[0] == "File Path":
            f = os.path.basename(r[1])
            continue
        if r[0] == "Line No":
            hdr = {c: i for i, c in enumerate(r)}
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0] != "":
            cur = (f, int(r[0]))
            continue
        try:
            n = int(r[hdr["Instructions Executed"]])
        except ValueError:
            continue
        freq[cur] = max(freq[cur], n / per_step)
    return freq


def line_map(path, rev):
    """new line number -> old line number (None for new lines)."""
    rel = os.path.relpath(path, ROOT)
    old = subprocess.run(["git", "show", "%s:%s" % (rev, rel)], capture_output=True, text=True, cwd=ROOT).stdout.splitlines()
    new = open(path).read().splitlines()
    m = {}
    sm = difflib.SequenceMatcher(None, [x.strip() for x in old], [x.strip() for x in new], autojunk=False)
    for tag, i1, i2, j1, j2 in sm.get_opcodes():
        if tag == "equal":
            for k in range(i2 - i1):
                m[j1 + k + 1] = i1 + k + 1
        elif tag == "replace" and (i2 - i1) == (j2 - j1):
            for k in range(i2 - i1):
                m[j1 + k + 1] = i1 + k + 1
    return m


def main():
    old_csv, rev, obj = sys.argv[1:4]
    pat = sys.argv[4] if len(sys.argv) > 4 else "react_kernelINS_3LayILi21ELi84ELi2EEELi0"
    ncells = int(sys.argv[5]) if len(sys.argv) > 5 else 32768
    spc = float(sys.argv[6]) if len(sys.argv) > 6 else 746.0
    freq = old_line_freq(old_csv, ncells * spc)
    tmp = "/tmp/hot_static"
    os.makedirs(tmp, exist_ok=True)
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, x) for x in os.listdir(tmp) if x.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    maps = {}
    on, chain, pending = False, [], []
    hot = []  # (addr, hotness, fn-ish)
    for ln in dis.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            on = pat in ln
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            pending.append((m.group(1), int(m.group(2))))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(\S.*);", ln)
        if not m:
            continue
        if pending:
            chain, pending = pending, []
        h = None
        for f, l in chain:
            b = os.path.basename(f)
            if "pelelm_b200" in f:
                if f not in maps:
                    maps[f] = line_map(f, rev)
                ol = maps[f].get(l)
                v = freq.get((b, ol)) if ol else None
            else:
                v = freq.get((b, l))
            if v is not None:
                h = v if h is None else min(h, v)
        hot.append((int(m.group(1), 16), h, chain[-1] if chain else None))
    unknown = sum(1 for _, h, _ in hot if h is None)
    print("%d instructions in %s (%d on lines the old capture does not know: counted as hot if their neighbours are)"
          % (len(hot), pat, unknown))
    # unknown -> inherit from the previous known instruction
    last = 0.0
    vals = []
    for a, h, c in hot:
        if h is None:
            h = last
        last = h
        vals.append((a, h, c))
    for thr in (1.0, 0.5, 0.2, 0.05):
        sel = [(a, c) for a, h, c in vals if h >= thr]
        lines = len(set(a // 128 for a, _ in sel))
        print("executed >= %.2f per step: %5d instructions (%.1f KB), %4d lines of 128 B (%.1f KB)"
              % (thr, len(sel), len(sel) * 16 / 1024.0, lines, lines * 128 / 1024.0))
    by = collections.Counter()
    for a, h, c in vals:
        if h >= 0.2 and c:
            by["%s:%d" % (os.path.basename(c[0]), c[1] // 10 * 10)] += 1
    print("hot instructions by outermost source line (decade):")
    for k, v in sorted(by.items(), key=lambda kv: -kv[1])[:40]:
        print("  %-28s %4d" % (k, v))


if __name__ == "__main__":
    main()
