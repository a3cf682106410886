"""AMReX VisMF MultiFab files and the ReactEval state dump built from them.

What is mirrored (reference file:line):
  * amrex::VisMF::Write / Read, header version 1 (amrex/Src/Base/AMReX_VisMF.cpp; the `<name>_H` text header with the
    BoxArray, the FabOnDisk table and the per-FAB min/max tables, and `<name>_D_xxxxx` holding `FAB (<real descriptor>)
    (<box>) ncomp\\n` + raw reals per FAB, amrex/Src/Base/AMReX_FabConv.cpp / AMReX_FArrayBox.cpp)
  * the chkfile a PeleLM run dumps for PelePhysics' ReactEval and how ReactEval loads it
    (PelePhysics/Testing/Exec/ReactEval/main.cpp:206-300 `Header`; :358-405 `Level_n/OldState` with
    D + N + 4 components (velocity, rho, rhoY_k, rhoh, T, rhoRT) and `Level_n/ChemForcing` with N + 1; MKS -> CGS on load)

`read_reacteval_chk` returns, per box, exactly the arrays `Reactor.react` takes (FAB layout, CGS), so a real PeleLM state
can be replayed through the CUDA path, the oracle port and the reference alike. Host-side I/O only: numpy.
"""
from __future__ import annotations

import os
import re

import numpy as np

_DBL_LE = "(8, (64 11 52 0 1 12 0 1023)),(8, (8 7 6 5 4 3 2 1))"
_DBL_BE = "(8, (64 11 52 0 1 12 0 1023)),(8, (1 2 3 4 5 6 7 8))"
_FLT_LE = "(4, (32 8 23 0 1 9 0 127)),(4, (4 3 2 1))"
_FLT_BE = "(4, (32 8 23 0 1 9 0 127)),(4, (1 2 3 4))"
_DTYPES = {_DBL_LE: "<f8", _DBL_BE: ">f8", _FLT_LE: "<f4", _FLT_BE: ">f4"}

_BOX_RE = re.compile(r"\(\(\s*(-?\d+)\s*,\s*(-?\d+)\s*(?:,\s*(-?\d+)\s*)?\)\s*\(\s*(-?\d+)\s*,\s*(-?\d+)\s*(?:,\s*(-?\d+)\s*)?\)"
                     r"\s*\(\s*(-?\d+)\s*,\s*(-?\d+)\s*(?:,\s*(-?\d+)\s*)?\)\)")


def _parse_box(text):
    """'((lo) (hi) (type))' -> (lo, hi) as 3-tuples (a 2-D box gets k = 0)."""
    m = _BOX_RE.search(text)
    if not m:
        raise ValueError("not a Box: %r" % text)
    g = m.groups()
    lo = (int(g[0]), int(g[1]), int(g[2]) if g[2] is not None else 0)
    hi = (int(g[3]), int(g[4]), int(g[5]) if g[5] is not None else 0)
    return lo, hi


def _fmt_box(lo, hi, dim=3):
    f = lambda v: "(" + ",".join(str(int(x)) for x in v[:dim]) + ")"
    return "(%s %s %s)" % (f(lo), f(hi), f((0, 0, 0)))


def _read_boxarray(lines, i):
    """BoxArray::readFrom: '(n hash' then n boxes then ')'. Returns (boxes, next line index)."""
    m = re.match(r"\s*\((\d+)\s+\d+", lines[i])
    if not m:
        raise ValueError("not a BoxArray header: %r" % lines[i])
    n = int(m.group(1))
    boxes = [_parse_box(lines[i + 1 + b]) for b in range(n)]
    i += 1 + n
    if lines[i].strip() != ")":
        raise ValueError("BoxArray not closed")
    return boxes, i + 1


def read_multifab(prefix):
    """VisMF::Read(mf, prefix): returns dict(boxes=[(lo, hi)], ncomp, ngrow, fabs=[array (ncomp, nz, ny, nx)]).
    A FAB written with ghost cells comes back with them (its own box is in `fab_boxes`)."""
    with open(prefix + "_H") as f:
        lines = f.read().split("\n")
    version = int(lines[0].split()[0])
    if version not in (1, 2, 3, 4):
        raise ValueError("unknown VisMF header version %d" % version)
    ncomp = int(lines[2].split()[0])
    g = lines[3].strip()
    ngrow = int(re.findall(r"-?\d+", g)[0])
    boxes, i = _read_boxarray(lines, 4)
    nfod = int(lines[i].split()[0])
    fod = []
    for b in range(nfod):
        _, name, off = lines[i + 1 + b].split()
        fod.append((name, int(off)))
    d = os.path.dirname(prefix)
    fabs, fab_boxes = [], []
    for name, off in fod:
        with open(os.path.join(d, name), "rb") as f:
            f.seek(off)
            hdr = f.readline().decode("ascii")
            m = re.match(r"FAB \((.*)\)\((\(.*\))\) (\d+)\s*$", hdr)
            if not m:
                raise ValueError("bad FAB header %r" % hdr)
            desc, nc = m.group(1), int(m.group(3))
            if desc not in _DTYPES:
                raise ValueError("unsupported real descriptor %r" % desc)
            lo, hi = _parse_box("(" + m.group(2) + ")")
            shape = (nc, hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1)
            cnt = int(np.prod(shape))
            a = np.frombuffer(f.read(cnt * np.dtype(_DTYPES[desc]).itemsize), dtype=_DTYPES[desc], count=cnt)
            fabs.append(a.astype(np.float64).reshape(shape))
            fab_boxes.append((lo, hi))
    return {"boxes": boxes, "ncomp": ncomp, "ngrow": ngrow, "fabs": fabs, "fab_boxes": fab_boxes}


def write_multifab(prefix, boxes, fabs, dim=3):
    """VisMF::Write(mf, prefix) with one data file (header version 1, native little-endian doubles, no ghost cells).
    fabs[b]: array (ncomp, nz, ny, nx) for boxes[b] = (lo, hi)."""
    ncomp = fabs[0].shape[0]
    d, base = os.path.dirname(prefix), os.path.basename(prefix)
    os.makedirs(d, exist_ok=True)
    dname = base + "_D_00000"
    offs = []
    with open(os.path.join(d, dname), "wb") as f:
        for (lo, hi), a in zip(boxes, fabs):
            shape = (ncomp, hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1)
            if tuple(a.shape) != shape:
                raise ValueError("FAB shape %r does not match its box %r" % (a.shape, shape))
            offs.append(f.tell())
            f.write(("FAB (%s)%s %d\n" % (_DBL_LE, _fmt_box(lo, hi, dim), ncomp)).encode("ascii"))
            f.write(np.ascontiguousarray(a, dtype="<f8").tobytes())
    with open(prefix + "_H", "w") as f:
        f.write("1\n1\n%d\n0\n" % ncomp)
        f.write("(%d 0\n" % len(boxes))
        for lo, hi in boxes:
            f.write(_fmt_box(lo, hi, dim) + "\n")
        f.write(")\n%d\n" % len(boxes))
        for o in offs:
            f.write("FabOnDisk: %s %d\n" % (dname, o))
        f.write("\n")
        for red in (np.min, np.max):
            f.write("%d,%d\n" % (len(boxes), ncomp))
            for a in fabs:
                f.write("".join("%.16e," % red(a[n]) for n in range(ncomp)) + "\n")
            f.write("\n")


# ---------------------------------------------------------------------------------------------------------------
# ReactEval chkfile
# ---------------------------------------------------------------------------------------------------------------
def read_reacteval_chk(chkfile, ns, dim=3):
    """Load a PeleLM -> ReactEval state dump as ReactEval does (main.cpp:206-300, :358-405).

    Returns dict(dt, finest_level, prob_lo, prob_hi, domain, levels=[[box dict, ...], ...]); every box dict holds
    lo, hi and the CGS arrays of the react() call in FAB layout: rY (ns, nz, ny, nx), rYsrc, T, rE, rEsrc."""
    with open(os.path.join(chkfile, "Header")) as f:
        lines = f.read().split("\n")
    finest = int(lines[1].split()[0])
    dt = float(lines[2].split()[0])
    prob_lo = [float(x) for x in lines[3].split()]
    prob_hi = [float(x) for x in lines[4].split()]
    domain = _parse_box(lines[5])
    i = 6
    grids = []
    for lev in range(finest + 1):
        while not lines[i].strip().startswith("("):
            i += 1
        ba, i = _read_boxarray(lines, i)
        grids.append(ba)
    levels = []
    for lev in range(finest + 1):
        st = read_multifab(os.path.join(chkfile, "Level_%d" % lev, "OldState"))
        fr = read_multifab(os.path.join(chkfile, "Level_%d" % lev, "ChemForcing"))
        if st["ncomp"] != ns + dim + 4 or fr["ncomp"] != ns + 1:
            raise ValueError("OldState/ChemForcing have %d/%d components, expected %d/%d" %
                             (st["ncomp"], fr["ncomp"], ns + dim + 4, ns + 1))
        if st["boxes"] != grids[lev]:
            raise ValueError("OldState BoxArray differs from the Header's at level %d" % lev)
        out = []
        for (lo, hi), S, F in zip(st["boxes"], st["fabs"], fr["fabs"]):
            out.append({"lo": lo, "hi": hi,
                        "rY": S[dim + 1:dim + 1 + ns] * 1.0e-3,      # main.cpp:391-392 (MKS -> CGS)
                        "rYsrc": F[:ns] * 1.0e-3,                    # :393
                        "T": S[dim + ns + 2].copy(),                 # :395
                        "rE": S[dim + ns + 1] * 10.0,                # :396-397
                        "rEsrc": F[ns] * 10.0})                      # :398
        levels.append(out)
    return {"dt": dt, "finest_level": finest, "prob_lo": prob_lo, "prob_hi": prob_hi, "domain": domain, "levels": levels}


def write_reacteval_chk(chkfile, dt, prob_lo, prob_hi, domain, levels, ns, dim=3):
    """The dump ReactEval's `initFile` expects, from MKS PeleLM-style data: levels[lev] = list of dicts with lo, hi,
    state (dim + ns + 4, nz, ny, nx) and forcing (ns + 1, nz, ny, nx)."""
    os.makedirs(chkfile, exist_ok=True)
    with open(os.path.join(chkfile, "Header"), "w") as f:
        f.write("ReactEval state dump (pelelm_b200.vismf)\n%d\n%.17g\n" % (len(levels) - 1, dt))
        f.write(" ".join("%.17g" % x for x in prob_lo[:dim]) + "\n")
        f.write(" ".join("%.17g" % x for x in prob_hi[:dim]) + "\n")
        f.write(_fmt_box(domain[0], domain[1], dim) + "\n")
        for lev in levels:
            f.write("(%d 0\n" % len(lev))
            for b in lev:
                f.write(_fmt_box(b["lo"], b["hi"], dim) + "\n")
            f.write(")\n")
    for l, lev in enumerate(levels):
        boxes = [(tuple(b["lo"]), tuple(b["hi"])) for b in lev]
        for b in lev:
            assert b["state"].shape[0] == dim + ns + 4 and b["forcing"].shape[0] == ns + 1
        write_multifab(os.path.join(chkfile, "Level_%d" % l, "OldState"), boxes, [b["state"] for b in lev], dim)
        write_multifab(os.path.join(chkfile, "Level_%d" % l, "ChemForcing"), boxes, [b["forcing"] for b in lev], dim)
