"""Emit straight-line, lane = cell mechanism kernels (``pelelm_b200/csrc/gen/lane_<mech>.h``) from ``data/<mech>.json``.

One THREAD evaluates one cell: every species / reaction index is a compile-time literal, every parameter is an
operand read from the constant bank, nothing is decoded or shuffled at run time.  The statements follow the
reference's generated mechanism kernels reaction by reaction and in the same order (drm19 line numbers;
SURVEY.md section 8a lists the other mechanisms):

  mix_hc / species_h   CKHBMS :1712, CKCPBS :1605, CKHMS :2770 (speciesEnthalpy :1207, cp_R :342)
  production_rate      productionRate :3650-4902 (gibbs :531)
  ajacobian            aJacobian :9696-14742 (consP = 1) + the mass scaling of cJac
                       (PelePhysics/Reactions/ReactorCvodeJacobian.cpp:101-148)

The generated header is plain C++ behind three macros (PLM_LHD, PLM_LCONST, the lane_* math functions of
csrc/lane_common.h), so the same text compiles for sm_100a (the product) and with g++ for the host-side model the
tests use to check the integrator logic without a GPU (tests/lane_model.cpp).
"""
from __future__ import annotations

import os
import struct

from .emit_tables import HERE, ROOT, load


def _is_imm(v: float) -> bool:
    """FP64 literals whose low word is zero are encodable as instruction immediates; the rest go to the constant bank."""
    return struct.pack(">d", float(v))[4:] == b"\0\0\0\0"


class Pool:
    def __init__(self):
        self.vals: list[float] = []
        self.index: dict[bytes, int] = {}

    def k(self, v: float) -> str:
        v = float(v)
        if _is_imm(v):
            return repr(v)
        key = struct.pack(">d", v)
        if key not in self.index:
            self.index[key] = len(self.vals)
            self.vals.append(v)
        return "KK[%d]" % self.index[key]


class Emit:
    def __init__(self, pool: Pool, indent: int = 4):
        self.pool = pool
        self.lines: list[str] = []
        self.n = 0
        self.ind = " " * indent

    def k(self, v):
        return self.pool.k(v)

    def let(self, expr: str, name: str | None = None) -> str:
        if name is None:
            name = "v%d" % self.n
            self.n += 1
        self.lines.append("%sconst double %s = %s;" % (self.ind, name, expr))
        return name

    def stmt(self, s: str):
        self.lines.append(self.ind + s)

    def text(self) -> str:
        return "\n".join(self.lines) + "\n"


def _net(r):
    d: dict[int, float] = {}
    for i, c in r["reactants"]:
        d[i] = d.get(i, 0.0) - c
    for i, c in r["products"]:
        d[i] = d.get(i, 0.0) + c
    return sorted((i, c) for i, c in d.items() if c != 0.0)


def _phase(e: Emit, side, sc="sc"):
    """phi = prod c^nu over the side's slots (oracle/port/chem.c phase())."""
    phi = None
    for i, nu in side:
        c = "%s[%d]" % (sc, i)
        f = c if nu == 1.0 else ("(%s * %s)" % (c, c) if nu == 2.0 else "lane_pow(%s, %s)" % (c, e.k(nu)))
        phi = f if phi is None else "%s * %s" % (phi, f)
    return phi if phi is not None else "1.0"


def _dphase(e: Emit, side, j, sc="sc"):
    """d(phi)/d(sc[j]) or None if j is not on the side (chem.c dphase(): running product started at 1.0)."""
    if not any(i == j for i, _ in side):
        return None
    fac = []
    for i, nu in side:
        c = "%s[%d]" % (sc, i)
        if i == j:
            if nu == 1.0:
                pass  # d *= 1.0
            elif nu == 2.0:
                fac.append("(2.0 * %s)" % c)
            else:
                fac.append("(%s * lane_pow(%s, %s))" % (e.k(nu), c, e.k(nu - 1.0)))
        else:
            fac.append(c if nu == 1.0 else "lane_pow(%s, %s)" % (c, e.k(nu)))
    return " * ".join(fac) if fac else "1.0"


def _poly(e: Emit, c, powers):
    """c0*p0 + c1*p1 + ... left to right, powers[i] None = bare constant."""
    terms = []
    for ci, p in zip(c, powers):
        terms.append(e.k(ci) if p is None else "%s * %s" % (e.k(ci), p))
    return " + ".join(terms)


def _thermo_groups(th, ns):
    groups: dict[float, list[int]] = {}
    for k in range(ns):
        groups.setdefault(float(th["tmid"][k]), []).append(k)
    return sorted(groups.items())


def emit_species_fn(mech, pool, name, body_of):
    """One function whose body evaluates a per-species expression on either temperature range (branch per T_mid group,
    as the generated reference does)."""
    ns = mech["n_species"]
    th = mech["thermo"]
    out = []
    for tmid, ks in _thermo_groups(th, ns):
        out.append("    if (T < %s) {" % repr(tmid))
        for k in ks:
            out.append("      " + body_of(k, 0))
        out.append("    } else {")
        for k in ks:
            out.append("      " + body_of(k, 1))
        out.append("    }")
    return "\n".join(out) + "\n"


def emit_thermo(mech, pool) -> str:
    ns = mech["n_species"]
    th = mech["thermo"]
    e = Emit(pool)
    from decimal import Decimal
    a0m1 = [[float(Decimal(repr(float(rng[0]))) - 1) for rng in sp] for sp in th["cp"]]
    imw = [1.0 / float(w) for w in mech["mw"]]
    s = []
    # ---- mixture enthalpy / heat capacity (CKHBMS + CKCPBS, or CKUBMS + CKCVBS), Y = mass fractions ----
    s.append("  // CKHBMS :1712 / CKCPBS :1605 (EINT: CKUBMS, CKCVBS): he = sum Y_k h_k/RT / W_k * RT, c = sum c_k/R Y_k / W_k * R\n")
    s.append("  template <int EINT>\n  PLM_LHD static void mix_hc(double T, const double (&Y)[NS], double& he, double& c)\n  {\n")
    s.append("    const double T2 = T * T, T3 = T * T * T, T4 = T * T * T * T, invT = 1 / T;\n")
    s.append("    double hs = 0.0, cs = 0.0;\n")

    def body(k, rg):
        ch = th["h"][k][rg]
        cc = th["cp"][k][rg]
        c0 = "(EINT ? %s : %s)" % (e.k(a0m1[k][rg]), e.k(ch[0]))
        c0c = "(EINT ? %s : %s)" % (e.k(a0m1[k][rg]), e.k(cc[0]))
        h = "%s + %s" % (c0, _poly(e, ch[1:], ["T", "T2", "T3", "T4", "invT"]))
        cp = "%s + %s" % (c0c, _poly(e, cc[1:], ["T", "T2", "T3", "T4"]))
        return "hs += Y[%d] * (%s) * %s; cs += (%s) * Y[%d] * %s;" % (k, h, e.k(imw[k]), cp, k, e.k(imw[k]))

    # the reference sums in species order; the T_mid groups preserve it only if every group is contiguous -- keep the
    # order exact by emitting species one at a time when groups interleave
    groups = _thermo_groups(th, ns)
    contiguous = all(ks == list(range(ks[0], ks[-1] + 1)) for _, ks in groups) and \
        [k for _, ks in sorted(groups, key=lambda g: g[1][0]) for k in ks] == list(range(ns))
    if contiguous:
        for tmid, ks in sorted(groups, key=lambda g: g[1][0]):
            s.append("    if (T < %s) {\n" % repr(tmid))
            for k in ks:
                s.append("      %s\n" % body(k, 0))
            s.append("    } else {\n")
            for k in ks:
                s.append("      %s\n" % body(k, 1))
            s.append("    }\n")
    else:
        for k in range(ns):
            s.append("    if (T < %s) { %s } else { %s }\n" % (repr(float(th["tmid"][k])), body(k, 0), body(k, 1)))
    s.append("    he = hs * (%s * T);\n    c = cs * %s;\n  }\n" % (e.k(8.31446261815324e+07), e.k(8.31446261815324e+07)))

    # ---- species enthalpies per mass (CKHMS :2770 / CKUMS) ----
    s.append("  // CKHMS :2770 (EINT: CKUMS): h_k = h_k/RT * RT / W_k  [erg/g]\n")
    s.append("  template <int EINT>\n  PLM_LHD static void species_h(double T, double (&hms)[NS])\n  {\n")
    s.append("    const double T2 = T * T, T3 = T * T * T, T4 = T * T * T * T, invT = 1 / T;\n")
    s.append("    const double RT = %s * T;\n" % e.k(8.31446261815324e+07))

    def body_h(k, rg):
        ch = th["h"][k][rg]
        c0 = "(EINT ? %s : %s)" % (e.k(a0m1[k][rg]), e.k(ch[0]))
        return "hms[%d] = (%s + %s) * (RT * %s);" % (k, c0, _poly(e, ch[1:], ["T", "T2", "T3", "T4", "invT"]), e.k(imw[k]))

    s.append(emit_species_fn(mech, pool, "species_h", body_h))
    s.append("  }\n")

    # ---- g/RT (gibbs :531), h/RT, cp/R, dcp/R dT for the rates and the Jacobian ----
    def fn(name, key, powers, comment):
        s.append("  // %s\n  PLM_LHD static void %s(double T, double logT, double (&o)[NS])\n  {\n" % (comment, name))
        s.append("    const double T2 = T * T, T3 = T * T * T, T4 = T * T * T * T, invT = 1 / T;\n    (void)T2; (void)T3; (void)T4; (void)invT; (void)logT;\n")
        s.append(emit_species_fn(mech, pool, name, lambda k, rg: "o[%d] = %s;" % (k, _poly(e, th[key][k][rg], powers))))
        s.append("  }\n")

    fn("g_RT", "g", ["invT", None, "logT", "T", "T2", "T3", "T4"], "gibbs :531")
    fn("h_RT", "h", [None, "T", "T2", "T3", "T4", "invT"], "speciesEnthalpy :1207")
    fn("cp_R", "cp", [None, "T", "T2", "T3", "T4"], "cp_R :342")
    fn("dcp_R", "dcp", [None, "T", "T2", "T3"], "dcvpRdT :5114")
    return "".join(s)


def _kf(e: Emit, r, always_exp=False):
    """forward rate constant; productionRate skips the exponential when b and Ea are both zero."""
    A, b, Ea = e.k(r["A"]), r["b"], r["Ea_R"]
    if b != 0.0 and Ea != 0.0:
        return "%s * lane_exp(%s * logT - %s * invT)" % (A, e.k(b), e.k(Ea))
    if b != 0.0:
        return "%s * lane_exp(%s * logT)" % (A, e.k(b))
    if Ea != 0.0:
        return "%s * lane_exp(-%s * invT)" % (A, e.k(Ea))
    return A  # exp(0) == 1 exactly


def _refc_pow(e: Emit, dn, inverse: bool):
    """refC^(-dn) (production rate, multiplies k_r) or refC^dn (Jacobian, multiplies Kc)."""
    if not inverse:
        if dn == -1.0:
            return "refC"
        if dn == 1.0:
            return "refCinv"
        return ("lane_pow(refC, %s)" % e.k(-dn)) if dn < 0 else ("lane_pow(refCinv, %s)" % e.k(dn))
    if dn == -1.0:
        return "refCinv"
    if dn == 1.0:
        return "refC"
    return "lane_pow(refC, %s)" % e.k(dn)


def emit_production_rate(mech, pool) -> str:
    ns = mech["n_species"]
    R = mech["reactions"]
    e = Emit(pool)
    e.stmt("const double logT = lane_log(T);")
    e.stmt("const double invT = lane_rcp(T);")
    e.stmt("const double refC = %s * invT;" % e.k(101325 / 8.31446))
    e.stmt("const double refCinv = lane_rcp(refC);")
    e.stmt("(void)refC; (void)refCinv;")
    e.stmt("double g[NS];")
    e.stmt("g_RT(T, logT, g);")
    e.stmt("double mixture = 0.0;")
    for k in range(ns):
        e.stmt("mixture += sc[%d];" % k)
    for k in range(ns):
        e.stmt("wdot[%d] = 0.0;" % k)
    for ri, r in enumerate(R):
        e.stmt("{  // reaction %d (input file #%d): %s" % (ri, r["orig"], r.get("equation", "")))
        e.ind = "      "
        typ = r["type"]
        kf = e.let(_kf(e, r), "k_f")
        corr = None
        if typ != 0:
            terms = ["mixture"]
            for i, ef in sorted(r["eff"]):
                em1 = ef - 1.0
                terms.append("sc[%d]" % i if em1 == 1.0 else "%s * sc[%d]" % (e.k(em1), i))
            corr = e.let(" + ".join(terms), "alpha")
        if typ >= 2:
            redP = e.let("%s / k_f * %s * lane_exp(%s * logT - %s * invT)" % (corr, e.k(r["A0"]), e.k(r["b0"]), e.k(r["Ea0_R"])), "redP")
            F = e.let("redP / (1.0 + redP)", "F")
            if typ == 3:
                tr = r["troe"]
                np4 = bool(r.get("troe_has_T2"))
                logPred = e.let("lane_log10(redP)", "logPred")
                fc = "0.0"
                if abs(tr[1]) > 1.0e-100 and (1.0 - tr[0]) != 0:
                    fc = "%s * lane_exp(-T * %s)" % (e.k(1.0 - tr[0]), e.k(1 / tr[1]))
                if abs(tr[2]) > 1.0e-100 and tr[0] != 0:
                    fc = "%s + %s * lane_exp(-T * %s)" % (fc, e.k(tr[0]), e.k(1 / tr[2]))
                if np4:
                    fc = "%s + lane_exp(-%s * invT)" % (fc, e.k(tr[3]))
                e.let(fc, "Fcent")
                e.let("lane_log10(Fcent)", "logFcent")
                e.let("-%s - %s * logFcent" % (e.k(0.4), e.k(0.67)), "troe_c")
                e.let("0.75 - %s * logFcent" % e.k(1.27), "troe_n")
                e.let("(troe_c + logPred) / (troe_n - %s * (troe_c + logPred))" % e.k(0.14), "troe")
                e.let("lane_exp10(logFcent / (1.0 + troe * troe))", "F_troe")
                corr = e.let("F * F_troe", "Corr")
            else:
                corr = F
        phi_f = _phase(e, sorted(r["reactants"]))
        qf = e.let(("%s * k_f * %s" % (corr, phi_f)) if typ != 0 else ("k_f * %s" % phi_f), "qf")
        if r["reversible"]:
            net = _net(r)
            dG = None
            for i, nv in net:
                c = -nv
                term = "g[%d]" % i if abs(c) == 1.0 else "%s * g[%d]" % (e.k(abs(c)), i)
                if dG is None:
                    dG = term if c > 0 else "-%s" % term
                else:
                    dG = "%s + %s" % (dG, term) if c > 0 else "%s - %s" % (dG, term)
            phi_r = _phase(e, sorted(r["products"]))
            kr = ("%s * k_f * lane_exp(-(%s))" % (corr, dG)) if typ != 0 else ("k_f * lane_exp(-(%s))" % dG)
            if r["dn"] != 0.0:
                kr = "%s * %s" % (kr, _refc_pow(e, r["dn"], False))
            e.let(kr, "k_r")
            e.let("qf - k_r * %s" % phi_r, "qdot")
        else:
            e.let("qf", "qdot")
        for i, c in _net(r):
            if c == 1.0:
                e.stmt("wdot[%d] += qdot;" % i)
            elif c == -1.0:
                e.stmt("wdot[%d] -= qdot;" % i)
            elif c > 0:
                e.stmt("wdot[%d] += %s * qdot;" % (i, e.k(c)))
            else:
                e.stmt("wdot[%d] -= %s * qdot;" % (i, e.k(-c)))
        e.ind = "    "
        e.stmt("}")
    return e.text()


def emit_ajacobian(mech, pool) -> str:
    """aJacobian with consP = 1 followed by cJac's mass scaling, for ONE cell.  The species block is accumulated in
    place in the caller's matrix (J(row, col) is an lvalue accessor: column-major like the SUNDenseMatrix); the T column is
    accumulated in registers.  Reaction order and the order of the updates of every entry are the reference's."""
    ns = mech["n_species"]
    R = mech["reactions"]
    mw = [float(w) for w in mech["mw"]]
    e = Emit(pool)
    e.stmt("const double logT = lane_log(T);")
    e.stmt("const double invT = lane_rcp(T);")
    e.stmt("const double invT2 = invT * invT;")
    e.stmt("const double refC = lane_div(%s, T);" % e.k(101325 / 8.31446))
    e.stmt("const double refCinv = lane_rcp(refC);")
    e.stmt("(void)refC; (void)refCinv;")
    e.stmt("const double ln10 = %s;" % e.k(2.302585092994046))
    e.stmt("const double log10e = %s;" % e.k(1.0 / 2.302585092994046))
    e.stmt("(void)ln10; (void)log10e;")
    e.stmt("double g[NS], h[NS];")
    e.stmt("g_RT(T, logT, g);")
    e.stmt("h_RT(T, logT, h);")
    e.stmt("double mixture = 0.0;")
    for k in range(ns):
        e.stmt("mixture += sc[%d];" % k)
    e.stmt("double wdot[NS], jT[NS];  // d(wdot_k)/dT accumulates in registers")
    for k in range(ns):
        e.stmt("wdot[%d] = 0.0; jT[%d] = 0.0;" % (k, k))
    e.stmt("for (int c = 0; c < NS; ++c)")
    e.stmt("  for (int r = 0; r < NS; ++r) J(r, c) = 0.0;")
    for ri, r in enumerate(R):
        typ = r["type"]
        reac, prod, net = sorted(r["reactants"]), sorted(r["products"]), _net(r)
        eff = sorted(r["eff"])
        rev = r["reversible"]
        e.stmt("{  // reaction %d (input file #%d): %s" % (ri, r["orig"], r.get("equation", "")))
        e.ind = "      "
        if typ != 0:
            terms = ["mixture"] + ["%s * sc[%d]" % (e.k(ef - 1), i) for i, ef in eff]
            e.let(" + ".join(terms), "alpha")
        e.let(_phase(e, reac), "phi_f")
        b, Ea = r["b"], r["Ea_R"]
        # always exp(b*logT - Ea*invT) in the reference; 0*x terms drop out exactly
        if b != 0.0 and Ea != 0.0:
            kf = "%s * lane_exp(%s * logT - %s * invT)" % (e.k(r["A"]), e.k(b), e.k(Ea))
        elif b != 0.0:
            kf = "%s * lane_exp(%s * logT)" % (e.k(r["A"]), e.k(b))
        elif Ea != 0.0:
            kf = "%s * lane_exp(-(%s * invT))" % (e.k(r["A"]), e.k(Ea))
        else:
            kf = e.k(r["A"])
        e.stmt("double k_f = %s;" % kf)
        e.let("%s * invT + %s * invT2" % (e.k(b), e.k(Ea)), "dlnkfdT")
        if typ >= 2:
            e.let("%s * lane_exp(%s * logT - %s * invT)" % (e.k(r["A0"]), e.k(r["b0"]), e.k(r["Ea0_R"])), "k_0")
            e.let("alpha / k_f * k_0", "Pr")
            e.let("Pr / (1.0 + Pr)", "fPr")
            e.let("%s * invT + %s * invT2" % (e.k(r["b0"]), e.k(r["Ea0_R"])), "dlnk0dT")
            e.let("log10e * (dlnk0dT - dlnkfdT)", "dlogPrdT")
            e.let("dlogPrdT / (1.0 + Pr)", "dlogfPrdT")
            if typ == 3:
                tr = r["troe"]
                np4 = bool(r.get("troe_has_T2"))
                e.let("lane_log10(Pr)", "logPr")
                e.let(("%s * lane_exp(-T / %s)" % (e.k(1.0 - tr[0]), e.k(tr[1]))) if abs(tr[1]) > 1e-100 else "0.0", "Fcent1")
                e.let(("%s * lane_exp(-T / %s)" % (e.k(tr[0]), e.k(tr[2]))) if abs(tr[2]) > 1e-100 else "0.0", "Fcent2")
                e.let(("lane_exp(-%s * invT)" % e.k(tr[3])) if np4 else "0.0", "Fcent3")
                e.let("Fcent1 + Fcent2 + Fcent3", "Fcent")
                e.let("lane_log10(Fcent)", "logFcent")
                e.let("-%s - %s * logFcent" % (e.k(0.4), e.k(0.67)), "troe_c")
                e.let("0.75 - %s * logFcent" % e.k(1.27), "troe_n")
                e.let("1.0 / (troe_n - %s * (troe_c + logPr))" % e.k(0.14), "troePr_den")
                e.let("(troe_c + logPr) * troePr_den", "troePr")
                e.let("1.0 / (1.0 + troePr * troePr)", "troe")
                e.let("lane_exp10(logFcent * troe)", "F")
                dl = "-Fcent1 / %s - Fcent2 / %s" % (e.k(tr[1]), e.k(tr[2]))
                if np4:
                    dl = "%s + Fcent3 * %s * invT2" % (dl, e.k(tr[3]))
                e.let("log10e / Fcent * (%s)" % dl, "dlogFcentdT")
                e.let("2.0 * logFcent * troe * troe * troePr * troePr_den", "dlogFdcn_fac")
                e.let("-troe_n * dlogFdcn_fac * troePr_den", "dlogFdc")
                e.let("dlogFdcn_fac * troePr", "dlogFdn")
                e.let("dlogFdc", "dlogFdlogPr")
                e.let("dlogFcentdT * (troe - %s * dlogFdc - %s * dlogFdn) + dlogFdlogPr * dlogPrdT" % (e.k(0.67), e.k(1.27)), "dlogFdT")
            else:
                e.let("1.0", "F")
                e.let("0.0", "dlogFdlogPr")
                e.let("0.0", "dlogFdT")
        if rev:
            e.let(_phase(e, prod), "phi_r")
            dG = " + ".join("%s * g[%d]" % (e.k(-nv), i) for i, nv in net)
            e.stmt("double Kc = lane_exp(%s);" % dG)
            dn = r["dn"]
            if dn != 0.0:
                e.stmt("Kc = %s * Kc;" % _refc_pow(e, dn, True))
            e.stmt("double k_r = k_f / Kc;")
            hr = " + ".join("%s * h[%d]" % (e.k(nu), i) for i, nu in reac)
            hp = " + ".join("%s * h[%d]" % (e.k(nu), i) for i, nu in prod)
            e.let("invT * (-(%s) + (%s) + (%s))" % (hr, hp, e.k(-dn)), "dlnKcdT")
            e.let("(dlnkfdT - dlnKcdT) * k_r", "dkrdT")
        else:
            e.let("0.0", "phi_r")
            e.stmt("double k_r = 0.0;")
            e.let("0.0", "dkrdT")
        if typ == 0:
            e.let("k_f * phi_f - k_r * phi_r", "q")
            e.let("(dlnkfdT * k_f * phi_f - dkrdT * phi_r)", "dqdT")
        elif typ == 1:
            e.let("k_f * phi_f - k_r * phi_r", "q_nocor")
            e.let("alpha * q_nocor", "q")
            e.let("alpha * (dlnkfdT * k_f * phi_f - dkrdT * phi_r)", "dqdT")
            e.stmt("k_f *= alpha; k_r *= alpha;")
            e.let("q_nocor", "dcdc_fac")
        else:
            e.let("k_f * phi_f - k_r * phi_r", "q_nocor")
            e.let("fPr * F", "Corr")
            e.let("Corr * q_nocor", "q")
            e.let("ln10 * (dlogfPrdT + dlogFdT)", "dlnCorrdT")
            e.let("Corr * (dlnkfdT * k_f * phi_f - dkrdT * phi_r) + dlnCorrdT * q", "dqdT")
            e.stmt("k_f *= Corr; k_r *= Corr;")
            e.let("q / alpha * (1.0 / (Pr + 1.0) + dlogFdlogPr)", "dcdc_fac")
        e.stmt("(void)k_r;")
        for i, nv in net:
            e.stmt("wdot[%d] += %s * q;" % (i, e.k(nv)))
        effd = dict(eff)
        for j in range(ns):
            terms = []
            if typ != 0 and j in effd:
                terms.append("%s * dcdc_fac" % e.k(effd[j] - 1))
            df = _dphase(e, reac, j)
            if df is not None:
                terms.append(("+ k_f * %s" % df) if df != "1.0" else "+ k_f * 1.0")
            dr = _dphase(e, prod, j) if rev else None
            if dr is not None:
                terms.append(("- k_r * %s" % dr) if dr != "1.0" else "- k_r * 1.0")
            if not terms:
                continue
            expr = " ".join(terms)
            if expr.startswith("+ "):
                expr = "0.0 " + expr
            elif expr.startswith("- "):
                expr = "0.0 " + expr
            d = e.let(expr)
            for i, nv in net:
                e.stmt("J(%d, %d) += %s * %s;" % (i, j, e.k(nv), d))
        for i, nv in net:
            e.stmt("jT[%d] += %s * dqdT;" % (i, e.k(nv)))
        e.ind = "    "
        e.stmt("}")
    # ---- closure :14705-14741 (consP) fused with the scaling of cJac (ReactorCvodeJacobian.cpp:101-148) ----
    # The (T,T) entry the closure computes is never copied out by cJac, so dcvpRdT / dcmixdT / dehmixdT are not formed.
    e.stmt("double c_R[NS];")
    e.stmt("cp_R(T, logT, c_R);")
    e.stmt("double cmix = 0.0, ehmix = 0.0;")
    for k in range(ns):
        e.stmt("cmix += c_R[%d] * sc[%d]; ehmix += h[%d] * wdot[%d];" % ((k,) * 4))
    e.stmt("const double cmixinv = 1.0 / cmix;")
    e.stmt("const double tmp1 = ehmix * cmixinv;")
    e.stmt("const double tmp3 = cmixinv * T;")
    e.stmt("const double tmp2 = tmp1 * tmp3;")
    for k in range(ns):  # column k: the dTdot/dy_k entry from the unscaled column, then the column scaled by W_m / W_k
        e.stmt("{")
        e.stmt("  double dehmixdc = 0.0, a;")
        for mm in range(ns):
            e.stmt("  a = J(%d, %d); dehmixdc += h[%d] * a; J(%d, %d) = lane_div(a * %s, %s);" %
                   (mm, k, mm, mm, k, e.k(mw[mm]), e.k(mw[k])))
        e.stmt("  J(NS, %d) = lane_div((tmp2 * c_R[%d] - tmp3 * dehmixdc) * 1.e6, %s);" % (k, k, e.k(mw[k])))
        e.stmt("}")
    for k in range(ns):
        e.stmt("J(%d, NS) = (jT[%d] * 1.e-6) * %s;" % (k, k, e.k(mw[k])))
    e.stmt("J(NS, NS) = 0.0;  // never written by the reference (ReactorCvodeJacobian.cpp:143-148): the matrix was zeroed")
    return e.text()


def emit_header(mech) -> str:
    m = mech["name"]
    ns, nr = mech["n_species"], mech["n_reactions"]
    pool = Pool()
    thermo = emit_thermo(mech, pool)
    prod = emit_production_rate(mech, pool)
    jac = emit_ajacobian(mech, pool)
    mw = [float(w) for w in mech["mw"]]
    out = []
    out.append("// GENERATED by pelelm_b200/mechgen/emit_lane.py from data/%s.json -- do not edit.\n" % m)
    out.append("// Straight-line lane = cell kernels of %s (%d species, %d reactions); see the generator for the reference lines.\n" % (m, ns, nr))
    out.append("#pragma once\n#include \"../lane_common.h\"\n\nnamespace plm {\nnamespace lane {\n\n")
    out.append("struct Mech_%s {\n" % m)
    out.append("  static constexpr int NS = %d, N = NS + 1, NR = %d;\n" % (ns, nr))
    out.append("  static constexpr const char* name() { return \"%s\"; }\n" % m)
    out.append("  static constexpr bool kTable = false;  // straight-line code (csrc/lane_chem.h)\n")
    # tables as functions of a constant index: the arrays live at namespace scope (a __constant__ cannot be a member)
    out.append("  PLM_LHD static double mw_of(int k);\n  PLM_LHD static double imw_of(int k);\n")
    out.append("%s" % thermo)
    out.append("  // productionRate :3650-4902 -- sc in mol/m^3, wdot in mol/m^3/s\n")
    out.append("  PLM_LHD static void production_rate(const double (&sc)[NS], double T, double (&wdot)[NS])\n  {\n%s  }\n" % prod)
    out.append("  // aJacobian (consP = 1) + cJac scaling: J(row, col) of d(ydot)/dy, y = [rhoY_k, T], sc = 1e6 * C (mol/m^3)\n")
    out.append("  template <class JA>\n  PLM_LHD static void ajacobian(JA& J, const double (&sc)[NS], double T)\n  {\n%s  }\n" % jac)
    out.append("};\n\n")
    kk = pool.vals

    def arr(name, vals):
        body = ",\n".join("    " + ", ".join(repr(float(v)) for v in vals[i:i + 4]) for i in range(0, len(vals), 4))
        return "PLM_LCONST double %s[%d] = {\n%s};\n" % (name, len(vals), body)

    pre = []
    pre.append(arr("KK_%s" % m, kk))
    pre.append(arr("MW_%s" % m, mw))
    pre.append(arr("IMW_%s" % m, [1.0 / w for w in mw]))
    text = "".join(out)
    # the constant pool precedes the struct; inside the struct `KK` is this mechanism's pool
    head, rest = text.split("struct Mech_%s {\n" % m, 1)
    text = head + "".join(pre) + "\nstruct Mech_%s {\n" % m + rest
    text = text.replace("KK[", "KK_%s[" % m)
    text += "PLM_LHD double Mech_%s::mw_of(int k) { return MW_%s[k]; }\n" % (m, m)
    text += "PLM_LHD double Mech_%s::imw_of(int k) { return IMW_%s[k]; }\n" % (m, m)
    text += "\n}  // namespace lane\n}  // namespace plm\n"
    return text


def main():
    outdir = os.path.join(ROOT, "pelelm_b200", "csrc", "gen")
    for m in ("LiDryer", "drm19", "dodecane_lu"):
        with open(os.path.join(outdir, "lane_%s.h" % m), "w") as f:
            f.write(emit_header(load(m)))
    print("lane kernels emitted")


# ---------------------------------------------------------------------------------------------------------------
# Table-driven lane = cell kernels: the mechanism as constant-memory tables (gen/lanetab_<mech>.h) for the loop-based
# evaluation in csrc/lane_tab.h.  Same data and order as the flat tables of emit_tables.py; one row per reaction so that
# every load of the loop body is a warp-uniform constant load.
# ---------------------------------------------------------------------------------------------------------------
def emit_tab_header(mech) -> str:
    from decimal import Decimal
    m = mech["name"]
    ns, nr = mech["n_species"], mech["n_reactions"]
    R = mech["reactions"]
    th = mech["thermo"]
    mw = [float(w) for w in mech["mw"]]

    def arr(ctype, name, vals, per=6, fmt=lambda v: repr(float(v))):
        vals = list(vals) or [0]
        body = ",\n".join("    " + ", ".join(fmt(v) for v in vals[i:i + per]) for i in range(0, len(vals), per))
        return "PLM_LCONST %s %s_%s[%d] = {\n%s};\n" % (ctype, name, m, len(vals), body)

    ii = lambda v: str(int(v))  # noqa: E731
    # thermo: [species][range][24] = cp[5] h[6] g[7] dcp[4] a0m1 pad
    TH = []
    for k in range(ns):
        for rg in (0, 1):
            row = list(th["cp"][k][rg]) + list(th["h"][k][rg]) + list(th["g"][k][rg]) + list(th["dcp"][k][rg])
            row.append(float(Decimal(repr(float(th["cp"][k][rg][0]))) - 1))
            row.append(0.0)
            assert len(row) == 24
            TH += row
    # reactions: [nr][12] doubles, [nr][4] ints
    RK, RF = [], []
    for r in R:
        tr = r.get("troe", [0, 0, 0, 0])
        typ = r["type"]
        t1 = typ == 3 and abs(tr[1]) > 1.0e-100 and (1.0 - tr[0]) != 0
        t2 = typ == 3 and abs(tr[2]) > 1.0e-100 and tr[0] != 0
        np4 = typ == 3 and bool(r.get("troe_has_T2"))
        RK += [r["A"], r["b"], r["Ea_R"], r.get("A0", 0.0), r.get("b0", 0.0), r.get("Ea0_R", 0.0),
               (1.0 - tr[0]) if typ == 3 else 0.0, tr[0], (1 / tr[1]) if t1 else 0.0, (1 / tr[2]) if t2 else 0.0, tr[3], r["dn"],
               tr[1], tr[2], 0.0, 0.0]
        # flags: bit0 term1, bit1 term2, bit2 T2 term, bit3 b and Ea both zero, bit4 Jacobian term1 (|T3| > 1e-100), bit5 term2
        fl = (1 if t1 else 0) | (2 if t2 else 0) | (4 if np4 else 0) | (8 if (r["b"] == 0.0 and r["Ea_R"] == 0.0) else 0) | \
            (16 if (typ == 3 and abs(tr[1]) > 1e-100) else 0) | (32 if (typ == 3 and abs(tr[2]) > 1e-100) else 0)
        RF += [typ, 1 if r["reversible"] else 0, fl, 0]

    def csr(lists):
        off, idx, val = [0], [], []
        for lst in lists:
            for i, v in lst:
                idx.append(int(i))
                val.append(float(v))
            off.append(len(idx))
        return off, idx, val

    reac = [sorted(r["reactants"]) for r in R]
    prod = [sorted(r["products"]) for r in R]
    net = [_net(r) for r in R]
    eff = [[(i, e - 1.0) for i, e in sorted(r["eff"])] for r in R]     # efficiency - 1
    # Jacobian columns of a reaction (consP = 1): every species the rate depends on, ascending; the third-body coefficient
    # (efficiency - 1) where listed, and where the species sits on either side
    jc = []
    for r, rr, pp, ee in zip(R, reac, prod, eff):
        cols = {}
        if r["type"] != 0:
            for i, e in ee:
                cols.setdefault(i, [0, 0.0])
                cols[i][0] |= 1
                cols[i][1] = e
        for i, _ in rr:
            cols.setdefault(i, [0, 0.0])[0] |= 2
        if r["reversible"]:
            for i, _ in pp:
                cols.setdefault(i, [0, 0.0])[0] |= 4
        jc.append(sorted((i, f, e) for i, (f, e) in cols.items()))
    joff, jidx, jfl, jef = [0], [], [], []
    for lst in jc:
        for i, f, e in lst:
            jidx.append(i)
            jfl.append(f)
            jef.append(e)
        joff.append(len(jidx))
    s = []
    s.append("// GENERATED by pelelm_b200/mechgen/emit_lane.py (emit_tab_header) from data/%s.json -- do not edit.\n" % m)
    s.append("// %s as constant-memory tables for the loop-based lane = cell kernels (csrc/lane_tab.h): %d species, %d reactions.\n" % (m, ns, nr))
    s.append("#pragma once\n#include \"../lane_common.h\"\n\nnamespace plm {\nnamespace lane {\n\n")
    s.append(arr("double", "LT_MW", mw))
    s.append(arr("double", "LT_IMW", [1.0 / w for w in mw]))
    s.append(arr("double", "LT_TMID", th["tmid"]))
    s.append(arr("double", "LT_TH", TH, per=6))
    s.append(arr("double", "LT_RK", RK, per=4))
    s.append(arr("int", "LT_RF", RF, per=16, fmt=ii))
    for nm, lists in (("R", reac), ("P", prod), ("N", net), ("E", eff)):
        off, idx, val = csr(lists)
        s.append(arr("int", "LT_%sOFF" % nm, off, per=24, fmt=ii))
        s.append(arr("int", "LT_%sIDX" % nm, idx, per=24, fmt=ii))
        s.append(arr("double", "LT_%sVAL" % nm, val, per=8))
    s.append(arr("int", "LT_JOFF", joff, per=24, fmt=ii))
    s.append(arr("int", "LT_JIDX", jidx, per=24, fmt=ii))
    s.append(arr("int", "LT_JFL", jfl, per=24, fmt=ii))
    s.append(arr("double", "LT_JEF", jef, per=8))
    s.append("\nstruct Tab_%s {\n" % m)
    s.append("  static constexpr int NS = %d, N = NS + 1, NR = %d;\n" % (ns, nr))
    s.append("  static constexpr const char* name() { return \"%s\"; }\n" % m)
    s.append("  static constexpr bool kTable = true;  // loops over constant-memory tables (csrc/lane_tab.h)\n")
    for nm in ("MW", "IMW", "TMID", "TH", "RK", "ROFF", "RIDX", "RVAL", "POFF", "PIDX", "PVAL", "NOFF", "NIDX", "NVAL", "EOFF", "EIDX",
               "EVAL", "JOFF", "JIDX", "JFL", "JEF", "RF"):
        typ = "int" if nm in ("RF", "ROFF", "RIDX", "POFF", "PIDX", "NOFF", "NIDX", "EOFF", "EIDX", "JOFF", "JIDX", "JFL") else "double"
        s.append("  PLM_LHD static %s %s(int i) { return LT_%s_%s[i]; }\n" % (typ, nm.lower(), nm, m))
    s.append("  PLM_LHD static double mw_of(int k) { return LT_MW_%s[k]; }\n  PLM_LHD static double imw_of(int k) { return LT_IMW_%s[k]; }\n" % (m, m))
    s.append("};\n\n}  // namespace lane\n}  // namespace plm\n")
    return "".join(s)


def main_tab():
    outdir = os.path.join(ROOT, "pelelm_b200", "csrc", "gen")
    for m in ("LiDryer", "drm19", "dodecane_lu"):
        with open(os.path.join(outdir, "lanetab_%s.h" % m), "w") as f:
            f.write(emit_tab_header(load(m)))
    print("lane tables emitted")


if __name__ == "__main__":
    main()
    main_tab()
