// Lane = cell mechanism kernels, loop form: one THREAD evaluates one cell, looping over species and reactions with the
// mechanism read from constant-memory tables (gen/lanetab_<mech>.h).  Every table load of a loop body is warp-uniform
// (one constant-cache access for 32 cells); the only per-lane table index is the NASA-7 temperature range.  The code of a
// whole right-hand side is a few KB, so the warps of an SM run it out of the instruction cache -- the straight-line form of
// the same arithmetic (gen/lane_<mech>.h) is ~180 KB for drm19 and is fetch-bound there (profiles/r02b_*).
//
// Arithmetic, order of operations and special cases follow the reference's generated mechanism code reaction by reaction
// (drm19 line numbers; SURVEY.md section 8a lists the other mechanisms):
//   lt_mix_hc / lt_species_h   CKHBMS :1712, CKCPBS :1605, CKHMS :2770 (EINT: CKUBMS, CKCVBS, CKUMS)
//   lt_get_T                   GET_T_GIVEN_HY :1967-2018 / GET_T_GIVEN_EY :1912-1963
//   lt_production_rate         productionRate :3650-4902 (gibbs :531), CKWC :4905-4923
//   lt_rhs                     utils::fKernelSpec, PelePhysics/Reactions/ReactorUtils.H:274-332
//   lt_cjac                    aJacobian :9696-14742 (consP = 1) + cvode::cJac's scaling, ReactorCvodeJacobian.cpp:84-153
// Per-cell arrays (concentrations, g/RT, production rates ...) are lane-strided vectors (LV): shared memory on the device.
#pragma once
#include "lane_common.h"

namespace plm {
namespace lane {

// element k of this lane's vector; the lanes of a warp interleave (stride PLM_LANE_W doubles)
template <int STRIDE>
struct LVt {
  double* p;
  PLM_LHD double& operator()(int k) const { return p[size_t(k) * STRIDE]; }
};
using LV = LVt<PLM_LANE_W>;

constexpr double kRUc = 8.31446261815324e+07;  // mechanism.H:145

// NASA-7 rows: [species][range][24] = cp/R a0..a4 | h/RT a0, a1/2 .. a4/5, a5 | g/RT a5, a0-a6, -a0, -a1/2, -a2/6, -a3/12,
// -a4/20 | d(cp/R)/dT a1, 2a2, 3a3, 4a4 | a0 - 1 (cv/R and e/RT constant term) | pad
template <class M>
PLM_LHD int th_row(int k, double T)
{
  return (k * 2 + (T < M::tmid(k) ? 0 : 1)) * 24;
}

// he = sum Y_k (h_k/RT) / W_k * RT, c = sum (c_k/R) Y_k / W_k * R with Y_k = y_k * rho_inv
template <class M, int EINT, class V>
PLM_LHD void lt_mix_hc(double T, const V& y, double rho_inv, double& he, double& c)
{
  const double T2 = T * T, T3 = T * T * T, T4 = T * T * T * T, invT = 1 / T;
  double hs = 0.0, cs = 0.0;
#pragma unroll 1
  for (int k = 0; k < M::NS; ++k) {
    const int r = th_row<M>(k, T);
    const double Yk = y(k) * rho_inv, imw = M::imw(k);
    const double h0 = EINT ? M::th(r + 22) : M::th(r + 5), c0 = EINT ? M::th(r + 22) : M::th(r + 0);
    const double h = h0 + M::th(r + 6) * T + M::th(r + 7) * T2 + M::th(r + 8) * T3 + M::th(r + 9) * T4 + M::th(r + 10) * invT;
    const double cp = c0 + M::th(r + 1) * T + M::th(r + 2) * T2 + M::th(r + 3) * T3 + M::th(r + 4) * T4;
    hs += Yk * h * imw;
    cs += cp * Yk * imw;
  }
  he = hs * (kRUc * T);
  c = cs * kRUc;
}

// the reference's range guard, verbatim; only reached when the guess or the iterate leaves [tmin, tmax]
template <class M, int EINT, class V>
PLM_LNOINL double lt_get_T_guarded(double hin, const V& y, double rho_inv, double Tguess, double& c_out)
{
  const double tmin = 90, tmax = 4000, tol = 1.e-6;
  double hmin, cmin, hmax, cmax;
  lt_mix_hc<M, EINT>(tmin, y, rho_inv, hmin, cmin);
  lt_mix_hc<M, EINT>(tmax, y, rho_inv, hmax, cmax);
  double t1, h1, c = cmin;
  if (hin < hmin || hin > hmax) {
    t1 = (hin < hmin) ? tmin - (hmin - hin) / cmin : tmax - (hmax - hin) / cmax;
    lt_mix_hc<M, EINT>(t1, y, rho_inv, h1, c_out);
    return t1;
  }
  t1 = Tguess;
  if (!(t1 >= tmin && t1 <= tmax)) t1 = tmin + (tmax - tmin) / (hmax - hmin) * (hin - hmin);
  bool fresh = false;
  for (int i = 0; i < 200; ++i) {
    lt_mix_hc<M, EINT>(t1, y, rho_inv, h1, c);
    fresh = true;
    double dt = (hin - h1) / c;
    if (dt > 100.) {
      dt = 100.;
    } else if (dt < -100.) {
      dt = -100.;
    } else if (lane_fabs(dt) < tol) {
      break;
    } else if (t1 + dt == t1) {
      break;
    }
    t1 += dt;
    fresh = false;
    if (t1 != t1) break;  // NaN: the reference iterates 200 times on it and returns NaN; same result, no work
  }
  if (!fresh) lt_mix_hc<M, EINT>(t1, y, rho_inv, h1, c);
  c_out = c;
  return t1;
}

// Newton iteration on T (tol 1e-6 K, <= 200 iterations, +-100 K clamp; the step that passes the |dt| < tol test is NOT
// applied).  c_out = cp (or cv) of the mixture at the returned T.
template <class M, int EINT, class V>
PLM_LHD double lt_get_T(double hin, const V& y, double rho_inv, double Tguess, double& c_out)
{
  const double tmin = 90, tmax = 4000, tol = 1.e-6;
  double t1 = Tguess, h1, c = 0.0;
  bool ok = (t1 >= tmin && t1 <= tmax);  // false for a NaN guess
  if (ok) {
    bool fresh = false;
#pragma unroll 1
    for (int i = 0; i < 200; ++i) {
      lt_mix_hc<M, EINT>(t1, y, rho_inv, h1, c);
      fresh = true;
      double dt = lane_div(hin - h1, c);
      if (dt > 100.) {
        dt = 100.;
      } else if (dt < -100.) {
        dt = -100.;
      } else if (lane_fabs(dt) < tol) {
        break;
      } else if (t1 + dt == t1) {
        break;
      }
      t1 += dt;
      fresh = false;
      if (t1 != t1) break;  // NaN: the reference iterates 200 times on it and returns NaN; same result, no work
    }
    ok = (t1 >= tmin && t1 <= tmax);
    if (ok && !fresh) lt_mix_hc<M, EINT>(t1, y, rho_inv, h1, c);
  }
  if (!ok) return lt_get_T_guarded<M, EINT>(hin, y, rho_inv, Tguess, c_out);
  c_out = c;
  return t1;
}

// phi = prod c^nu over one side of a reaction (CSR slice [a, b)); the generated code writes c*c for nu = 2
template <class M, class V>
PLM_LHD double lt_phase(const V& sc, int a, int b, bool reactants)
{
  double phi = 1.0;
#pragma unroll 1
  for (int s = a; s < b; ++s) {
    const double c = sc(reactants ? M::ridx(s) : M::pidx(s));
    const double nu = reactants ? M::rval(s) : M::pval(s);
    const double f = (nu == 1.0) ? c : (nu == 2.0 ? c * c : lane_pow(c, nu));
    phi = (s == a) ? f : phi * f;
  }
  return phi;
}

// productionRate: sc in mol/m^3 -> wdot in mol/m^3/s; g = scratch for g/RT
template <class M, class V>
PLM_LHD void lt_production_rate(const V& wdot, const V& sc, const V& g, double T)
{
  const double logT = lane_log(T);
  const double invT = lane_rcp(T);
  const double refC = (101325 / 8.31446) * invT;
  const double refCinv = lane_rcp(refC);
  const double T2 = T * T, T3 = T * T * T, T4 = T * T * T * T;
  double mixture = 0.0;
#pragma unroll 1
  for (int k = 0; k < M::NS; ++k) {
    const int r = th_row<M>(k, T);
    g(k) = M::th(r + 11) * invT + M::th(r + 12) + M::th(r + 13) * logT + M::th(r + 14) * T + M::th(r + 15) * T2 +
           M::th(r + 16) * T3 + M::th(r + 17) * T4;
    mixture += sc(k);
    wdot(k) = 0.0;
  }
#pragma unroll 1
  for (int r = 0; r < M::NR; ++r) {
    const int type = M::rf(4 * r), rev = M::rf(4 * r + 1), fl = M::rf(4 * r + 2);
    const int ko = 16 * r;
    double k_f = M::rk(ko);
    if (!(fl & 8)) k_f = k_f * lane_exp(M::rk(ko + 1) * logT - M::rk(ko + 2) * invT);  // b or Ea non-zero (a zero term drops out exactly)
    double Corr = 1.0;
    if (type != 0) {
      Corr = mixture;
#pragma unroll 1
      for (int s = M::eoff(r); s < M::eoff(r + 1); ++s) {
        const double em1 = M::eval(s);
        Corr += (em1 == 1.0) ? sc(M::eidx(s)) : em1 * sc(M::eidx(s));
      }
      if (type >= 2) {
        const double redP = Corr / k_f * M::rk(ko + 3) * lane_exp(M::rk(ko + 4) * logT - M::rk(ko + 5) * invT);
        const double F = redP / (1.0 + redP);
        if (type == 3) {
          const double logPred = lane_log10(redP);
          double fc = 0.0;
          if (fl & 1) fc = M::rk(ko + 6) * lane_exp(-T * M::rk(ko + 8));
          if (fl & 2) fc = fc + M::rk(ko + 7) * lane_exp(-T * M::rk(ko + 9));
          if (fl & 4) fc = fc + lane_exp(-M::rk(ko + 10) * invT);
          const double logFcent = lane_log10(fc);
          const double troe_c = -0.4 - 0.67 * logFcent;
          const double troe_n = 0.75 - 1.27 * logFcent;
          const double troe = (troe_c + logPred) / (troe_n - 0.14 * (troe_c + logPred));
          const double F_troe = lane_exp10(logFcent / (1.0 + troe * troe));
          Corr = F * F_troe;
        } else {
          Corr = F;
        }
      }
    }
    const double phi_f = lt_phase<M>(sc, M::roff(r), M::roff(r + 1), true);
    const double qf = (type != 0) ? Corr * k_f * phi_f : k_f * phi_f;
    double qr = 0.0;
    const int n0 = M::noff(r), n1 = M::noff(r + 1);
    if (rev) {
      double dG = 0.0;  // sum nu' g - sum nu'' g, species order
#pragma unroll 1
      for (int s = n0; s < n1; ++s) {
        const double c = -M::nval(s);
        const double gk = g(M::nidx(s));
        const double term = (c == 1.0 || c == -1.0) ? gk : lane_fabs(c) * gk;
        if (s == n0) dG = (c > 0) ? term : -term;
        else dG = (c > 0) ? dG + term : dG - term;
      }
      const double phi_r = lt_phase<M>(sc, M::poff(r), M::poff(r + 1), false);
      double kr = (type != 0) ? Corr * k_f * lane_exp(-(dG)) : k_f * lane_exp(-(dG));
      const double dn = M::rk(ko + 11);
      if (dn != 0.0)
        kr = kr * ((dn == -1.0) ? refC : (dn == 1.0 ? refCinv : (dn < 0 ? lane_pow(refC, -dn) : lane_pow(refCinv, dn))));
      qr = kr * phi_r;
    }
    const double qdot = qf - qr;
#pragma unroll 1
    for (int s = n0; s < n1; ++s) {
      const double c = M::nval(s);
      const int k = M::nidx(s);
      if (c == 1.0) wdot(k) += qdot;
      else if (c == -1.0) wdot(k) -= qdot;
      else if (c > 0) wdot(k) += c * qdot;
      else wdot(k) -= (-c) * qdot;
    }
  }
}

// fKernelSpec for one cell: y = [rhoY_k, T guess] -> f; sc, g: scratch vectors of NS entries; rys(k) = external species
// source; returns the temperature found.  only_T: stop after the temperature solve (f untouched).
template <class M, int EINT, class SRC, class V>
PLM_LHD double lt_rhs(const V& y, const V& f, const V& sc, const V& g, double hin_rho, const SRC& rys, bool only_T)
{
  constexpr int ns = M::NS;
  double rho = 0.0;
#pragma unroll 1
  for (int k = 0; k < ns; ++k) rho += y(k);
  const double rho_inv = lane_rcp(rho);
  const double nrg = hin_rho * rho_inv;  // (rhoe_init + rhoesrc_ext * dt_save) / rho
  double c;
  const double T = lt_get_T<M, EINT>(nrg, y, rho_inv, y(ns), c);
  if (only_T) return T;
  if (T != T) {  // a NaN state: every component of the reference's result is NaN as well
#pragma unroll 1
    for (int k = 0; k <= ns; ++k) f(k) = T;
    return T;
  }
#pragma unroll 1
  for (int k = 0; k < ns; ++k) sc(k) = (rho * (y(k) * rho_inv) * M::imw(k)) * 1.0e6;  // CKYTCR :2403, CKWC :4911
  lt_production_rate<M>(f, sc, g, T);
  // CKHMS / CKUMS and the energy equation
  const double T2 = T * T, T3 = T * T * T, T4 = T * T * T * T, invT = 1 / T;
  const double RT = kRUc * T;
  double rhoesrc = rys.energy();
#pragma unroll 1
  for (int k = 0; k < ns; ++k) {
    const int r = th_row<M>(k, T);
    const double h0 = EINT ? M::th(r + 22) : M::th(r + 5);
    const double ei = (h0 + M::th(r + 6) * T + M::th(r + 7) * T2 + M::th(r + 8) * T3 + M::th(r + 9) * T4 + M::th(r + 10) * invT) *
                      (RT * M::imw(k));
    const double cdot_rYs = (f(k) * 1.0e-6) * M::mw(k) + rys(k);
    f(k) = cdot_rYs;
    rhoesrc -= cdot_rYs * ei;
  }
  f(ns) = rhoesrc * lane_div(rho_inv, c);
  return T;
}

// d(phi)/d(sc[j]) over one side (CSR slice); the running product starts at 1.0 as in the generated code
template <class M, class V>
PLM_LHD double lt_dphase(const V& sc, int a, int b, int j, bool reactants)
{
  double d = 1.0;
#pragma unroll 1
  for (int s = a; s < b; ++s) {
    const int i = reactants ? M::ridx(s) : M::pidx(s);
    const double nu = reactants ? M::rval(s) : M::pval(s);
    const double c = sc(i);
    if (i == j) {
      if (nu == 1.0) d *= 1.0;
      else if (nu == 2.0) d *= 2.0 * c;
      else d *= nu * lane_pow(c, nu - 1.0);
    } else {
      d *= (nu == 1.0) ? c : lane_pow(c, nu);
    }
  }
  return d;
}

// cJac for one cell: J(row, col) <- d ydot / d y, y = [rhoY_k, T].  sc, g: scratch vectors (NS) in fast memory; h, wd, jT, cR:
// scratch vectors (NS) anywhere.  The species block accumulates in place in J, reaction by reaction and column by column in
// the reference's order; the T column accumulates in jT; the closure (:14705-14741) and cJac's scaling run column-wise at
// the end.  The (T,T) entry is never written by the reference (ReactorCvodeJacobian.cpp:143-148): 0.
template <class M, class JA, class V>
PLM_LHD void lt_cjac(const V& y, JA& J, const V& sc, const V& g, const V& h, const V& wd, const V& jT, const V& cR)
{
  constexpr int ns = M::NS;
  double rho = 0.0;
#pragma unroll 1
  for (int k = 0; k < ns; ++k) rho = rho + y(k);
  const double T = y(ns);
  const double logT = lane_log(T);
  const double invT = lane_rcp(T);
  const double invT2 = invT * invT;
  const double refC = lane_div(101325 / 8.31446, T);
  const double refCinv = lane_rcp(refC);
  const double ln10 = 2.302585092994046, log10e = 1.0 / 2.302585092994046;
  const double T2 = T * T, T3 = T * T * T, T4 = T * T * T * T;
  double mixture = 0.0;
#pragma unroll 1
  for (int k = 0; k < ns; ++k) {
    const double massfrac = y(k) / rho;
    const double c = 1.e6 * (rho * massfrac * M::imw(k));
    sc(k) = c;
    mixture += c;
    const int r = th_row<M>(k, T);
    g(k) = M::th(r + 11) * invT + M::th(r + 12) + M::th(r + 13) * logT + M::th(r + 14) * T + M::th(r + 15) * T2 +
           M::th(r + 16) * T3 + M::th(r + 17) * T4;
    h(k) = M::th(r + 5) + M::th(r + 6) * T + M::th(r + 7) * T2 + M::th(r + 8) * T3 + M::th(r + 9) * T4 + M::th(r + 10) * invT;
    cR(k) = M::th(r + 0) + M::th(r + 1) * T + M::th(r + 2) * T2 + M::th(r + 3) * T3 + M::th(r + 4) * T4;
    wd(k) = 0.0;
    jT(k) = 0.0;
  }
#pragma unroll 1
  for (int c = 0; c < ns; ++c) {
#pragma unroll 1
    for (int r = 0; r < ns; ++r) J(r, c) = 0.0;
  }
#pragma unroll 1
  for (int r = 0; r < M::NR; ++r) {
    const int type = M::rf(4 * r), rev = M::rf(4 * r + 1), fl = M::rf(4 * r + 2);
    const int ko = 16 * r;
    double alpha = 0.0;
    if (type != 0) {
      alpha = mixture;
#pragma unroll 1
      for (int s = M::eoff(r); s < M::eoff(r + 1); ++s) alpha += M::eval(s) * sc(M::eidx(s));
    }
    const int r0 = M::roff(r), r1 = M::roff(r + 1), p0 = M::poff(r), p1 = M::poff(r + 1), n0 = M::noff(r), n1 = M::noff(r + 1);
    const double phi_f = lt_phase<M>(sc, r0, r1, true);
    const double b = M::rk(ko + 1), EaR = M::rk(ko + 2);
    double k_f = M::rk(ko) * lane_exp(b * logT - EaR * invT);
    const double dlnkfdT = b * invT + EaR * invT2;
    double Pr = 0, fPr = 0, F = 1.0, dlogfPrdT = 0, dlogFdT = 0, dlogFdlogPr = 0;
    if (type >= 2) {
      const double b0 = M::rk(ko + 4), Ea0R = M::rk(ko + 5);
      const double k_0 = M::rk(ko + 3) * lane_exp(b0 * logT - Ea0R * invT);
      Pr = alpha / k_f * k_0;
      fPr = Pr / (1.0 + Pr);
      const double dlnk0dT = b0 * invT + Ea0R * invT2;
      const double dlogPrdT = log10e * (dlnk0dT - dlnkfdT);
      dlogfPrdT = dlogPrdT / (1.0 + Pr);
      if (type == 3) {
        const double tr0 = M::rk(ko + 7), tr1 = M::rk(ko + 12), tr2 = M::rk(ko + 13), tr3 = M::rk(ko + 10);
        const double logPr = lane_log10(Pr);
        const double Fcent1 = (fl & 16) ? (1. - tr0) * lane_exp(-T / tr1) : 0.;
        const double Fcent2 = (fl & 32) ? tr0 * lane_exp(-T / tr2) : 0.;
        const double Fcent3 = (fl & 4) ? lane_exp(-tr3 * invT) : 0.;
        const double Fcent = Fcent1 + Fcent2 + Fcent3;
        const double logFcent = lane_log10(Fcent);
        const double troe_c = -.4 - .67 * logFcent;
        const double troe_n = .75 - 1.27 * logFcent;
        const double troePr_den = 1.0 / (troe_n - .14 * (troe_c + logPr));
        const double troePr = (troe_c + logPr) * troePr_den;
        const double troe = 1.0 / (1.0 + troePr * troePr);
        F = lane_exp10(logFcent * troe);
        double dlogFcentdT = -Fcent1 / tr1 - Fcent2 / tr2;
        if (fl & 4) dlogFcentdT += Fcent3 * tr3 * invT2;
        dlogFcentdT = log10e / Fcent * dlogFcentdT;
        const double dlogFdcn_fac = 2.0 * logFcent * troe * troe * troePr * troePr_den;
        const double dlogFdc = -troe_n * dlogFdcn_fac * troePr_den;
        const double dlogFdn = dlogFdcn_fac * troePr;
        dlogFdlogPr = dlogFdc;
        dlogFdT = dlogFcentdT * (troe - 0.67 * dlogFdc - 1.27 * dlogFdn) + dlogFdlogPr * dlogPrdT;
      }
    }
    double phi_r = 0.0, k_r = 0.0, dkrdT = 0.0;
    if (rev) {
      phi_r = lt_phase<M>(sc, p0, p1, false);
      double dG = 0.0;
#pragma unroll 1
      for (int s = n0; s < n1; ++s) dG += -M::nval(s) * g(M::nidx(s));
      double Kc = lane_exp(dG);
      const double dn = M::rk(ko + 11);
      if (dn != 0.0) Kc = ((dn == -1.0) ? refCinv : (dn == 1.0 ? refC : lane_pow(refC, dn))) * Kc;
      k_r = k_f / Kc;
      double hr = 0.0, hp = 0.0;
#pragma unroll 1
      for (int s = r0; s < r1; ++s) hr += M::rval(s) * h(M::ridx(s));
#pragma unroll 1
      for (int s = p0; s < p1; ++s) hp += M::pval(s) * h(M::pidx(s));
      const double dlnKcdT = invT * (-(hr) + (hp) + (-dn));
      dkrdT = (dlnkfdT - dlnKcdT) * k_r;
    }
    double q, dqdT, dcdc_fac = 0.0;
    if (type == 0) {
      q = k_f * phi_f - k_r * phi_r;
      dqdT = (dlnkfdT * k_f * phi_f - dkrdT * phi_r);
    } else if (type == 1) {
      const double q_nocor = k_f * phi_f - k_r * phi_r;
      q = alpha * q_nocor;
      dqdT = alpha * (dlnkfdT * k_f * phi_f - dkrdT * phi_r);
      k_f *= alpha;
      k_r *= alpha;
      dcdc_fac = q_nocor;
    } else {
      const double q_nocor = k_f * phi_f - k_r * phi_r;
      const double Corr = fPr * F;
      q = Corr * q_nocor;
      const double dlnCorrdT = ln10 * (dlogfPrdT + dlogFdT);
      dqdT = Corr * (dlnkfdT * k_f * phi_f - dkrdT * phi_r) + dlnCorrdT * q;
      k_f *= Corr;
      k_r *= Corr;
      dcdc_fac = q / alpha * (1.0 / (Pr + 1.0) + dlogFdlogPr);
    }
#pragma unroll 1
    for (int s = n0; s < n1; ++s) {
      wd(M::nidx(s)) += M::nval(s) * q;
      jT(M::nidx(s)) += M::nval(s) * dqdT;
    }
    // columns: every species the rate depends on, ascending
#pragma unroll 1
    for (int c = M::joff(r); c < M::joff(r + 1); ++c) {
      const int j = M::jidx(c), jf = M::jfl(c);
      double dqdci = 0.0;
      if (jf & 1) dqdci = M::jef(c) * dcdc_fac;
      if (jf & 2) dqdci += k_f * lt_dphase<M>(sc, r0, r1, j, true);
      if (jf & 4) dqdci -= k_r * lt_dphase<M>(sc, p0, p1, j, false);
#pragma unroll 1
      for (int s = n0; s < n1; ++s) J(M::nidx(s), j) += M::nval(s) * dqdci;
    }
  }
  // closure (consP) fused with the scaling of cJac
  double cmix = 0.0, ehmix = 0.0;
#pragma unroll 1
  for (int k = 0; k < ns; ++k) {
    cmix += cR(k) * sc(k);
    ehmix += h(k) * wd(k);
  }
  const double cmixinv = 1.0 / cmix;
  const double tmp1 = ehmix * cmixinv;
  const double tmp3 = cmixinv * T;
  const double tmp2 = tmp1 * tmp3;
#pragma unroll 1
  for (int k = 0; k < ns; ++k) {  // column k: the dTdot/dy_k entry from the unscaled column, then the column scaled by W_m / W_k
    double dehmixdc = 0.0;
    const double mwk = M::mw(k);
#pragma unroll 1
    for (int mm = 0; mm < ns; ++mm) {
      const double a = J(mm, k);
      dehmixdc += h(mm) * a;
      J(mm, k) = lane_div(a * M::mw(mm), mwk);
    }
    J(ns, k) = lane_div((tmp2 * cR(k) - tmp3 * dehmixdc) * 1.e6, mwk);
  }
#pragma unroll 1
  for (int k = 0; k < ns; ++k) J(k, ns) = (jT(k) * 1.e-6) * M::mw(k);
  J(ns, ns) = 0.0;
}

}  // namespace lane
}  // namespace plm
