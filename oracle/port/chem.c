/* TEST INFRASTRUCTURE -- CPU restatement ("port") of the reference's per-cell chemistry kernels.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it; the product
 * (pelelm_b200/csrc) never links or calls anything here.
 *
 * Parity status: PINNED -- every function below is checked against the compiled, unmodified
 * reference (oracle/_ref, built by oracle/Makefile) in tests/test_oracle_port_vs_ref.py, and
 * against committed fixtures generated from it (tests/golden/) where /root/reference is absent.
 *
 * The mechanism is table driven (pelelm_b200/csrc/gen/tables_<mech>.c: DATA shared with the product) instead of generated straight-line
 * code; evaluation order per reaction follows the reference's generated kernels so results agree to
 * rounding. Reference locations cited are for drm19 (Support/Mechanism/Models/drm19/mechanism.H);
 * LiDryer and dodecane_lu have the same functions at other lines.
 */
#include "oracle.h"

#include <math.h>
#include <string.h>

#define RU 8.31446261815324e+07 /* mechanism.H:145 */

const plm_mech_tables* orc_mech(const char* name)
{
  if (!strcmp(name, "LiDryer")) return &plm_tables_LiDryer;
  if (!strcmp(name, "drm19")) return &plm_tables_drm19;
  if (!strcmp(name, "dodecane_lu")) return &plm_tables_dodecane_lu;
  return 0;
}

/* ---- NASA-7 polynomials: cp_R :342, gibbs :531, speciesEnthalpy :1207, dcvpRdT :5114 ---- */
void orc_cp_R(const plm_mech_tables* m, double* s, const double* tc)
{
  const double T = tc[1];
  for (int k = 0; k < m->ns; ++k) {
    const double* c = m->cp + (k * 2 + (T < m->tmid[k] ? 0 : 1)) * 5;
    s[k] = c[0] + c[1] * tc[1] + c[2] * tc[2] + c[3] * tc[3] + c[4] * tc[4];
  }
}
void orc_h_RT(const plm_mech_tables* m, double* s, const double* tc)
{
  const double T = tc[1], invT = 1 / T;
  for (int k = 0; k < m->ns; ++k) {
    const double* c = m->h + (k * 2 + (T < m->tmid[k] ? 0 : 1)) * 6;
    s[k] = c[0] + c[1] * tc[1] + c[2] * tc[2] + c[3] * tc[3] + c[4] * tc[4] + c[5] * invT;
  }
}
void orc_g_RT(const plm_mech_tables* m, double* s, const double* tc)
{
  const double T = tc[1], invT = 1 / T;
  for (int k = 0; k < m->ns; ++k) {
    const double* c = m->g + (k * 2 + (T < m->tmid[k] ? 0 : 1)) * 7;
    s[k] = c[0] * invT + c[1] + c[2] * tc[0] + c[3] * tc[1] + c[4] * tc[2] + c[5] * tc[3] + c[6] * tc[4];
  }
}
void orc_dcp_R(const plm_mech_tables* m, double* s, const double* tc)
{
  const double T = tc[1];
  for (int k = 0; k < m->ns; ++k) {
    const double* c = m->dcp + (k * 2 + (T < m->tmid[k] ? 0 : 1)) * 4;
    s[k] = c[0] + c[1] * tc[1] + c[2] * tc[2] + c[3] * tc[3];
  }
}

/* cv_R, speciesInternalEnergy (LiDryer mechanism.H:115, :505): the cp_R / speciesEnthalpy polynomials with the constant term
 * a0 - 1 (tables: a0m1) */
void orc_cv_R(const plm_mech_tables* m, double* s, const double* tc)
{
  const double T = tc[1];
  for (int k = 0; k < m->ns; ++k) {
    const int rg = (T < m->tmid[k] ? 0 : 1);
    const double* c = m->cp + (k * 2 + rg) * 5;
    s[k] = m->a0m1[k * 2 + rg] + c[1] * tc[1] + c[2] * tc[2] + c[3] * tc[3] + c[4] * tc[4];
  }
}
void orc_e_RT(const plm_mech_tables* m, double* s, const double* tc)
{
  const double T = tc[1], invT = 1 / T;
  for (int k = 0; k < m->ns; ++k) {
    const int rg = (T < m->tmid[k] ? 0 : 1);
    const double* c = m->h + (k * 2 + rg) * 6;
    s[k] = m->a0m1[k * 2 + rg] + c[1] * tc[1] + c[2] * tc[2] + c[3] * tc[3] + c[4] * tc[4] + c[5] * invT;
  }
}
/* CKCVBS (LiDryer mechanism.H:835), CKUBMS :933, CKUMS :1708 */
double orc_cvbs(const plm_mech_tables* m, double T, const double* y)
{
  const double tc[5] = {0, T, T * T, T * T * T, T * T * T * T};
  double cvor[ORC_MAXS], result = 0;
  orc_cv_R(m, cvor, tc);
  for (int i = 0; i < m->ns; ++i) result += cvor[i] * y[i] * (1.0 / m->mw[i]);
  return result * RU;
}
double orc_ubms(const plm_mech_tables* m, double T, const double* y)
{
  const double tc[5] = {0, T, T * T, T * T * T, T * T * T * T};
  double ums[ORC_MAXS], result = 0;
  const double RT = RU * T;
  orc_e_RT(m, ums, tc);
  for (int i = 0; i < m->ns; ++i) result += y[i] * ums[i] * (1.0 / m->mw[i]);
  return result * RT;
}
void orc_ums(const plm_mech_tables* m, double T, double* ums)
{
  const double tc[5] = {0, T, T * T, T * T * T, T * T * T * T};
  const double RT = RU * T;
  orc_e_RT(m, ums, tc);
  for (int i = 0; i < m->ns; ++i) ums[i] *= RT * (1.0 / m->mw[i]);
}
/* GET_T_GIVEN_EY (drm19 mechanism.H:1912-1963): same iteration as GET_T_GIVEN_HY on e(T), cv */
int orc_get_T_given_eY(const plm_mech_tables* m, double ein, const double* y, double* t)
{
  const int maxiter = 200;
  const double tol = 1.e-6;
  double tmin = 90, tmax = 4000;
  double e1, cv, t1, dt;
  const double emin = orc_ubms(m, tmin, y);
  const double emax = orc_ubms(m, tmax, y);
  if (ein < emin) {
    cv = orc_cvbs(m, tmin, y);
    *t = tmin - (emin - ein) / cv;
    return 1;
  }
  if (ein > emax) {
    cv = orc_cvbs(m, tmax, y);
    *t = tmax - (emax - ein) / cv;
    return 1;
  }
  t1 = *t;
  if (t1 < tmin || t1 > tmax) t1 = tmin + (tmax - tmin) / (emax - emin) * (ein - emin);
  for (int i = 0; i < maxiter; ++i) {
    e1 = orc_ubms(m, t1, y);
    cv = orc_cvbs(m, t1, y);
    dt = (ein - e1) / cv;
    if (dt > 100.) {
      dt = 100.;
    } else if (dt < -100.) {
      dt = -100.;
    } else if (fabs(dt) < tol) {
      break;
    } else if (t1 + dt == t1) {
      break;
    }
    t1 += dt;
  }
  *t = t1;
  return 0;
}

/* CKCPBS :1605-1626 */
double orc_cpbs(const plm_mech_tables* m, double T, const double* y)
{
  const double tc[5] = {0, T, T * T, T * T * T, T * T * T * T};
  double cpor[ORC_MAXS], tr[ORC_MAXS], result = 0;
  orc_cp_R(m, cpor, tc);
  for (int i = 0; i < m->ns; ++i) tr[i] = cpor[i] * y[i] * (1.0 / m->mw[i]);
  for (int i = 0; i < m->ns; ++i) result += tr[i];
  return result * RU;
}
/* CKHBMS :1712-1733 */
double orc_hbms(const plm_mech_tables* m, double T, const double* y)
{
  const double tc[5] = {0, T, T * T, T * T * T, T * T * T * T};
  double hml[ORC_MAXS], tmp[ORC_MAXS], result = 0;
  const double RT = RU * T;
  orc_h_RT(m, hml, tc);
  for (int i = 0; i < m->ns; ++i) tmp[i] = y[i] * hml[i] * (1.0 / m->mw[i]);
  for (int i = 0; i < m->ns; ++i) result += tmp[i];
  return result * RT;
}
/* CKHMS :2770-2786 */
void orc_hms(const plm_mech_tables* m, double T, double* hms)
{
  const double tc[5] = {0, T, T * T, T * T * T, T * T * T * T};
  const double RT = RU * T;
  orc_h_RT(m, hms, tc);
  for (int i = 0; i < m->ns; ++i) hms[i] *= RT * (1.0 / m->mw[i]);
}

/* GET_T_GIVEN_HY :1967-2018 (tol 1e-6 K, <=200 iterations, +-100 K clamp; the step that passes
 * the |dt|<tol test is NOT applied). Returns ierr. */
int orc_get_T_given_hY(const plm_mech_tables* m, double hin, const double* y, double* t)
{
  const int maxiter = 200;
  const double tol = 1.e-6;
  double tmin = 90, tmax = 4000;
  double h1, cp, t1, dt;
  const double hmin = orc_hbms(m, tmin, y);
  const double hmax = orc_hbms(m, tmax, y);
  if (hin < hmin) {
    cp = orc_cpbs(m, tmin, y);
    *t = tmin - (hmin - hin) / cp;
    return 1;
  }
  if (hin > hmax) {
    cp = orc_cpbs(m, tmax, y);
    *t = tmax - (hmax - hin) / cp;
    return 1;
  }
  t1 = *t;
  if (t1 < tmin || t1 > tmax) t1 = tmin + (tmax - tmin) / (hmax - hmin) * (hin - hmin);
  for (int i = 0; i < maxiter; ++i) {
    h1 = orc_hbms(m, t1, y);
    cp = orc_cpbs(m, t1, y);
    dt = (hin - h1) / cp;
    if (dt > 100.) {
      dt = 100.;
    } else if (dt < -100.) {
      dt = -100.;
    } else if (fabs(dt) < tol) {
      break;
    } else if (t1 + dt == t1) {
      break;
    }
    t1 += dt;
  }
  *t = t1;
  return 0;
}

static double phase(const double* sc, const int* idx, const double* nu, int a, int b)
{
  double phi = 1.0;
  int first = 1;
  for (int s = a; s < b; ++s) {
    const double c = sc[idx[s]];
    double f = (nu[s] == 1.0) ? c : (nu[s] == 2.0 ? c * c : pow(c, nu[s]));
    phi = first ? f : phi * f;
    first = 0;
  }
  return phi;
}

static double refc_pow(double refC, double refCinv, double dn)
{ /* factor multiplying k_f*exp(-dG) in the reverse rate: refC^(-dn) */
  if (dn == 0.0) return 1.0;
  if (dn == -1.0) return refC;
  if (dn == 1.0) return refCinv;
  return dn < 0 ? pow(refC, -dn) : pow(refCinv, dn);
}

/* productionRate :3650-4902 -- sc in mol/m^3, wdot in mol/m^3/s */
void orc_production_rate(const plm_mech_tables* m, double* wdot, const double* sc, double T)
{
  const double tc[5] = {log(T), T, T * T, T * T * T, T * T * T * T};
  const double invT = 1.0 / tc[1];
  const double refC = 101325 / 8.31446 * invT;
  const double refCinv = 1 / refC;
  double g_RT[ORC_MAXS];
  double mixture = 0.0;
  for (int i = 0; i < m->ns; ++i) wdot[i] = 0.0;
  for (int i = 0; i < m->ns; ++i) mixture += sc[i];
  orc_g_RT(m, g_RT, tc);

  for (int r = 0; r < m->nr; ++r) {
    const double b = m->b[r], EaR = m->EaR[r];
    double k_f = m->A[r];
    if (b != 0.0 && EaR != 0.0) k_f = k_f * exp(b * tc[0] - EaR * invT);
    else if (b != 0.0) k_f = k_f * exp(b * tc[0]);
    else if (EaR != 0.0) k_f = k_f * exp(-EaR * invT);
    double Corr = 1.0;
    const int type = m->type[r];
    if (type != 0) {
      Corr = mixture;
      for (int s = m->eoff[r]; s < m->eoff[r + 1]; ++s) {
        const double em1 = m->eval[s] - 1.0;
        Corr += (em1 == 1.0) ? sc[m->eidx[s]] : em1 * sc[m->eidx[s]];
      }
    }
    if (type >= 2) {
      const double redP = Corr / k_f * m->A0[r] * exp(m->b0[r] * tc[0] - m->Ea0R[r] * invT);
      const double F = redP / (1.0 + redP);
      if (type == 3) {
        const double* tr = m->troe + 4 * r;
        const double logPred = log10(redP);
        double fc = 0.0;
        if (fabs(tr[1]) > 1.0e-100 && (1.0 - tr[0]) != 0) fc = (1.0 - tr[0]) * exp(-tc[1] * (1 / tr[1]));
        if (fabs(tr[2]) > 1.0e-100 && tr[0] != 0) fc = fc + tr[0] * exp(-tc[1] * (1 / tr[2]));
        if (m->troe_np[r] == 4) fc = fc + exp(-tr[3] * invT);
        else fc = fc + 0.0;
        const double logFcent = log10(fc);
        const double troe_c = -0.4 - 0.67 * logFcent;
        const double troe_n = 0.75 - 1.27 * logFcent;
        const double troe = (troe_c + logPred) / (troe_n - 0.14 * (troe_c + logPred));
        const double F_troe = pow(10, logFcent / (1.0 + troe * troe));
        Corr = F * F_troe;
      } else {
        Corr = F;
      }
    }
    const double phi_f = phase(sc, m->ridx, m->rval, m->roff[r], m->roff[r + 1]);
    const double qf = (type != 0) ? Corr * k_f * phi_f : k_f * phi_f;
    double qr = 0.0;
    if (m->rev[r]) {
      double dG = 0.0; /* sum nu' g - sum nu'' g, species order */
      int first = 1;
      for (int s = m->noff[r]; s < m->noff[r + 1]; ++s) {
        const double c = -m->nval[s];
        const double term = (c == 1.0 || c == -1.0) ? g_RT[m->nidx[s]] : fabs(c) * g_RT[m->nidx[s]];
        if (first) dG = (c > 0) ? term : -term;
        else dG = (c > 0) ? dG + term : dG - term;
        first = 0;
      }
      const double phi_r = phase(sc, m->pidx, m->pval, m->poff[r], m->poff[r + 1]);
      double kr = (type != 0) ? Corr * k_f * exp(-(dG)) : k_f * exp(-(dG));
      if (m->dn[r] != 0.0) kr = kr * refc_pow(refC, refCinv, m->dn[r]);
      qr = kr * phi_r;
    }
    const double qdot = qf - qr;
    for (int s = m->noff[r]; s < m->noff[r + 1]; ++s) {
      const double c = m->nval[s];
      if (c == 1.0) wdot[m->nidx[s]] += qdot;
      else if (c == -1.0) wdot[m->nidx[s]] -= qdot;
      else if (c > 0) wdot[m->nidx[s]] += c * qdot;
      else wdot[m->nidx[s]] -= (-c) * qdot;
    }
  }
}

/* CKWC :4905-4923 -- C mol/cm^3 in, wdot mol/cm^3/s out */
void orc_ckwc(const plm_mech_tables* m, double T, const double* C, double* wdot)
{
  double c[ORC_MAXS];
  for (int i = 0; i < m->ns; ++i) c[i] = C[i] * 1.0e6;
  orc_production_rate(m, wdot, c, T);
  for (int i = 0; i < m->ns; ++i) wdot[i] *= 1.0e-6;
}

/* d(phi)/d(sc[j]) for a side given as CSR slice; 0 if j absent */
static int dphase(const double* sc, const int* idx, const double* nu, int a, int b, int j, double* out)
{
  int found = 0;
  double d = 1.0;
  for (int s = a; s < b; ++s) {
    const double c = sc[idx[s]];
    if (idx[s] == j) {
      found = 1;
      if (nu[s] == 1.0) d *= 1.0;
      else if (nu[s] == 2.0) d *= 2.0 * c;
      else d *= nu[s] * pow(c, nu[s] - 1.0);
    } else {
      d *= (nu[s] == 1.0) ? c : pow(c, nu[s]);
    }
  }
  *out = d;
  return found;
}

/* aJacobian :9696-14742 with consP = 1 (h-reactor; ReactorCvodeJacobian.cpp:129-131).
 * J column-major n x n, n = ns+1, J[col*n+row]; sc in mol/m^3. */
void orc_ajacobian(const plm_mech_tables* m, double* J, const double* sc, double T, int consP)
{
  const int ns = m->ns, n = ns + 1;
  double wdot[ORC_MAXS], g_RT[ORC_MAXS], h_RT[ORC_MAXS];
  for (int i = 0; i < n * n; ++i) J[i] = 0.0;
  for (int k = 0; k < ns; ++k) wdot[k] = 0.0;
  const double tc[5] = {log(T), T, T * T, T * T * T, T * T * T * T};
  const double invT = 1.0 / tc[1];
  const double invT2 = invT * invT;
  const double refC = 101325 / 8.31446 / T;
  const double refCinv = 1.0 / refC;
  double mixture = 0.0;
  for (int k = 0; k < ns; ++k) mixture += sc[k];
  orc_g_RT(m, g_RT, tc);
  orc_h_RT(m, h_RT, tc);
  const double ln10 = log(10.0);
  const double log10e = 1.0 / log(10.0);

  double dqdc_keep[ORC_MAXS];
  for (int j = 0; j < ns; ++j) dqdc_keep[j] = 0.0;
  for (int r = 0; r < m->nr; ++r) {
    const int type = m->type[r];
    double alpha = 0.0;
    if (type != 0) {
      alpha = mixture;
      for (int s = m->eoff[r]; s < m->eoff[r + 1]; ++s) alpha += (m->eval[s] - 1) * sc[m->eidx[s]];
    }
    const double phi_f = phase(sc, m->ridx, m->rval, m->roff[r], m->roff[r + 1]);
    double k_f = m->A[r] * exp(m->b[r] * tc[0] - m->EaR[r] * invT);
    const double dlnkfdT = m->b[r] * invT + m->EaR[r] * invT2;
    double Pr = 0, fPr = 0, F = 1.0, dlogfPrdT = 0, dlogFdT = 0, dlogFdlogPr = 0;
    if (type >= 2) {
      const double k_0 = m->A0[r] * exp(m->b0[r] * tc[0] - m->Ea0R[r] * invT);
      Pr = alpha / k_f * k_0;
      fPr = Pr / (1.0 + Pr);
      const double dlnk0dT = m->b0[r] * invT + m->Ea0R[r] * invT2;
      const double dlogPrdT = log10e * (dlnk0dT - dlnkfdT);
      dlogfPrdT = dlogPrdT / (1.0 + Pr);
      if (type == 3) {
        const double* tr = m->troe + 4 * r;
        const double logPr = log10(Pr);
        const double Fcent1 = (fabs(tr[1]) > 1.e-100) ? (1. - tr[0]) * exp(-T / tr[1]) : 0.;
        const double Fcent2 = (fabs(tr[2]) > 1.e-100) ? tr[0] * exp(-T / tr[2]) : 0.;
        const double Fcent3 = (m->troe_np[r] == 4) ? exp(-tr[3] * invT) : 0.;
        const double Fcent = Fcent1 + Fcent2 + Fcent3;
        const double logFcent = log10(Fcent);
        const double troe_c = -.4 - .67 * logFcent;
        const double troe_n = .75 - 1.27 * logFcent;
        const double troePr_den = 1.0 / (troe_n - .14 * (troe_c + logPr));
        const double troePr = (troe_c + logPr) * troePr_den;
        const double troe = 1.0 / (1.0 + troePr * troePr);
        F = pow(10.0, logFcent * troe);
        double dlogFcentdT = -Fcent1 / tr[1] - Fcent2 / tr[2];
        if (m->troe_np[r] == 4) dlogFcentdT += Fcent3 * tr[3] * invT2;
        dlogFcentdT = log10e / Fcent * dlogFcentdT;
        const double dlogFdcn_fac = 2.0 * logFcent * troe * troe * troePr * troePr_den;
        const double dlogFdc = -troe_n * dlogFdcn_fac * troePr_den;
        const double dlogFdn = dlogFdcn_fac * troePr;
        dlogFdlogPr = dlogFdc;
        dlogFdT = dlogFcentdT * (troe - 0.67 * dlogFdc - 1.27 * dlogFdn) + dlogFdlogPr * dlogPrdT;
      }
    }
    double phi_r = 0.0, k_r = 0.0, dkrdT = 0.0;
    if (m->rev[r]) {
      phi_r = phase(sc, m->pidx, m->pval, m->poff[r], m->poff[r + 1]);
      double dG = 0.0;
      for (int s = m->noff[r]; s < m->noff[r + 1]; ++s) dG += -m->nval[s] * g_RT[m->nidx[s]];
      double Kc = exp(dG);
      const double dn = m->dn[r];
      if (dn != 0.0) Kc = ((dn == -1.0) ? refCinv : (dn == 1.0 ? refC : pow(refC, dn))) * Kc;
      k_r = k_f / Kc;
      double hr = 0.0, hp = 0.0;
      for (int s = m->roff[r]; s < m->roff[r + 1]; ++s) hr += m->rval[s] * h_RT[m->ridx[s]];
      for (int s = m->poff[r]; s < m->poff[r + 1]; ++s) hp += m->pval[s] * h_RT[m->pidx[s]];
      const double dlnKcdT = invT * (-(hr) + (hp) + (-dn));
      dkrdT = (dlnkfdT - dlnKcdT) * k_r;
    }
    double q, dqdT, q_nocor = 0.0, dcdc_fac = 0.0;
    if (type == 0) {
      q = k_f * phi_f - k_r * phi_r;
      dqdT = (dlnkfdT * k_f * phi_f - dkrdT * phi_r);
    } else if (type == 1) {
      q_nocor = k_f * phi_f - k_r * phi_r;
      q = alpha * q_nocor;
      dqdT = alpha * (dlnkfdT * k_f * phi_f - dkrdT * phi_r);
      k_f *= alpha;
      k_r *= alpha;
      dcdc_fac = q_nocor;
    } else {
      q_nocor = k_f * phi_f - k_r * phi_r;
      const double Corr = fPr * F;
      q = Corr * q_nocor;
      const double dlnCorrdT = ln10 * (dlogfPrdT + dlogFdT);
      dqdT = Corr * (dlnkfdT * k_f * phi_f - dkrdT * phi_r) + dlnCorrdT * q;
      k_f *= Corr;
      k_r *= Corr;
      dcdc_fac = q / alpha * (1.0 / (Pr + 1.0) + dlogFdlogPr);
    }
    for (int s = m->noff[r]; s < m->noff[r + 1]; ++s) wdot[m->nidx[s]] += m->nval[s] * q;
    /* columns: every species the rate depends on */
    for (int j = 0; j < ns; ++j) {
      double dqdci = 0.0, d;
      int dep = 0, zero_eff = 0;
      if (type != 0) {
        if (consP) {
          for (int s = m->eoff[r]; s < m->eoff[r + 1]; ++s)
            if (m->eidx[s] == j) {
              dqdci = (m->eval[s] - 1) * dcdc_fac;
              dep = 1;
            }
        } else { /* constant volume: every species is a third body (efficiency 1 unless listed), mechanism.H "else" branches */
          dqdci = dcdc_fac;
          for (int s = m->eoff[r]; s < m->eoff[r + 1]; ++s)
            if (m->eidx[s] == j) {
              dqdci = m->eval[s] * dcdc_fac;
              zero_eff = (m->eval[s] == 0.0);
            }
          dep = 1;
        }
      }
      int phase_dep = 0;
      if (dphase(sc, m->ridx, m->rval, m->roff[r], m->roff[r + 1], j, &d)) {
        dqdci += k_f * d;
        dep = phase_dep = 1;
      }
      if (m->rev[r] && dphase(sc, m->pidx, m->pval, m->poff[r], m->poff[r + 1], j, &d)) {
        dqdci -= k_r * d;
        dep = phase_dep = 1;
      }
      if (!dep) continue;
      if (type != 0 && !consP) {
        /* The generated reference keeps ONE dqdc[] array for the whole routine and, in these branches, does not assign
         * the entry of a species whose efficiency is 0 and that is neither reactant nor product (e.g. drm19
         * mechanism.H reaction 22, dqdc[5] / dqdc[12]) -- yet adds all ns entries to J: such an entry still holds
         * the value the previous third-body reaction left there. Reproduced, so that cJac matches entry by entry. */
        if (zero_eff && !phase_dep) dqdci = dqdc_keep[j];
        else dqdc_keep[j] = dqdci;
      }
      for (int s = m->noff[r]; s < m->noff[r + 1]; ++s) J[j * n + m->nidx[s]] += m->nval[s] * dqdci;
    }
    for (int s = m->noff[r]; s < m->noff[r + 1]; ++s) J[ns * n + m->nidx[s]] += m->nval[s] * dqdT;
  }

  /* closure :14705-14741 */
  double c_R[ORC_MAXS], dcRdT[ORC_MAXS], e_RT[ORC_MAXS];
  const double* eh_RT = h_RT;
  if (consP) {
    orc_cp_R(m, c_R, tc);
    orc_dcp_R(m, dcRdT, tc);
  } else {
    orc_cv_R(m, c_R, tc);
    orc_dcp_R(m, dcRdT, tc);
    orc_e_RT(m, e_RT, tc);
    eh_RT = e_RT;
  }
  double cmix = 0.0, ehmix = 0.0, dcmixdT = 0.0, dehmixdT = 0.0;
  for (int k = 0; k < ns; ++k) {
    cmix += c_R[k] * sc[k];
    dcmixdT += dcRdT[k] * sc[k];
    ehmix += eh_RT[k] * wdot[k];
    dehmixdT += invT * (c_R[k] - eh_RT[k]) * wdot[k] + eh_RT[k] * J[ns * n + k];
  }
  const double cmixinv = 1.0 / cmix;
  const double tmp1 = ehmix * cmixinv;
  const double tmp3 = cmixinv * T;
  const double tmp2 = tmp1 * tmp3;
  for (int k = 0; k < ns; ++k) {
    double dehmixdc = 0.0;
    for (int mm = 0; mm < ns; ++mm) dehmixdc += eh_RT[mm] * J[k * n + mm];
    J[k * n + ns] = tmp2 * c_R[k] - tmp3 * dehmixdc;
  }
  J[ns * n + ns] = -tmp1 + tmp2 * dcmixdT - tmp3 * dehmixdT;
}

/* DWDOT :14745-14765 + mass scaling of cJac (ReactorCvodeJacobian.cpp:101-148); y = [rhoY_k, T] */
void orc_cjac(const plm_mech_tables* m, const double* y, double* Jout) { orc_cjac_rt(m, 2, y, Jout); }
/* cJac with the reactor type: consP = (reactor_type != eint) (ReactorCvodeJacobian.cpp:109-118) */
void orc_cjac_rt(const plm_mech_tables* m, int reactor_type, const double* y, double* Jout)
{
  const int ns = m->ns, n = ns + 1;
  double rho = 0.0, massfrac[ORC_MAXS], C[ORC_MAXS], c[ORC_MAXS];
  double Jt[(ORC_MAXS + 1) * (ORC_MAXS + 1)];
  for (int i = 0; i < ns; ++i) rho = rho + y[i];
  const double temp = y[ns];
  for (int i = 0; i < ns; ++i) massfrac[i] = y[i] / rho;
  for (int i = 0; i < ns; ++i) C[i] = rho * massfrac[i] * (1.0 / m->mw[i]); /* CKYTCR :2403 */
  for (int k = 0; k < ns; ++k) c[k] = 1.e6 * C[k];
  orc_ajacobian(m, Jt, c, temp, reactor_type == 1 ? 0 : 1);
  for (int k = 0; k < ns; ++k) {
    Jt[ns * n + k] *= 1.e-6;
    Jt[k * n + ns] *= 1.e6;
  }
  for (int i = 0; i < ns; ++i) {
    double* col = Jout + i * n;
    for (int k = 0; k < ns; ++k) col[k] = Jt[i * n + k] * m->mw[k] / m->mw[i];
    col[ns] = Jt[i * n + ns] / m->mw[i];
  }
  double* col = Jout + ns * n;
  for (int i = 0; i < ns; ++i) col[i] = Jt[ns * n + i] * m->mw[i];
  /* The reference never writes the (T,T) entry (ReactorCvodeJacobian.cpp:143-148, "Never read"): the
   * matrix was zeroed by cvLsLinSys (cvode_ls.c:1311), so dTdot/dT = 0 in the Newton matrix. Mirrored. */
  col[ns] = 0.0;
}

/* fKernelSpec, h-reactor branch (ReactorUtils.H:274-332) with Fuego wrappers RHY2T :128, RTY2Hi :293,
 * RTY2Cp :36, RTY2WDOT :168. y[ns+1] = [rhoY, T guess]. */
void orc_rhs(const plm_mech_tables* m, double dt_save, const double* y, double* ydot, double rhoe_init,
             double rhoesrc_ext, const double* rYs)
{
  orc_rhs_rt(m, 2, dt_save, y, ydot, rhoe_init, rhoesrc_ext, rYs);
}
/* reactor_type 1 = eint (REY2T, RTY2Ei, RTY2Cv), 2 = enthalpy (RHY2T, RTY2Hi, RTY2Cp): ReactorUtils.H:300-316 */
void orc_rhs_rt(const plm_mech_tables* m, int reactor_type, double dt_save, const double* y, double* ydot, double rhoe_init,
                double rhoesrc_ext, const double* rYs)
{
  const int ns = m->ns;
  double massfrac[ORC_MAXS], ei[ORC_MAXS], cdots[ORC_MAXS], C[ORC_MAXS];
  double rho = 0.0;
  for (int k = 0; k < ns; ++k) {
    massfrac[k] = y[k];
    rho += massfrac[k];
  }
  const double rho_inv = 1.0 / rho;
  for (int k = 0; k < ns; ++k) massfrac[k] *= rho_inv;
  const double nrg = (rhoe_init + rhoesrc_ext * dt_save) * rho_inv;
  double temp = y[ns];
  double cp;
  if (reactor_type == 1) {
    orc_get_T_given_eY(m, nrg, massfrac, &temp);
    orc_ums(m, temp, ei);
    cp = orc_cvbs(m, temp, massfrac);
  } else {
    orc_get_T_given_hY(m, nrg, massfrac, &temp);
    orc_hms(m, temp, ei);
    cp = orc_cpbs(m, temp, massfrac);
  }
  for (int i = 0; i < ns; ++i) C[i] = rho * massfrac[i] * (1.0 / m->mw[i]);
  orc_ckwc(m, temp, C, cdots);
  for (int k = 0; k < ns; ++k) cdots[k] *= m->mw[k];
  double rhoesrc = rhoesrc_ext;
  for (int k = 0; k < ns; ++k) {
    const double cdot_rYs = cdots[k] + rYs[k];
    ydot[k] = cdot_rYs;
    rhoesrc -= cdot_rYs * ei[k];
  }
  ydot[ns] = rhoesrc * (rho_inv / cp);
}

/* reactionRateRhoY for one cell (reference Source/PeleLM_K.H:57-101): MKS rhoY[ns], rhoH, T guess -> rhoYdot[ns] MKS.
 * HY2T -> GET_T_GIVEN_HY (Fuego.H:120-126); RTY2WDOT (Fuego.H:168-183) = CKYTCR, CKWC, x mw. */
void orc_reaction_rate_rhoY(const plm_mech_tables* m, const double* rhoY, double rhoH, double T, int mask, double* rhoYdot)
{
  const int ns = m->ns;
  if (mask < 0) {
    for (int n = 0; n < ns; n++) rhoYdot[n] = 0.0;
    return;
  }
  double rho = 0.0;
  for (int n = 0; n < ns; n++) rho += rhoY[n];
  const double rhoinv = 1.0 / rho;
  double y[ORC_MAXS], C[ORC_MAXS], wdot[ORC_MAXS];
  for (int n = 0; n < ns; n++) y[n] = rhoY[n] * rhoinv;
  double Tloc = T;
  const double H = rhoH * rhoinv * 1.0e4;
  orc_get_T_given_hY(m, H, y, &Tloc);
  rho = rho * 0.001;
  for (int i = 0; i < ns; ++i) C[i] = rho * y[i] * (1.0 / m->mw[i]);
  orc_ckwc(m, Tloc, C, wdot);
  for (int n = 0; n < ns; n++) wdot[n] *= m->mw[n];
  for (int n = 0; n < ns; n++) rhoYdot[n] = wdot[n] * 1000.0;
}
