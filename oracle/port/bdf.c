/* TEST INFRASTRUCTURE -- per-cell restatement of SUNDIALS CVODE 6.2.0 (BDF, Newton, dense LU) exactly
 * as PelePhysics' CPU ReactorCvode configures and drives it (ReactorCvode.cpp:9-295, :1372-1453):
 * CVodeReInit + CVode(tout, CV_NORMAL) per cell, no tstop, no constraints, no root finding,
 * stability-limit detection off, maxnef 100, mxstep 10000, msbj 100, maxcor = ode.max_nls_iters.
 * Pinned against the compiled reference in tests/test_oracle_port_vs_ref.py.
 *
 * Source map (SUNDIALS src/): CVode cvode/cvode.c:990-1476, cvHin :1945-2048, cvStep :2143-2249,
 * cvAdjustParams/Order/BDF :2265-2451, cvRescale :2457, cvPredict :2485, cvSetBDF/cvSetTqBDF :2691-2779,
 * cvNls :2781, cvHandleNFlag :2954, cvRestore :3008, cvDoErrorTest :3048, cvCompleteStep :3162,
 * cvPrepareNextStep :3218, cvSetEta :3262, cvComputeEtaqm1/qp1 :3299/:3318, cvChooseEta :3347,
 * CVodeGetDky :1490; Newton loop sunnonlinsol/newton/sunnonlinsol_newton.c:187-335;
 * cvNlsLSetup/ConvTest/Residual cvode/cvode_nls.c:251-388; cvLsSetup/LinSys/Solve cvode/cvode_ls.c:1270-1720;
 * denseGETRF/GETRS sundials/sundials_dense.c:135-232; WRMS norm nvector/serial/nvector_serial.c:638-660.
 */
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <string.h>

#define ZERO 0.0
#define ONE 1.0
#define TWO 2.0
/* cvode_impl.h:57-130, cvode.c:111-118, cvode_nls.c:29-31, cvode_ls_impl.h:40-42 */
#define ETA_MIN_FX 0.0
#define ETA_MAX_FX 1.5
#define ETA_MAX_FS 10000.0
#define ETA_MAX_ES 10.0
#define ETA_MAX_GS 10.0
#define ETA_MIN 0.1
#define ETA_MAX_EF 0.2
#define ETA_MIN_EF 0.1
#define ETA_CF 0.25
#define SMALL_NST 10
#define SMALL_NEF 2
#define ADDON 0.000001
#define BIAS1 6.0
#define BIAS2 6.0
#define BIAS3 10.0
#define LONG_WAIT 10
#define MXNCF 10
#define MXNEF1 3
#define HLB_FACTOR 100.0
#define HUB_FACTOR 0.1
#define H_BIAS 0.5
#define MAX_ITERS 4
#define CORTES 0.1
#define CRDOWN 0.3
#define RDIV 2.0
#define MSBP 20
#define DGMAX_LSETUP 0.3
#define DGMAX_JBAD 0.2

#define CV_TOO_MUCH_WORK (-1)
#define CV_TOO_MUCH_ACC (-2)
#define CV_ERR_FAILURE (-3)
#define CV_CONV_FAILURE (-4)
#define CV_TOO_CLOSE (-27)

enum { DO_ERROR_TEST = 2, PREDICT_AGAIN = 3, TRY_AGAIN = 5 };
enum { FIRST_CALL = 6, PREV_CONV_FAIL = 7, PREV_ERR_FAIL = 8 };
enum { CV_NO_FAILURES = 0, CV_FAIL_BAD_J = 1, CV_FAIL_OTHER = 2 };
enum { NLS_SUCCESS = 0, NLS_CONTINUE = 901, NLS_CONV_RECVR = 902 };

typedef struct {
  const plm_mech_tables* m;
  int n, qmax, maxcor, maxnef, mxstep, msbj, ajac;
  int rt; /* ode.reactor_type */
  double rtol, atolv[ORC_MAXN];
  double rhoe_init, rhoesrc;
  const double* rYsrc;
  double zn[ORC_QMAX + 1][ORC_MAXN];
  double ewt[ORC_MAXN], y[ORC_MAXN], acor[ORC_MAXN], tempv[ORC_MAXN], ftemp[ORC_MAXN], delta[ORC_MAXN];
  double A[ORC_MAXN * ORC_MAXN], savedJ[ORC_MAXN * ORC_MAXN];
  int piv[ORC_MAXN];
  double tn, h, hprime, hscale, eta, etamax, hu;
  int q, qprime, L, qwait, indx_acor;
  double tau[ORC_QMAX + 3], tq[6], l[ORC_QMAX + 3];
  double rl1, gamma, gammap, gamrat, crate, delp, acnrm, saved_tq5, etaq, etaqm1, etaqp1, nlscoef;
  long nst, nfe, ncfn, netf, nni, nsetups, nje, nstlp, nstlj;
  int jcur, convfail;
} cvmem;

static double rpowR(double base, double e) { return base <= 0.0 ? 0.0 : pow(base, e); }
static double rpowI(double base, int e)
{
  double prod = 1.0;
  int expt = e < 0 ? -e : e;
  for (int i = 1; i <= expt; i++) prod *= base;
  if (e < 0) prod = 1.0 / prod;
  return prod;
}
static double wrms(const double* x, const double* w, int n)
{
  double sum = 0.0;
  for (int i = 0; i < n; i++) {
    const double p = x[i] * w[i];
    sum += p * p;
  }
  return sqrt(sum / n);
}
static int ewtset(cvmem* cv, const double* ycur, double* w)
{
  for (int i = 0; i < cv->n; i++) {
    const double t = cv->rtol * fabs(ycur[i]) + cv->atolv[i];
    if (t <= 0.0) return -1;
    w[i] = 1.0 / t;
  }
  return 0;
}
static void f(cvmem* cv, double t, const double* y, double* ydot)
{
  orc_rhs_rt(cv->m, cv->rt, t, y, ydot, cv->rhoe_init, cv->rhoesrc, cv->rYsrc);
  cv->nfe++;
}

/* ---- dense LU, column-major a[j*n+i] ---- */
static int getrf(double* a, int n, int* p)
{
  for (int k = 0; k < n; k++) {
    double* col_k = a + k * n;
    int l = k;
    for (int i = k + 1; i < n; i++)
      if (fabs(col_k[i]) > fabs(col_k[l])) l = i;
    p[k] = l;
    if (col_k[l] == ZERO) return k + 1;
    if (l != k) {
      for (int i = 0; i < n; i++) {
        const double temp = a[i * n + l];
        a[i * n + l] = a[i * n + k];
        a[i * n + k] = temp;
      }
    }
    const double mult = ONE / col_k[k];
    for (int i = k + 1; i < n; i++) col_k[i] *= mult;
    for (int j = k + 1; j < n; j++) {
      double* col_j = a + j * n;
      const double a_kj = col_j[k];
      if (a_kj != ZERO)
        for (int i = k + 1; i < n; i++) col_j[i] -= a_kj * col_k[i];
    }
  }
  return 0;
}
static void getrs(const double* a, int n, const int* p, double* b)
{
  for (int k = 0; k < n; k++) {
    const int pk = p[k];
    if (pk != k) {
      const double tmp = b[k];
      b[k] = b[pk];
      b[pk] = tmp;
    }
  }
  for (int k = 0; k < n - 1; k++) {
    const double* col_k = a + k * n;
    for (int i = k + 1; i < n; i++) b[i] -= col_k[i] * b[k];
  }
  for (int k = n - 1; k > 0; k--) {
    const double* col_k = a + k * n;
    b[k] /= col_k[k];
    for (int i = 0; i < k; i++) b[i] -= col_k[i] * b[k];
  }
  b[0] /= a[0];
}

/* cvLsDenseDQJac cvode_ls.c:1034-1107 (dense_direct); fy = f(t,y) */
static void dq_jac(cvmem* cv, double t, double* y, const double* fy, double* J)
{
  const int n = cv->n;
  const double uround = DBL_EPSILON;
  const double srur = (uround > 0.0) ? sqrt(uround) : 0.0;
  const double fnorm = wrms(fy, cv->ewt, n);
  const double minInc = (fnorm != ZERO) ? (1000.0 * fabs(cv->h) * uround * n * fnorm) : ONE;
  double ft[ORC_MAXN];
  for (int j = 0; j < n; j++) {
    const double yjsaved = y[j];
    double inc = srur * fabs(yjsaved);
    if (minInc / cv->ewt[j] > inc) inc = minInc / cv->ewt[j];
    y[j] += inc;
    orc_rhs_rt(cv->m, cv->rt, t, y, ft, cv->rhoe_init, cv->rhoesrc, cv->rYsrc);
    cv->nfe++; /* counted as nfeLS by CVODE; PelePhysics adds both into FC */
    y[j] = yjsaved;
    const double inc_inv = ONE / inc;
    for (int i = 0; i < n; i++) J[j * n + i] = inc_inv * ft[i] - inc_inv * fy[i];
  }
}

/* cvNlsLSetup + cvLsSetup + cvLsLinSys + SUNLinSolSetup_Dense */
static int lsetup(cvmem* cv)
{
  const int n = cv->n;
  const double dgamma = fabs((cv->gamma / cv->gammap) - ONE);
  const int jbad = (cv->nst == 0) || (cv->nst >= cv->nstlj + cv->msbj) ||
                   ((cv->convfail == CV_FAIL_BAD_J) && (dgamma < DGMAX_JBAD)) || (cv->convfail == CV_FAIL_OTHER);
  if (!jbad) {
    cv->jcur = 0;
    memcpy(cv->A, cv->savedJ, sizeof(double) * n * n);
  } else {
    cv->jcur = 1;
    if (cv->ajac) orc_cjac_rt(cv->m, cv->rt, cv->y, cv->A);
    else dq_jac(cv, cv->tn, cv->y, cv->ftemp, cv->A);
    memcpy(cv->savedJ, cv->A, sizeof(double) * n * n);
  }
  const double c = -cv->gamma;
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++) {
      cv->A[j * n + i] *= c;
      if (i == j) cv->A[j * n + i] += ONE;
    }
  if (cv->jcur) {
    cv->nje++;
    cv->nstlj = cv->nst;
  }
  const int ier = getrf(cv->A, n, cv->piv);
  cv->nsetups++;
  cv->gamrat = ONE;
  cv->gammap = cv->gamma;
  cv->crate = ONE;
  cv->nstlp = cv->nst;
  return ier > 0 ? NLS_CONV_RECVR : NLS_SUCCESS;
}

/* cvNlsResidual */
static void residual(cvmem* cv, double* res)
{
  const int n = cv->n;
  for (int i = 0; i < n; i++) cv->y[i] = cv->zn[0][i] + cv->acor[i];
  f(cv, cv->tn, cv->y, cv->ftemp);
  for (int i = 0; i < n; i++) res[i] = cv->rl1 * cv->zn[1][i] + cv->acor[i];
  for (int i = 0; i < n; i++) res[i] = -cv->gamma * cv->ftemp[i] + res[i];
}

/* SUNNonlinSolSolve_Newton with cvNlsConvTest / cvLsSolve */
static int newton(cvmem* cv, int callSetup)
{
  const int n = cv->n;
  int jbad = 0, retval, nls_jcur = 0;
  double* delta = cv->delta;
  for (;;) {
    residual(cv, delta);
    retval = NLS_SUCCESS;
    if (callSetup) {
      if (jbad) cv->convfail = CV_FAIL_BAD_J;
      retval = lsetup(cv);
      nls_jcur = cv->jcur;
      if (retval != NLS_SUCCESS) return retval; /* LSetup failure leaves the setup loop directly */
    }
    if (retval == NLS_SUCCESS) {
      int m = 0;
      for (;;) {
        cv->nni++;
        for (int i = 0; i < n; i++) delta[i] = -ONE * delta[i];
        getrs(cv->A, n, cv->piv, delta);
        if (cv->gamrat != ONE) {
          const double s = TWO / (ONE + cv->gamrat);
          for (int i = 0; i < n; i++) delta[i] *= s;
        }
        for (int i = 0; i < n; i++) cv->acor[i] = cv->acor[i] + delta[i];
        /* cvNlsConvTest */
        const double del = wrms(delta, cv->ewt, n);
        if (m > 0) {
          const double r = del / cv->delp;
          cv->crate = (CRDOWN * cv->crate > r) ? CRDOWN * cv->crate : r;
        }
        const double dcon = del * ((ONE < cv->crate) ? ONE : cv->crate) / cv->tq[4];
        if (dcon <= ONE) {
          cv->acnrm = (m == 0) ? del : wrms(cv->acor, cv->ewt, n);
          return NLS_SUCCESS;
        }
        if ((m >= 1) && (del > RDIV * cv->delp)) {
          retval = NLS_CONV_RECVR;
          break;
        }
        cv->delp = del;
        m++;
        if (m >= cv->maxcor) {
          retval = NLS_CONV_RECVR;
          break;
        }
        residual(cv, delta);
      }
    }
    if ((retval > 0) && !nls_jcur) {
      callSetup = 1;
      jbad = 1;
      for (int i = 0; i < n; i++) cv->acor[i] = ZERO;
      continue;
    }
    break;
  }
  return retval;
}

static void rescale(cvmem* cv)
{
  double fac = cv->eta;
  for (int j = 1; j <= cv->q; j++) {
    for (int i = 0; i < cv->n; i++) cv->zn[j][i] *= fac;
    fac = cv->eta * fac;
  }
  cv->h = cv->hscale * cv->eta;
  cv->hscale = cv->h;
}

static void increase_bdf(cvmem* cv)
{
  double alpha0, alpha1, prod, xi, xiold, hsum, A1;
  for (int i = 0; i <= cv->qmax; i++) cv->l[i] = ZERO;
  cv->l[2] = alpha1 = prod = xiold = ONE;
  alpha0 = -ONE;
  hsum = cv->hscale;
  if (cv->q > 1) {
    for (int j = 1; j < cv->q; j++) {
      hsum += cv->tau[j + 1];
      xi = hsum / cv->hscale;
      prod *= xi;
      alpha0 -= ONE / (j + 1);
      alpha1 += ONE / xi;
      for (int i = j + 2; i >= 2; i--) cv->l[i] = cv->l[i] * xiold + cv->l[i - 1];
      xiold = xi;
    }
  }
  A1 = (-alpha0 - alpha1) / prod;
  for (int i = 0; i < cv->n; i++) cv->zn[cv->L][i] = A1 * cv->zn[cv->indx_acor][i];
  if (cv->q > 1)
    for (int j = 2; j <= cv->q; j++)
      for (int i = 0; i < cv->n; i++) cv->zn[j][i] = cv->l[j] * cv->zn[cv->L][i] + cv->zn[j][i];
}
static void decrease_bdf(cvmem* cv)
{
  double hsum, xi;
  for (int i = 0; i <= cv->qmax; i++) cv->l[i] = ZERO;
  cv->l[2] = ONE;
  hsum = ZERO;
  for (int j = 1; j <= cv->q - 2; j++) {
    hsum += cv->tau[j];
    xi = hsum / cv->hscale;
    for (int i = j + 2; i >= 2; i--) cv->l[i] = cv->l[i] * xi + cv->l[i - 1];
  }
  if (cv->q > 2)
    for (int j = 2; j < cv->q; j++)
      for (int i = 0; i < cv->n; i++) cv->zn[j][i] = -cv->l[j] * cv->zn[cv->q][i] + cv->zn[j][i];
}
static void adjust_order(cvmem* cv, int deltaq)
{
  if ((cv->q == 2) && (deltaq != 1)) return;
  if (deltaq == 1) increase_bdf(cv);
  else if (deltaq == -1) decrease_bdf(cv);
}
static void adjust_params(cvmem* cv)
{
  if (cv->qprime != cv->q) {
    adjust_order(cv, cv->qprime - cv->q);
    cv->q = cv->qprime;
    cv->L = cv->q + 1;
    cv->qwait = cv->L;
  }
  rescale(cv);
}
static void predict(cvmem* cv)
{
  cv->tn += cv->h;
  for (int k = 1; k <= cv->q; k++)
    for (int j = cv->q; j >= k; j--)
      for (int i = 0; i < cv->n; i++) cv->zn[j - 1][i] = cv->zn[j - 1][i] + cv->zn[j][i];
}
static void restore(cvmem* cv, double saved_t)
{
  cv->tn = saved_t;
  for (int k = 1; k <= cv->q; k++)
    for (int j = cv->q; j >= k; j--)
      for (int i = 0; i < cv->n; i++) cv->zn[j - 1][i] = cv->zn[j - 1][i] - cv->zn[j][i];
}
static void set_bdf(cvmem* cv)
{
  double alpha0, alpha0_hat, xi_inv, xistar_inv, hsum;
  const int q = cv->q;
  cv->l[0] = cv->l[1] = xi_inv = xistar_inv = ONE;
  for (int i = 2; i <= q; i++) cv->l[i] = ZERO;
  alpha0 = alpha0_hat = -ONE;
  hsum = cv->h;
  if (q > 1) {
    for (int j = 2; j < q; j++) {
      hsum += cv->tau[j - 1];
      xi_inv = cv->h / hsum;
      alpha0 -= ONE / j;
      for (int i = j; i >= 1; i--) cv->l[i] += cv->l[i - 1] * xi_inv;
    }
    alpha0 -= ONE / q;
    xistar_inv = -cv->l[1] - alpha0;
    hsum += cv->tau[q - 1];
    xi_inv = cv->h / hsum;
    alpha0_hat = -cv->l[1] - xi_inv;
    for (int i = q; i >= 1; i--) cv->l[i] += cv->l[i - 1] * xistar_inv;
  }
  /* cvSetTqBDF */
  double A1, A2, A3, A4, A5, A6, C, Cpinv, Cppinv;
  A1 = ONE - alpha0_hat + alpha0;
  A2 = ONE + q * A1;
  cv->tq[2] = fabs(A1 / (alpha0 * A2));
  cv->tq[5] = fabs(A2 * xistar_inv / (cv->l[q] * xi_inv));
  if (cv->qwait == 1) {
    if (q > 1) {
      C = xistar_inv / cv->l[q];
      A3 = alpha0 + ONE / q;
      A4 = alpha0_hat + xi_inv;
      Cpinv = (ONE - A4 + A3) / A3;
      cv->tq[1] = fabs(C * Cpinv);
    } else
      cv->tq[1] = ONE;
    hsum += cv->tau[q];
    xi_inv = cv->h / hsum;
    A5 = alpha0 - (ONE / (q + 1));
    A6 = alpha0_hat - xi_inv;
    Cppinv = (ONE - A6 + A5) / A2;
    cv->tq[3] = fabs(Cppinv / (xi_inv * (q + 2) * A5));
  }
  cv->tq[4] = cv->nlscoef / cv->tq[2];
  /* cvSet tail */
  cv->rl1 = ONE / cv->l[1];
  cv->gamma = cv->h * cv->rl1;
  if (cv->nst == 0) cv->gammap = cv->gamma;
  cv->gamrat = (cv->nst > 0) ? cv->gamma / cv->gammap : ONE;
}

static int nls(cvmem* cv, int nflag)
{
  cv->convfail = ((nflag == FIRST_CALL) || (nflag == PREV_ERR_FAIL)) ? CV_NO_FAILURES : CV_FAIL_OTHER;
  const int callSetup = (nflag == PREV_CONV_FAIL) || (nflag == PREV_ERR_FAIL) || (cv->nst == 0) ||
                        (cv->nst >= cv->nstlp + MSBP) || (fabs(cv->gamrat - ONE) > DGMAX_LSETUP);
  for (int i = 0; i < cv->n; i++) cv->acor[i] = ZERO;
  const int flag = newton(cv, callSetup);
  if (flag != NLS_SUCCESS) return flag;
  for (int i = 0; i < cv->n; i++) cv->y[i] = cv->zn[0][i] + cv->acor[i];
  cv->jcur = 0;
  return 0;
}

static int handle_nflag(cvmem* cv, int* nflag, double saved_t, int* ncf)
{
  if (*nflag == 0) return DO_ERROR_TEST;
  cv->ncfn++;
  restore(cv, saved_t);
  (*ncf)++;
  cv->etamax = ONE;
  if (*ncf == MXNCF) return CV_CONV_FAILURE; /* hmin = 0 */
  cv->eta = ETA_CF;
  *nflag = PREV_CONV_FAIL;
  rescale(cv);
  return PREDICT_AGAIN;
}

static int do_error_test(cvmem* cv, int* nflag, double saved_t, int* nef, double* dsmPtr)
{
  const double dsm = cv->acnrm * cv->tq[2];
  *dsmPtr = dsm;
  if (dsm <= ONE) return 0;
  (*nef)++;
  cv->netf++;
  *nflag = PREV_ERR_FAIL;
  restore(cv, saved_t);
  if (*nef == cv->maxnef) return CV_ERR_FAILURE;
  cv->etamax = ONE;
  if (*nef <= MXNEF1) {
    cv->eta = ONE / (rpowR(BIAS2 * dsm, ONE / cv->L) + ADDON);
    cv->eta = (ETA_MIN_EF > cv->eta) ? ETA_MIN_EF : cv->eta; /* hmin/|h| = 0 */
    if (*nef >= SMALL_NEF) cv->eta = (cv->eta < ETA_MAX_EF) ? cv->eta : ETA_MAX_EF;
    rescale(cv);
    return TRY_AGAIN;
  }
  if (cv->q > 1) {
    cv->eta = ETA_MIN_EF;
    adjust_order(cv, -1);
    cv->L = cv->q;
    cv->q--;
    cv->qwait = cv->L;
    rescale(cv);
    return TRY_AGAIN;
  }
  cv->eta = ETA_MIN_EF;
  cv->h *= cv->eta;
  cv->hscale = cv->h;
  cv->qwait = LONG_WAIT;
  f(cv, cv->tn, cv->zn[0], cv->tempv);
  for (int i = 0; i < cv->n; i++) cv->zn[1][i] = cv->h * cv->tempv[i];
  return TRY_AGAIN;
}

static void complete_step(cvmem* cv)
{
  cv->nst++;
  cv->hu = cv->h;
  for (int i = cv->q; i >= 2; i--) cv->tau[i] = cv->tau[i - 1];
  if ((cv->q == 1) && (cv->nst > 1)) cv->tau[2] = cv->tau[1];
  cv->tau[1] = cv->h;
  for (int j = 0; j <= cv->q; j++)
    for (int i = 0; i < cv->n; i++) cv->zn[j][i] = cv->l[j] * cv->acor[i] + cv->zn[j][i];
  cv->qwait--;
  if ((cv->qwait == 1) && (cv->q != cv->qmax)) {
    for (int i = 0; i < cv->n; i++) cv->zn[cv->qmax][i] = cv->acor[i];
    cv->saved_tq5 = cv->tq[5];
    cv->indx_acor = cv->qmax;
  }
}

static void set_eta(cvmem* cv)
{
  if ((cv->eta > ETA_MIN_FX) && (cv->eta < ETA_MAX_FX)) {
    cv->eta = ONE;
    cv->hprime = cv->h;
  } else {
    if (cv->eta >= ETA_MAX_FX) {
      cv->eta = (cv->eta < cv->etamax) ? cv->eta : cv->etamax;
      /* hmax_inv = 0: eta /= max(1, 0) */
    } else {
      cv->eta = (cv->eta > ETA_MIN) ? cv->eta : ETA_MIN;
    }
    cv->hprime = cv->h * cv->eta;
  }
}

static void prepare_next_step(cvmem* cv, double dsm)
{
  const int n = cv->n;
  if (cv->etamax == ONE) {
    cv->qwait = (cv->qwait > 2) ? cv->qwait : 2;
    cv->qprime = cv->q;
    cv->hprime = cv->h;
    cv->eta = ONE;
    return;
  }
  cv->etaq = ONE / (rpowR(BIAS2 * dsm, ONE / cv->L) + ADDON);
  if (cv->qwait != 0) {
    cv->eta = cv->etaq;
    cv->qprime = cv->q;
    set_eta(cv);
    return;
  }
  cv->qwait = 2;
  /* cvComputeEtaqm1 */
  cv->etaqm1 = ZERO;
  if (cv->q > 1) {
    const double ddn = wrms(cv->zn[cv->q], cv->ewt, n) * cv->tq[1];
    cv->etaqm1 = ONE / (rpowR(BIAS1 * ddn, ONE / cv->q) + ADDON);
  }
  /* cvComputeEtaqp1 */
  cv->etaqp1 = ZERO;
  if (cv->q != cv->qmax && cv->saved_tq5 != ZERO) {
    const double cquot = (cv->tq[5] / cv->saved_tq5) * rpowI(cv->h / cv->tau[2], cv->L);
    for (int i = 0; i < n; i++) cv->tempv[i] = -cquot * cv->zn[cv->qmax][i] + cv->acor[i];
    const double dup = wrms(cv->tempv, cv->ewt, n) * cv->tq[3];
    cv->etaqp1 = ONE / (rpowR(BIAS3 * dup, ONE / (cv->L + 1)) + ADDON);
  }
  /* cvChooseEta */
  double etam = (cv->etaq > cv->etaqp1) ? cv->etaq : cv->etaqp1;
  etam = (cv->etaqm1 > etam) ? cv->etaqm1 : etam;
  if ((etam > ETA_MIN_FX) && (etam < ETA_MAX_FX)) {
    cv->eta = ONE;
    cv->qprime = cv->q;
  } else if (etam == cv->etaq) {
    cv->eta = cv->etaq;
    cv->qprime = cv->q;
  } else if (etam == cv->etaqm1) {
    cv->eta = cv->etaqm1;
    cv->qprime = cv->q - 1;
  } else {
    cv->eta = cv->etaqp1;
    cv->qprime = cv->q + 1;
    for (int i = 0; i < n; i++) cv->zn[cv->qmax][i] = cv->acor[i];
  }
  set_eta(cv);
}

static int step(cvmem* cv)
{
  int ncf = 0, nef = 0, nflag, kflag, eflag;
  double dsm = 0.0;
  if ((cv->nst > 0) && (cv->hprime != cv->h)) adjust_params(cv);
  const double saved_t = cv->tn;
  nflag = FIRST_CALL;
  for (;;) {
    predict(cv);
    set_bdf(cv);
    nflag = nls(cv, nflag);
    kflag = handle_nflag(cv, &nflag, saved_t, &ncf);
    if (kflag == PREDICT_AGAIN) continue;
    if (kflag != DO_ERROR_TEST) return kflag;
    eflag = do_error_test(cv, &nflag, saved_t, &nef, &dsm);
    if (eflag == TRY_AGAIN) continue;
    if (eflag != 0) return eflag;
    break;
  }
  complete_step(cv);
  prepare_next_step(cv, dsm);
  cv->etamax = (cv->nst <= SMALL_NST) ? ETA_MAX_ES : ETA_MAX_GS;
  for (int i = 0; i < cv->n; i++) cv->acor[i] *= cv->tq[2];
  return 0;
}

static double upper_bound_h0(cvmem* cv, double tdist)
{
  double hub_inv = 0.0;
  for (int i = 0; i < cv->n; i++) {
    const double t2 = fabs(cv->zn[0][i]);
    const double t1 = HUB_FACTOR * t2 + 1.0 / cv->ewt[i]; /* cv_efun then N_VInv: 1/(1/(rtol|y|+atol)) */
    const double r = fabs(cv->zn[1][i]) / t1;
    if (r > hub_inv) hub_inv = r;
  }
  double hub = HUB_FACTOR * tdist;
  if (hub * hub_inv > ONE) hub = ONE / hub_inv;
  return hub;
}

static int hin(cvmem* cv, double tout)
{
  const double uround = DBL_EPSILON;
  double tdiff = tout - cv->tn;
  if (tdiff == ZERO) return CV_TOO_CLOSE;
  const int sign = (tdiff > ZERO) ? 1 : -1;
  const double tdist = fabs(tdiff);
  const double tround = uround * ((fabs(cv->tn) > fabs(tout)) ? fabs(cv->tn) : fabs(tout));
  if (tdist < TWO * tround) return CV_TOO_CLOSE;
  const double hlb = HLB_FACTOR * tround;
  const double hub = upper_bound_h0(cv, tdist);
  double hg = sqrt(hlb * hub);
  if (hub < hlb) {
    cv->h = (sign == -1) ? -hg : hg;
    return 0;
  }
  double hnew = hg, hrat, yddnrm;
  for (int count1 = 1; count1 <= MAX_ITERS; count1++) {
    const double hgs = hg * sign;
    /* cvYddNorm */
    for (int i = 0; i < cv->n; i++) cv->y[i] = hgs * cv->zn[1][i] + cv->zn[0][i];
    f(cv, cv->tn + hgs, cv->y, cv->tempv);
    /* N_VLinearSum(1/hg, tempv, -1/hg, zn1): a == -b dispatches to VScaleDiff (nvector_serial.c:452) */
    for (int i = 0; i < cv->n; i++) cv->tempv[i] = (ONE / hgs) * (cv->tempv[i] - cv->zn[1][i]);
    yddnrm = wrms(cv->tempv, cv->ewt, cv->n);
    hnew = (yddnrm * hub * hub > TWO) ? sqrt(TWO / yddnrm) : sqrt(hg * hub);
    if (count1 == MAX_ITERS) break;
    hrat = hnew / hg;
    if ((hrat > 0.5) && (hrat < TWO)) break;
    if ((count1 > 1) && (hrat > TWO)) {
      hnew = hg;
      break;
    }
    hg = hnew;
  }
  double h0 = H_BIAS * hnew;
  if (h0 < hlb) h0 = hlb;
  if (h0 > hub) h0 = hub;
  if (sign == -1) h0 = -h0;
  cv->h = h0;
  return 0;
}

int orc_cvode_cell(const plm_mech_tables* m, const orc_opts* o, double* yio, double rhoe_init, double rhoesrc_ext,
                   const double* rYsrc, double t0, double tout, orc_stats* st)
{
  static _Thread_local cvmem cvs;
  cvmem* cv = &cvs;
  const int n = m->ns + 1;
  memset(st, 0, sizeof(*st));
  cv->m = m;
  cv->n = n;
  cv->qmax = o->max_order;
  cv->maxcor = o->max_nls_iters;
  cv->maxnef = 100;
  cv->mxstep = 10000;
  cv->msbj = 100;
  cv->ajac = o->analytical_jacobian;
  cv->rt = o->reactor_type;
  cv->rtol = o->rtol;
  for (int i = 0; i < n; i++) cv->atolv[i] = (o->typ_vals[0] > 0.0) ? o->typ_vals[i] * o->atol : o->atol;
  cv->rhoe_init = rhoe_init;
  cv->rhoesrc = rhoesrc_ext;
  cv->rYsrc = rYsrc;
  cv->nlscoef = CORTES;
  /* CVodeReInit */
  cv->tn = t0;
  cv->q = 1;
  cv->L = 2;
  cv->qwait = cv->L;
  cv->etamax = ETA_MAX_FS;
  cv->hu = ZERO;
  for (int i = 0; i < n; i++) cv->zn[0][i] = yio[i];
  cv->nst = cv->nfe = cv->ncfn = cv->netf = cv->nni = cv->nsetups = cv->nstlp = 0;
  cv->nje = cv->nstlj = 0;
  cv->saved_tq5 = ZERO; /* always overwritten before first use (see cvCompleteStep) */
  cv->indx_acor = cv->qmax;
  cv->crate = ONE;
  cv->delp = ZERO;
  cv->jcur = 0;
  cv->eta = ONE;
  cv->qprime = 1;
  for (int i = 0; i < ORC_QMAX + 3; i++) cv->tau[i] = cv->l[i] = ZERO;
  for (int i = 0; i < 6; i++) cv->tq[i] = ZERO;

  int istate = 0;
  /* first-call block */
  if (ewtset(cv, cv->zn[0], cv->ewt) != 0) {
    st->flag = -22;
    return st->flag;
  }
  f(cv, cv->tn, cv->zn[0], cv->zn[1]);
  istate = hin(cv, tout);
  if (istate != 0) {
    st->flag = istate;
    st->nfe = cv->nfe;
    return istate;
  }
  cv->hscale = cv->h;
  cv->hprime = cv->h;
  for (int i = 0; i < n; i++) cv->zn[1][i] *= cv->h;

  long nstloc = 0;
  for (;;) {
    if (cv->nst > 0) {
      if (ewtset(cv, cv->zn[0], cv->ewt) != 0) {
        istate = -22;
        for (int i = 0; i < n; i++) yio[i] = cv->zn[0][i];
        break;
      }
    }
    if (nstloc >= cv->mxstep) {
      istate = CV_TOO_MUCH_WORK;
      for (int i = 0; i < n; i++) yio[i] = cv->zn[0][i];
      break;
    }
    const double nrm = wrms(cv->zn[0], cv->ewt, n);
    if (DBL_EPSILON * nrm > ONE) {
      istate = CV_TOO_MUCH_ACC;
      for (int i = 0; i < n; i++) yio[i] = cv->zn[0][i];
      break;
    }
    const int kflag = step(cv);
    if (kflag != 0) {
      istate = kflag;
      for (int i = 0; i < n; i++) yio[i] = cv->zn[0][i];
      break;
    }
    nstloc++;
    if ((cv->tn - tout) * cv->h >= ZERO) {
      /* CVodeGetDky(tout, 0): N_VLinearCombination over zn[q..0] with cvals s^j */
      const double s = (tout - cv->tn) / cv->h;
      double cvals[ORC_QMAX + 1];
      for (int j = cv->q; j >= 0; j--) {
        cvals[j] = ONE;
        for (int i = 0; i < j; i++) cvals[j] *= s;
      }
      for (int i = 0; i < n; i++) {
        /* N_VLinearCombination_Serial: z = c0*X0, then z += ci*Xi in the order q, q-1, ..., 0 */
        double z = cvals[cv->q] * cv->zn[cv->q][i];
        for (int j = cv->q - 1; j >= 0; j--) z += cvals[j] * cv->zn[j][i];
        yio[i] = z;
      }
      istate = 0;
      break;
    }
  }
  st->nst = cv->nst;
  st->nfe = cv->nfe;
  st->nje = cv->nje;
  st->nsetups = cv->nsetups;
  st->nni = cv->nni;
  st->ncfn = cv->ncfn;
  st->netf = cv->netf;
  st->flag = istate;
  return istate;
}
