/* TEST INFRASTRUCTURE (tests/test_gj_model.py): a scalar C model of the two Gauss-Jordan inverse routines of
 * pelelm_b200/csrc/bdf_warp.cuh -- gjinv_warp (one pivot per sweep) and gjinv2_warp (two pivots per sweep) -- with the
 * lanes of the warp played by loops, phase by phase, and the same explicit fused multiply-adds. It pins the claim the
 * device code relies on: the two-pivot routine gives every element the same operations in the same order, so both
 * return the same matrix (up to the sign of exact zeros), the same permutation and the same singularity code.
 * Compile with -ffp-contract=off: only the fma() calls below may fuse. */
#include <math.h>
#include <string.h>

#define BIG 0x7fffffff

/* first largest |a| wins (denseGETRF, sundials_dense.c:150-152); NaN never wins; returns -1.0 / BIG when no candidate */
static void argmax_first(const double* v, const int* idx, int nl, double* vout, int* iout)
{
  double bv = -1.0;
  int bi = BIG;
  for (int l = 0; l < nl; ++l)
    if (v[l] >= 0.0 && (v[l] > bv || (v[l] == bv && idx[l] < bi))) {
      bv = v[l];
      bi = idx[l];
    }
  *vout = bv;
  *iout = bi;
}

/* gjinv_warp: lane = row (rows lane, lane + 32, ...) */
int gj1(double* A, int n, int ld, int* perm)
{
  for (int i = 0; i < n; ++i) perm[i] = i;
  for (int k = 0; k < n; ++k) {
    double v = -1.0;
    int l = BIG;
    for (int i = k; i < n; ++i) {
      const double a = fabs(A[i * ld + k]);
      if (a > v) {
        v = a;
        l = i;
      }
    }
    if (!(v > 0.0)) return k + 1;
    const double pivinv = 1.0 / A[l * ld + k];
    for (int j = 0; j < n; ++j) {
      const double t = A[l * ld + j];
      if (l != k) A[l * ld + j] = A[k * ld + j];
      A[k * ld + j] = (j == k) ? pivinv : t * pivinv;
    }
    if (l != k) {
      const int t = perm[l];
      perm[l] = perm[k];
      perm[k] = t;
    }
    const double* rk = A + k * ld;
    for (int i = 0; i < n; ++i) {
      if (i == k) continue;
      double* row = A + i * ld;
      const double f = row[k];
      row[k] = 0.0;
      for (int j = 0; j < n; ++j) row[j] = fma(-rk[j], f, row[j]);
    }
  }
  return 0;
}

/* gjinv2_warp, n even, two rows per lane (rows lane and lane + 32; 32 < n <= 64) */
int gj2(double* A, int n, int ld, int* perm)
{
  double fA[64], fB[64], v[64];
  int idx[64], on[64];
  for (int i = 0; i < n; ++i) perm[i] = i;
  for (int k = 0; k < n; k += 2) {
    const int k1 = k + 1;
    /* pivot k */
    for (int lane = 0; lane < 32; ++lane) {
      const int i0 = lane, i1 = lane + 32;
      double bv = -1.0;
      int bi = BIG;
      if (i0 >= k) {
        bv = fabs(A[i0 * ld + k]);
        bi = i0;
        if (!(bv > -1.0)) {
          bv = -1.0;
          bi = BIG;
        }
      }
      if (i1 < n && i1 >= k) {
        const double a = fabs(A[i1 * ld + k]);
        if (a > bv) {
          bv = a;
          bi = i1;
        }
      }
      v[lane] = bv;
      idx[lane] = bi;
    }
    double pv;
    int l;
    argmax_first(v, idx, 32, &pv, &l);
    if (!(pv > 0.0)) return k + 1;
    const double pivinv = 1.0 / A[l * ld + k];
    for (int j = 0; j < n; ++j) {
      const double t = A[l * ld + j];
      if (l != k) A[l * ld + j] = A[k * ld + j];
      A[k * ld + j] = (j == k) ? pivinv : t * pivinv;
    }
    if (l != k) {
      const int t = perm[l];
      perm[l] = perm[k];
      perm[k] = t;
    }
    const double* rk = A + k * ld;
    double* rq = A + k1 * ld;
    const double rk1 = rk[k1];
    /* multipliers of pivot k, column k+1 as pivot k leaves it (index = row) */
    for (int i = 0; i < n; ++i) {
      fA[i] = (i != k) ? A[i * ld + k] : 0.0;
      fB[i] = (i != k) ? fma(-rk1, fA[i], A[i * ld + k1]) : rk1;
    }
    /* pivot k+1 */
    for (int lane = 0; lane < 32; ++lane) {
      const int i0 = lane, i1 = lane + 32;
      double bv = -1.0;
      int bi = BIG;
      if (i0 >= k1) {
        bv = fabs(fB[i0]);
        bi = i0;
        if (!(bv > -1.0)) {
          bv = -1.0;
          bi = BIG;
        }
      }
      if (i1 < n && i1 >= k1) {
        const double a = fabs(fB[i1]);
        if (a > bv) {
          bv = a;
          bi = i1;
        }
      }
      v[lane] = bv;
      idx[lane] = bi;
    }
    int l1;
    argmax_first(v, idx, 32, &pv, &l1);
    if (!(pv > 0.0)) return k1 + 1;
    const double cl1 = fB[l1], fl1 = fA[l1], ck1 = fB[k1], fk1 = fA[k1];
    const double pivinv1 = 1.0 / cl1;
    /* rows l1 and k+1 through pivot k, swapped, new row k+1 scaled */
    for (int j = 0; j < n; ++j) {
      const double r = rk[j];
      const double tl = fma(-r, fl1, (j == k) ? 0.0 : A[l1 * ld + j]);
      if (l1 != k1) A[l1 * ld + j] = fma(-r, fk1, (j == k) ? 0.0 : rq[j]);
      rq[j] = (j == k1) ? pivinv1 : tl * pivinv1;
    }
    if (l1 != k1) {
      const int t = perm[l1];
      perm[l1] = perm[k1];
      perm[k1] = t;
    }
    for (int i = 0; i < n; ++i) {
      on[i] = (i != k && i != k1);
      if (l1 != k1 && i == l1) {
        fA[i] = 0.0;
        fB[i] = ck1;
      } else if (on[i]) {
        A[i * ld + k] = 0.0;
      }
    }
    /* the sweep: pivot k then pivot k+1 on every element of the swept rows */
    for (int i = 0; i < n; ++i) {
      if (!on[i]) continue;
      double* row = A + i * ld;
      for (int j = 0; j < n; ++j) row[j] = fma(-rq[j], fB[i], fma(-rk[j], fA[i], row[j]));
      row[k1] = fma(-pivinv1, fB[i], 0.0);
    }
    /* row k through pivot k+1 */
    for (int j = 0; j < n; ++j) A[k * ld + j] = fma(-rq[j], rk1, (j == k1) ? 0.0 : A[k * ld + j]);
  }
  return 0;
}
