// Per-cell variable-order (1..5) variable-step BDF integrator with modified Newton and dense LU,
// one warp per cell: state vectors are lane-resident (component i = lane + 32*slot), the Nordsieck
// array and the Newton matrix live in the warp's shared memory (compile-time layout, chem_warp.cuh Lay<>), the saved
// Jacobian in a per-warp global scratch (L2 resident; touched only at the ~50 linear-solver setups per cell), the
// step-size / order scalars and the BDF coefficient arrays in registers (every lane computes the same bits:
// butterfly reductions), so all control flow is warp-uniform and needs no broadcasts.
//
// Behavioural specification = SUNDIALS CVODE 6.2.0 as configured by PelePhysics' CPU ReactorCvode
// (constants cvode_impl.h:57-130, cvode.c:111-118, cvode_nls.c:29-31,
// cvode_ls_impl.h:40-42). Function names below name the CVODE routine whose decisions they reproduce:
//   CVode cvode.c:990-1476 | cvHin :1945 | cvStep :2143 | cvAdjustParams/IncreaseBDF/DecreaseBDF :2265-2451
//   cvRescale :2457 | cvPredict :2485 | cvSetBDF/cvSetTqBDF :2691-2779 | cvNls :2781 | cvHandleNFlag :2954
//   cvRestore :3008 | cvDoErrorTest :3048 | cvCompleteStep :3162 | cvPrepareNextStep :3218 | cvSetEta :3262
//   cvComputeEtaqm1/qp1 :3299/:3318 | cvChooseEta :3347 | CVodeGetDky :1490
//   SUNNonlinSolSolve_Newton sunnonlinsol_newton.c:187 | cvNlsConvTest/Residual/LSetup cvode_nls.c:251-388
//   cvLsSetup/cvLsLinSys/cvLsSolve cvode_ls.c:1270-1720 | denseGETRF/GETRS sundials_dense.c:135-232
// One deliberate arithmetic difference: the linear-solver setup forms the inverse of the Newton matrix (Gauss-Jordan
// with the same partial-pivoting rule) and each Newton iteration is a mat-vec, see gjinv_warp below.
#pragma once
#include <cfloat>

#include "chem_warp.cuh"

namespace plm {

struct BdfOpts {
  double rtol, atol;
  int qmax;      // cvode.max_order
  int maxcor;    // ode.max_nls_iters
  int ajac;      // 1 analytic Jacobian (denseAJ_direct), 0 difference quotient (dense_direct)
  int mxstep;    // 10000 (ReactorCvode.cpp:276)
  int maxnef;    // 100   (ReactorCvode.cpp:267)
  int msbj;      // 100   (ReactorCvode.cpp:285)
};

struct CellStats {
  int nst, nfe, nje, nsetups, nni, ncfn, netf, flag;
};

namespace cvc {  // CVODE constants
constexpr double ETA_MIN_FX = 0.0, ETA_MAX_FX = 1.5, ETA_MAX_FS = 10000.0, ETA_MAX_ES = 10.0, ETA_MAX_GS = 10.0,
                 ETA_MIN = 0.1, ETA_MAX_EF = 0.2, ETA_MIN_EF = 0.1, ETA_CF = 0.25, ADDON = 0.000001, BIAS1 = 6.0,
                 BIAS2 = 6.0, BIAS3 = 10.0, HLB_FACTOR = 100.0, HUB_FACTOR = 0.1, H_BIAS = 0.5, CORTES = 0.1,
                 CRDOWN = 0.3, RDIV = 2.0, DGMAX_LSETUP = 0.3, DGMAX_JBAD = 0.2;
constexpr int SMALL_NST = 10, SMALL_NEF = 2, LONG_WAIT = 10, MXNCF = 10, MXNEF1 = 3, MAX_ITERS = 4, MSBP = 20;
constexpr int CV_TOO_MUCH_WORK = -1, CV_TOO_MUCH_ACC = -2, CV_ERR_FAILURE = -3, CV_CONV_FAILURE = -4,
              CV_ILL_INPUT = -22, CV_TOO_CLOSE = -27;
enum { DO_ERROR_TEST = 2, PREDICT_AGAIN = 3, TRY_AGAIN = 5 };
enum { FIRST_CALL = 6, PREV_CONV_FAIL = 7, PREV_ERR_FAIL = 8 };
enum { NO_FAILURES = 0, FAIL_BAD_J = 1, FAIL_OTHER = 2 };
enum { NLS_SUCCESS = 0, NLS_CONV_RECVR = 902 };
}  // namespace cvc

// SUNRpowerR(base, 1/k) (sundials_math.c:56): pow(base, 1.0/k) for base > 0 else 0. The step-size controller only
// takes k-th roots with k = L or L+1 (2..7): square and cube roots have dedicated short sequences.
__device__ __noinline__ double rroot(double base, int k)
{
  if (base <= 0.0) return 0.0;
  if (k == 2) return sqrt(base);
  if (k == 3) return cbrt(base);
  if (k == 4) return sqrt(sqrt(base));
  return dexp(dlog(base) / double(k));
}
// 1.0 / k for the small integers the BDF coefficient formulas divide by (k = 1..7), correctly rounded constants
__constant__ double c_inv_int[8] = {0.0, 1.0, 0.5, 1.0 / 3.0, 0.25, 0.2, 1.0 / 6.0, 1.0 / 7.0};
__device__ __forceinline__ double inv_int(int k) { return c_inv_int[k & 7]; }
__device__ __forceinline__ double rpowI(double base, int e)
{
  double prod = 1.0;
  const int expt = e < 0 ? -e : e;
  for (int i = 1; i <= expt; i++) prod *= base;
  if (e < 0) prod = 1.0 / prod;
  return prod;
}

// ---- Newton matrix: in-place Gauss-Jordan inverse with partial pivoting (row-major, lane = row) ----
// CVODE factors A = I - gamma*J once per setup (denseGETRF) and then runs two triangular solves per Newton iteration
// (denseGETRS); with ~16 iterations per setup and one warp per system the triangular solves are ~44 dependent
// shuffle/FMA steps each. Here the setup forms W = (P A)^-1 (same pivot rule as denseGETRF: first largest |a_ik|,
// sundials_dense.c:150-152) and every Newton iteration is one dense mat-vec x = W (P b): no dependent chain, no
// predicates. The iteration matrix is an approximation to begin with (stale J, stale gamma); the converged Newton
// solution does not depend on how W b is rounded.
template <class L>
__device__ __noinline__ int gjinv_warp(double* A)
{
  constexpr int NSLOT = L::NSLOT, n = L::N, ld = L::LD;
  double* wb = warp_base<L>();
  int* perm = reinterpret_cast<int*>(wb + L::PERM);
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int s = 0; s < NSLOT; ++s)
    if (lane + 32 * s < n) perm[lane + 32 * s] = lane + 32 * s;
  __syncwarp();
#pragma unroll 1
  for (int k = 0; k < n; ++k) {
    double v = -1.0;
    int idx = 0x7fffffff;
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      const int i = lane + 32 * s;
      if (i >= k && i < n) {
        const double a = fabs(A[i * ld + k]);
        if (a > v) {  // strict: first occurrence wins
          v = a;
          idx = i;
        }
      }
    }
    if (!L::BIG) {  // small systems: the butterfly measured 4 % faster on drm19 than the reductions below
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(kFull, v, o);
        const int oi = __shfl_xor_sync(kFull, idx, o);
        if (ov > v || (ov == v && oi < idx)) {
          v = ov;
          idx = oi;
        }
      }
    } else {
      // arg-max over the warp with three hardware reductions (redux.sync) instead of a 5-level butterfly of
      // (value, index) pairs: non-negative doubles order like their bit patterns, so the maximum is the maximum of the
      // high words, then of the low words among those, and the first row wins ties (denseGETRF's rule). A NaN entry
      // (never larger in the reference's comparison) gets the smallest key.
      const bool ok = v >= 0.0;  // false for "no candidate row in this lane" (v = -1) and for NaN
      const unsigned hi = ok ? unsigned(__double2hiint(v)) + 1u : 0u;  // +1: a valid zero pivot candidate still beats "none"
      const unsigned mhi = __reduce_max_sync(kFull, hi);
      const bool c1 = ok && hi == mhi;
      const unsigned lo = c1 ? unsigned(__double2loint(v)) : 0u;
      const unsigned mlo = __reduce_max_sync(kFull, lo);
      const bool win = c1 && lo == mlo;
      const unsigned first = __reduce_min_sync(kFull, win ? unsigned(idx) : 0x7fffffffu);
      const int src = int(first & 31u);  // the winner row's lane (rows are lane + 32 * slot)
      v = __shfl_sync(kFull, v, src);
      idx = (mhi == 0u) ? 0x7fffffff : int(first);
      if (mhi == 0u) v = -1.0;
    }
    const int l = idx;
    if (!(v > 0.0)) return k + 1;  // zero pivot column (or nothing but NaN in it): singular
    const double pivinv = 1.0 / A[l * ld + k];
    __syncwarp();
    // swap rows l and k and scale the pivot row: lane = column
    for (int j = lane; j < n; j += 32) {
      const double t = A[l * ld + j];
      if (l != k) A[l * ld + j] = A[k * ld + j];
      A[k * ld + j] = (j == k) ? pivinv : t * pivinv;
    }
    if (lane == 0 && l != k) {
      const int t = perm[l];
      perm[l] = perm[k];
      perm[k] = t;
    }
    __syncwarp();
    const double* rk = A + k * ld;
    if (NSLOT == 2 && L::PAIRS) {
      // both rows of the lane in one sweep: the pivot row is read once for the two, and the two updates are independent
      // chains (the inverse is 26 % of the dodecane_lu kernel, stalled on shared-memory latency)
      const int i1 = lane + 32;
      const bool on0 = lane != k, on1 = (i1 != k) && (i1 < n);
      double2* r0 = reinterpret_cast<double2*>(A + lane * ld);
      double2* r1 = reinterpret_cast<double2*>(A + (i1 < n ? i1 : n - 1) * ld);
      const double f0 = on0 ? A[lane * ld + k] : 0.0, f1 = on1 ? A[i1 * ld + k] : 0.0;
      if (on0) A[lane * ld + k] = 0.0;
      if (on1) A[i1 * ld + k] = 0.0;
      const double2* rk2 = reinterpret_cast<const double2*>(rk);
#pragma unroll 3
      for (int j = 0; j < n / 2; ++j) {
        const double2 p = rk2[j];
        double2 a = r0[j], b = r1[j];
        a.x -= p.x * f0;
        a.y -= p.y * f0;
        b.x -= p.x * f1;
        b.y -= p.y * f1;
        if (on0) r0[j] = a;
        if (on1) r1[j] = b;
      }
    } else
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      const int i = lane + 32 * s;
      if (i != k && i < n) {
        double* row = A + i * ld;
        const double f = row[k];
        row[k] = 0.0;
        if (L::PAIRS) {  // rows are 16-byte aligned and an odd number of 16-byte units apart: conflict-free pair sweeps
          const double2* rk2 = reinterpret_cast<const double2*>(rk);
          double2* row2 = reinterpret_cast<double2*>(row);
          // next column pair loaded before this one is stored (see gjinv2_warp): same arithmetic, no load -> store -> load chain
          double2 p = rk2[0], r = row2[0];
#pragma unroll 2
          for (int j = 0; j < n / 2; ++j) {
#if defined(PLM_GJPIPE) && !PLM_GJPIPE
            const int jn = j;  // development: the unpipelined loop (loads of pair j issued in iteration j)
            p = rk2[j];
            r = row2[j];
#else
            // the last iteration re-reads its own pair (not the pair past the row, which belongs to another lane's row: an
            // unused value, but a read racecheck rightly reports)
            const int jn = (j + 1 < n / 2) ? j + 1 : j;
#endif
            const double2 pn = rk2[jn], rn = row2[jn];
            r.x -= p.x * f;
            r.y -= p.y * f;
            row2[j] = r;
            p = pn;
            r = rn;
          }
        } else {
#pragma unroll 2
          for (int j = 0; j < n; ++j) row[j] -= rk[j] * f;
        }
      }
    }
    __syncwarp();
  }
  return 0;
}

// arg-max of (v, idx) over the warp, first row winning ties, v < 0 / NaN = "no candidate" (three hardware reductions, see
// gjinv_warp); returns warp-uniform v (-1 when no lane has a candidate) and idx.
__device__ __forceinline__ void argmax_first(double& v, int& idx)
{
  const bool ok = v >= 0.0;
  const unsigned hi = ok ? unsigned(__double2hiint(v)) + 1u : 0u;
  const unsigned mhi = __reduce_max_sync(kFull, hi);
  const bool c1 = ok && hi == mhi;
  const unsigned lo = c1 ? unsigned(__double2loint(v)) : 0u;
  const unsigned mlo = __reduce_max_sync(kFull, lo);
  const bool win = c1 && lo == mlo;
  const unsigned first = __reduce_min_sync(kFull, win ? unsigned(idx) : 0x7fffffffu);
  v = __shfl_sync(kFull, v, int(first & 31u));
  idx = (mhi == 0u) ? 0x7fffffff : int(first);
  if (mhi == 0u) v = -1.0;
}

// The same inverse, two pivots per sweep (systems of two rows per lane, n even: dodecane_lu). Every element receives the
// same fused multiply-adds in the same order as in gjinv_warp -- pivot k, then pivot k+1 -- so W is the same matrix (up
// to the sign of exact zeros); what changes is the traffic: a lane's rows are read and written once per PAIR of pivots
// (18 instead of 34 shared-memory wavefronts per 16-byte column pair) and eight independent FMAs ride on every round
// trip. The panel work that makes this possible: after pivot row k is final, column k+1 as pivot k leaves it is
// formed in registers (c), the second pivot is searched in it, and the two rows that swap places for pivot k+1 are
// brought up to date with pivot k on their own (lane = column) before the swap and the scaling. Row k itself, which
// every lane reads as the first pivot row during the sweep, gets its pivot-(k+1) update after the sweep.
template <class L>
__device__ __noinline__ int gjinv2_warp(double* A)
{
  constexpr int n = L::N, ld = L::LD;
  static_assert(L::NSLOT == 2 && L::PAIRS && n % 2 == 0 && n > 32, "two rows per lane, even order");
  double* wb = warp_base<L>();
  int* perm = reinterpret_cast<int*>(wb + L::PERM);
  const int lane = threadIdx.x & 31;
  const int i0 = lane, i1 = lane + 32;
  const bool has1 = i1 < n;
  const int i1c = has1 ? i1 : n - 1;  // clamped for addressing; its results are never stored
  perm[i0] = i0;
  if (has1) perm[i1] = i1;
  __syncwarp();
  double* row0 = A + i0 * ld;
  double* row1 = A + i1c * ld;
#pragma unroll 1
  for (int k = 0; k < n; k += 2) {
    const int k1 = k + 1;
    // ---- pivot k: search, swap, scale (as gjinv_warp)
    double v = -1.0;
    int idx = 0x7fffffff;
    if (i0 >= k) {
      v = fabs(row0[k]);
      idx = i0;
      if (!(v > -1.0)) {  // NaN: never larger in the reference's comparison
        v = -1.0;
        idx = 0x7fffffff;
      }
    }
    if (has1 && i1 >= k) {
      const double a = fabs(row1[k]);
      if (a > v) {
        v = a;
        idx = i1;
      }
    }
    argmax_first(v, idx);
    if (!(v > 0.0)) return k + 1;
    const int l = idx;
    const double pivinv = 1.0 / A[l * ld + k];
    __syncwarp();
    for (int j = lane; j < n; j += 32) {
      const double t = A[l * ld + j];
      if (l != k) A[l * ld + j] = A[k * ld + j];
      A[k * ld + j] = (j == k) ? pivinv : t * pivinv;
    }
    if (lane == 0 && l != k) {
      const int t = perm[l];
      perm[l] = perm[k];
      perm[k] = t;
    }
    __syncwarp();
    const double* rk = A + k * ld;
    double* rq = A + k1 * ld;
    // ---- multipliers of pivot k and column k+1 as pivot k leaves it
    const double rk1 = rk[k1];
    double fA0 = (i0 != k) ? row0[k] : 0.0;
    double fA1 = (has1 && i1 != k) ? row1[k] : 0.0;
    double fB0 = (i0 != k) ? fma(-rk1, fA0, row0[k1]) : rk1;
    double fB1 = (i1 != k) ? fma(-rk1, fA1, row1[k1]) : rk1;
    // ---- pivot k+1: search in that column
    v = -1.0;
    idx = 0x7fffffff;
    if (i0 >= k1) {
      v = fabs(fB0);
      idx = i0;
      if (!(v > -1.0)) {
        v = -1.0;
        idx = 0x7fffffff;
      }
    }
    if (has1 && i1 >= k1) {
      const double a = fabs(fB1);
      if (a > v) {
        v = a;
        idx = i1;
      }
    }
    argmax_first(v, idx);
    if (!(v > 0.0)) return k1 + 1;
    const int l1 = idx;
    const double cl1 = __shfl_sync(kFull, l1 < 32 ? fB0 : fB1, l1 & 31), fl1 = __shfl_sync(kFull, l1 < 32 ? fA0 : fA1, l1 & 31);
    const double ck1 = __shfl_sync(kFull, k1 < 32 ? fB0 : fB1, k1 & 31), fk1 = __shfl_sync(kFull, k1 < 32 ? fA0 : fA1, k1 & 31);
    const double pivinv1 = 1.0 / cl1;
    __syncwarp();
    // ---- rows l1 and k+1 brought up to date with pivot k, swapped, the new row k+1 scaled (lane = column)
    for (int j = lane; j < n; j += 32) {
      const double r = rk[j];
      const double tl = fma(-r, fl1, (j == k) ? 0.0 : A[l1 * ld + j]);
      if (l1 != k1) A[l1 * ld + j] = fma(-r, fk1, (j == k) ? 0.0 : rq[j]);
      rq[j] = (j == k1) ? pivinv1 : tl * pivinv1;
    }
    if (lane == 0 && l1 != k1) {
      const int t = perm[l1];
      perm[l1] = perm[k1];
      perm[k1] = t;
    }
    // ---- what is left for the sweep, per row: pivot k (fA) then pivot k+1 (fB)
    //   row k      : nothing here (its pivot-(k+1) update follows the sweep: every lane reads it meanwhile)
    //   row k+1    : the second pivot row, final
    //   row l1     : holds the old row k+1, already through pivot k; multiplier of pivot k+1 = its column-(k+1) entry ck1
    bool on0 = i0 != k && i0 != k1, on1 = has1 && i1 != k && i1 != k1;
    if (l1 != k1) {
      if (i0 == l1) {
        fA0 = 0.0;
        fB0 = ck1;
      } else if (on0) {
        row0[k] = 0.0;
      }
      if (i1 == l1) {
        fA1 = 0.0;
        fB1 = ck1;
      } else if (on1) {
        row1[k] = 0.0;
      }
    } else {
      if (on0) row0[k] = 0.0;
      if (on1) row1[k] = 0.0;
    }
    __syncwarp();
    {
      // Software-pipelined by hand: the operands of column pair j+1 are loaded before pair j is stored. The compiler keeps
      // shared-memory loads behind earlier stores it cannot prove disjoint (everything here derives from one base
      // pointer), which left every iteration a full load -> FMA -> store round trip (the short-scoreboard stalls of the
      // capture). No lane stores into a row it reads as a pivot row (rows k and k+1 are "off"), and a lane's own rows
      // are only its own, so the early loads are safe.
      const double2* p2 = reinterpret_cast<const double2*>(rk);
      const double2* q2 = reinterpret_cast<const double2*>(rq);
      double2* r0 = reinterpret_cast<double2*>(row0);
      // a lane without a second row reads the pivot row instead (read-only during the sweep; its b is never stored): the
      // clamped row n-1 is being written by its owner meanwhile, an unused value but a read racecheck rightly reports
      double2* r1 = has1 ? reinterpret_cast<double2*>(row1) : const_cast<double2*>(p2);
      double2 p = p2[0], q = q2[0], a = r0[0], b = r1[0];
#pragma unroll 3
      for (int j = 0; j < n / 2; ++j) {
        const int jn = (j + 1 < n / 2) ? j + 1 : j;  // the last iteration re-reads its own pair, not another lane's row
        const double2 pn = p2[jn], qn = q2[jn], an = r0[jn], bn = r1[jn];
        a.x = fma(-q.x, fB0, fma(-p.x, fA0, a.x));
        a.y = fma(-q.y, fB0, fma(-p.y, fA0, a.y));
        b.x = fma(-q.x, fB1, fma(-p.x, fA1, b.x));
        b.y = fma(-q.y, fB1, fma(-p.y, fA1, b.y));
        if (on0) r0[j] = a;
        if (on1) r1[j] = b;
        p = pn;
        q = qn;
        a = an;
        b = bn;
      }
    }
    // column k+1 of the swept rows: zeroed before pivot k+1 acts on it (what the sweep left there is overwritten)
    if (on0) row0[k1] = fma(-pivinv1, fB0, 0.0);
    if (on1) row1[k1] = fma(-pivinv1, fB1, 0.0);
    __syncwarp();
    // ---- row k through pivot k+1 (multiplier = its column-(k+1) entry rk1)
    for (int j = lane; j < n; j += 32) A[k * ld + j] = fma(-rq[j], rk1, (j == k1) ? 0.0 : A[k * ld + j]);
    __syncwarp();
  }
  return 0;
}

// x = W (P b): in wb[FB] = b (every lane its own components), out = x of the lane's components.
template <class L>
__device__ __forceinline__ void solve_warp(double* wb, int lane, double (&x)[L::NSLOT], const double* __restrict__ Wt)
{
  constexpr int NSLOT = L::NSLOT, n = L::N, ld = L::LD;
  const int* perm = reinterpret_cast<const int*>(wb + L::PERM);
  __syncwarp();
  double t[NSLOT];
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) t[s] = (lane + 32 * s < n) ? wb[L::FB + perm[lane + 32 * s]] : 0.0;
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) wb[L::YB + lane + 32 * s] = t[s];  // YB is free between right-hand-side calls
  __syncwarp();
  const double* pb = wb + L::YB;
  if constexpr (L::POOL > 0) {
    // W streamed from the warp's global buffer (L2), transposed: element (row, col) at Wt[col * WLD + row], so a column is
    // one coalesced 256-byte load per slot. Same products, same order of accumulation as the shared-memory form below.
    static_assert(NSLOT == 2 && n % 2 == 0, "pooled matrices: two rows per lane");
    constexpr int wld = L::WLD;
    const double2* pb2 = reinterpret_cast<const double2*>(pb);
    const double* w0 = Wt + lane;
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll 9
    for (int j = 0; j < n / 2; ++j) {
      const double u0 = __ldcg(w0 + (2 * j) * wld), u1 = __ldcg(w0 + (2 * j + 1) * wld);
      const double v0 = __ldcg(w0 + (2 * j) * wld + 32), v1 = __ldcg(w0 + (2 * j + 1) * wld + 32);
      const double2 b = pb2[j];
      a0 = fma(u0, b.x, a0);
      a1 = fma(u1, b.y, a1);
      b0 = fma(v0, b.x, b0);
      b1 = fma(v1, b.y, b1);
    }
    x[0] = a0 + a1;
    x[NSLOT - 1] = b0 + b1;
  } else if (NSLOT == 2 && L::PAIRS) {  // both rows of the lane against one pass over the right-hand side: four independent chains
    const double2* r0 = reinterpret_cast<const double2*>(wb + L::A + lane * ld);
    const double2* r1 = reinterpret_cast<const double2*>(wb + L::A + min(lane + 32, n - 1) * ld);
    const double2* pb2 = reinterpret_cast<const double2*>(pb);
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll 3
    for (int j = 0; j < n / 2; ++j) {
      const double2 b = pb2[j], u = r0[j], v = r1[j];
      a0 = fma(u.x, b.x, a0);
      a1 = fma(u.y, b.y, a1);
      b0 = fma(v.x, b.x, b0);
      b1 = fma(v.y, b.y, b1);
    }
    x[0] = a0 + a1;
    x[NSLOT - 1] = b0 + b1;
  } else
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) {
    const double* row = wb + L::A + min(lane + 32 * s, n - 1) * ld;
    double a0 = 0.0, a1 = 0.0;
    if (L::PAIRS) {
      const double2* row2 = reinterpret_cast<const double2*>(row);
      const double2* pb2 = reinterpret_cast<const double2*>(pb);
#pragma unroll 4
      for (int j = 0; j < n / 2; ++j) {
        const double2 w = row2[j], b = pb2[j];
        a0 = fma(w.x, b.x, a0);
        a1 = fma(w.y, b.y, a1);
      }
    } else {
#pragma unroll 4
      for (int j = 0; j + 1 < n; j += 2) {
        a0 = fma(row[j], pb[j], a0);
        a1 = fma(row[j + 1], pb[j + 1], a1);
      }
      if (n & 1) a0 = fma(row[n - 1], pb[n - 1], a0);
    }
    x[s] = a0 + a1;
  }
  __syncwarp();
}

// Integrator state of the warp's cell. Everything that lives longer than one statement sits in the warp's shared-memory
// area (scalars: every lane holds the same bits and stores the same value; per-lane vectors: lane-indexed), reached
// through references, so that only a handful of registers are live across the right-hand-side call. The first version
// kept it all in registers: ~90 live values across a call into a function that needs ~120 registers itself, i.e. spills
// around every call at 128 registers per thread and no way to run more than 16 warps per SM.
// A warp-uniform scalar of the integrator in the warp's shared-memory area: every lane reads it, lane 0 alone writes it,
// with a warp barrier on both sides of the store.  Round 1 let all 32 lanes store the same bits with no barrier ("benign"
// races, 58 racecheck hazards): correct with today's code generation, but not by the CUDA memory model.
template <class T>
struct WScalar {
  T* p;
  __device__ __forceinline__ operator T() const { return *p; }
  __device__ __forceinline__ const WScalar& operator=(T v) const
  {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) *p = v;
    __syncwarp();
    return *this;
  }
  // store of a GROUP of scalars: the caller brackets the group with one pair of __syncwarp() (every value of the group is
  // computed into registers first, so nothing inside the group reads what the group writes)
  __device__ __forceinline__ void put(T v) const
  {
    if ((threadIdx.x & 31) == 0) *p = v;
  }
  __device__ __forceinline__ const WScalar& operator=(const WScalar& o) const { return *this = T(o); }
  __device__ __forceinline__ const WScalar& operator+=(T v) const { return *this = T(*p + v); }
  __device__ __forceinline__ const WScalar& operator-=(T v) const { return *this = T(*p - v); }
  __device__ __forceinline__ const WScalar& operator*=(T v) const { return *this = T(*p * v); }
  __device__ __forceinline__ T operator++(int) const
  {
    const T old = *p;
    *this = T(old + 1);
    return old;
  }
  __device__ __forceinline__ T operator--(int) const
  {
    const T old = *p;
    *this = T(old - 1);
    return old;
  }
};
// A statistics counter (right-hand sides, Newton iterations, Jacobians, setups, failures): nothing in the integration reads
// it, so lane 0 keeps it alone -- plain increments, no barrier -- and every lane reads it once, behind a warp barrier, when
// the cell is done.  (As WScalar these six cost two warp barriers per increment: four per right-hand side.)
struct WCount {
  int* p;
  __device__ __forceinline__ void operator++(int) const
  {
    if ((threadIdx.x & 31) == 0) *p = *p + 1;
  }
  __device__ __forceinline__ void reset() const
  {
    if ((threadIdx.x & 31) == 0) *p = 0;
  }
  __device__ __forceinline__ int final_value() const { return *p; }  // after a __syncwarp()
};
template <class T>
struct WArray {
  T* p;
  __device__ __forceinline__ WScalar<T> operator[](int i) const { return WScalar<T>{p + i}; }
};

template <class L>
struct BdfWarp {
  static constexpr int NSLOT = L::NSLOT, n = L::N, NP = L::NP, QM = L::QM, ld = L::LD;
  static_assert(QM + 1 <= 6, "scalar area holds 6 entries per coefficient array");
  enum { SC_TAU = 0, SC_TQ = 8, SC_LC = 16, SC_RHOE = 24, SC_RHOESRC = 25, SC_D = 26, SC_I = 40 };
  struct LaneVec {  // per-lane vector in shared memory: component lane + 32*s
    double* p;
    __device__ __forceinline__ double& operator[](int s) const { return p[32 * s]; }
  };
  const BdfOpts& o;
  const int lane;
  double* const wb;
  double* const Jsave;  // [n*ld] per-warp global scratch
  double* const Wt;     // pooled matrices (Lay<>::POOL): [n][WLD] the transposed inverse of the last setup, behind Jsave
  const LaneVec atolv, ewt, acor;
  double yv[NSLOT], ft[NSLOT], dl[NSLOT], tv[NSLOT];  // short-lived temporaries
#ifdef PLM_SOFTALIGN
  bool in_setup = false;  // the warp is outside the soft barrier's group (difference-quotient Jacobian: f without arrival)
#endif
  const WArray<double> tau, lc, tq;
  const WScalar<double> tn, h, hprime, hscale, eta, etamax;
  const WScalar<double> rl1, gamma, gammap, gamrat, crate, delp, saved_tq5;
  double acnrm = 0.0;  // norm of the accepted correction: from the end of the Newton iteration to the error test, a register
  const WScalar<int> q, qprime, L_, qwait, indx_acor;
  const WScalar<int> nst, nstlp, nstlj;
  const WCount nfe, ncfn, netf, nni, nsetups, nje;
  const WScalar<int> jcur, convfail;

  __device__ __forceinline__ double* sc() const { return wb + L::SM; }
  __device__ __forceinline__ int* si() const { return reinterpret_cast<int*>(wb + L::SM + SC_I); }
  static __device__ __forceinline__ int* sip(double* w) { return reinterpret_cast<int*>(w + L::SM + SC_I); }
  __device__ BdfWarp(const BdfOpts& o_, int lane_, double* wb_, double* js)
      : o(o_), lane(lane_), wb(wb_), Jsave(js), Wt(js + n * L::LD),
        atolv{wb_ + L::ATOL + lane_}, ewt{wb_ + L::EWT + lane_}, acor{wb_ + L::ACOR + lane_},
        tau{wb_ + L::SM + SC_TAU}, lc{wb_ + L::SM + SC_LC}, tq{wb_ + L::SM + SC_TQ},
        tn{wb_ + L::SM + SC_D + 0}, h{wb_ + L::SM + SC_D + 1}, hprime{wb_ + L::SM + SC_D + 2}, hscale{wb_ + L::SM + SC_D + 3},
        eta{wb_ + L::SM + SC_D + 4}, etamax{wb_ + L::SM + SC_D + 5}, rl1{wb_ + L::SM + SC_D + 6}, gamma{wb_ + L::SM + SC_D + 7},
        gammap{wb_ + L::SM + SC_D + 8}, gamrat{wb_ + L::SM + SC_D + 9}, crate{wb_ + L::SM + SC_D + 10},
        delp{wb_ + L::SM + SC_D + 11}, saved_tq5{wb_ + L::SM + SC_D + 13},
        q{sip(wb_) + 0}, qprime{sip(wb_) + 1}, L_{sip(wb_) + 2}, qwait{sip(wb_) + 3}, indx_acor{sip(wb_) + 4}, nst{sip(wb_) + 5},
        nfe{sip(wb_) + 6}, ncfn{sip(wb_) + 7}, netf{sip(wb_) + 8}, nni{sip(wb_) + 9}, nsetups{sip(wb_) + 10}, nje{sip(wb_) + 11},
        nstlp{sip(wb_) + 12}, nstlj{sip(wb_) + 13}, jcur{sip(wb_) + 14}, convfail{sip(wb_) + 15}
  {
  }

  __device__ __forceinline__ bool valid(int s) const { return lane + 32 * s < n; }
  __device__ __forceinline__ double& zn(int j, int s) const { return wb[L::ZN + j * NP + lane + 32 * s]; }
  template <int SZ>
  static __device__ __forceinline__ double sel(const double (&a)[SZ], int i)
  {
    double v = a[0];
#pragma unroll
    for (int k = 1; k < SZ; ++k) v = (i == k) ? a[k] : v;
    return v;
  }
  static __device__ __forceinline__ double sel(const WArray<double>& a, int i) { return a[i]; }

  // sqrt(sum / n); a zero (or NaN / inf) sum goes through the library routine. One copy (instruction-cache footprint).
  static __device__ __noinline__ double wrms_finish(double sum)
  {
    const double v = sum * (1.0 / n);
    return (v > 1e-290 && v < 1e290) ? fsqrt(v) : sqrt(v);
  }
  template <class X, class W>
  __device__ __forceinline__ double wrms(const X& x, const W& w) const
  {
    double a = 0.0;
#pragma unroll
    for (int s = 0; s < NSLOT; ++s)
      if (valid(s)) {
        const double p = x[s] * w[s];
        a += p * p;
      }
    return wrms_finish(warp_sum(a));
  }
  // sum_i (zn[0]_i * ewt_i)^2 ; the caller compares it with n / uround^2 (cvode.c:1316-1321: uround * wrms > 1)
  __device__ __forceinline__ double wss_zn0() const
  {
    double a = 0.0;
#pragma unroll
    for (int s = 0; s < NSLOT; ++s)
      if (valid(s)) {
        const double p = zn(0, s) * ewt[s];
        a += p * p;
      }
    return warp_sum(a);
  }
  // cvEwtSetSV on zn[0]; false if a weight is non-positive
  __device__ __forceinline__ bool ewtset()
  {
    int bad = 0;
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      ewt[s] = 1.0;
      if (valid(s)) {
        const double t = o.rtol * fabs(zn(0, s)) + atolv[s];
        if (t <= 0.0) bad = 1;
        ewt[s] = (t > 1e-290) ? frcp(t) : 1.0 / t;
      }
    }
    return __any_sync(kFull, bad) == 0;
  }
  // ydot = f(t, y) through the shared-memory argument buffers of rhs_warp
  __device__ __forceinline__ void f(double t, const double (&y)[NSLOT], double (&ydot)[NSLOT])
  {
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) wb[L::YB + lane + 32 * s] = y[s];
#ifdef PLM_ALIGN
    // Development experiment (profiles/r02n_*): the warps of the CTA enter every right-hand side
    // together, so that they stream its code as one instruction stream.  Warps that ran out of cells keep arriving
    // (react_kernel's epilogue) until every warp of the CTA has: all threads execute the same number of barriers.
    __syncthreads_and(0);
#elif defined(PLM_SOFTALIGN)
    if (!in_setup) softbar_arrive_wait();  // (ends with a warp barrier)
    else __syncwarp();
#else
    __syncwarp();
#endif
    rhs_warp<L>(t);
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) ydot[s] = wb[L::FB + lane + 32 * s];
    nfe++;
  }

  // cvLsDenseDQJac (cvode_ls.c:1034-1107): fy = f(tn, yv) in ft
  __device__ __forceinline__ void dq_jac(double* Jm)
  {
    const double uround = DBL_EPSILON, srur = sqrt(uround);
    const double fnorm = wrms(ft, ewt);
    const double minInc = (fnorm != 0.0) ? (1000.0 * fabs(h) * uround * n * fnorm) : 1.0;
    double ysave[NSLOT], ft2[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) ysave[s] = yv[s];
#pragma unroll 1
    for (int j = 0; j < n; ++j) {
      double inc_l = 0.0;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s)
        if (lane + 32 * s == j) {
          inc_l = fmax(srur * fabs(ysave[s]), minInc / ewt[s]);
          yv[s] += inc_l;
        }
      const double inc = __shfl_sync(kFull, inc_l, j & 31);
      f(tn, yv, ft2);
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) yv[s] = ysave[s];
      const double inc_inv = 1.0 / inc;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s)
        if (valid(s)) Jm[(lane + 32 * s) * ld + j] = inc_inv * ft2[s] - inc_inv * ft[s];
    }
    __syncwarp();
  }

  // cvNlsLSetup + cvLsSetup + cvLsLinSys + dense factorisation
  __device__ __forceinline__ int lsetup()
  {
#ifdef PLM_SOFTALIGN
    softbar_leave();
    in_setup = true;
    const int ier_ = lsetup_body();
    in_setup = false;
    softbar_join();
    return ier_;
  }
  __device__ __forceinline__ int lsetup_body()
  {
#endif
    int pool = 0;
    double* A;
    if constexpr (L::POOL > 0) {
      pool = pool_acquire<L>();
      A = pool_matrix<L>(pool);
    } else {
      A = wb + L::A;
    }
    const double dgamma = fabs(fdiv(gamma, gammap) - 1.0);
    const bool jbad = (nst == 0) || (nst >= nstlj + o.msbj) ||
                      ((convfail == cvc::FAIL_BAD_J) && (dgamma < cvc::DGMAX_JBAD)) || (convfail == cvc::FAIL_OTHER);
    __syncwarp();
    if (jbad) {
      jcur = 1;
      if (o.ajac) {
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) wb[L::YB + lane + 32 * s] = yv[s];
        __syncwarp();
        jac_warp<L>(A);
      } else {
        dq_jac(A);
      }
      // large mechanisms: L2-only stores / loads (.cg) -- the small L1 left beside 190 KB of shared memory stays with the
      // mechanism tables (dodecane_lu +5 % against L1-cached accesses) and the 23 KB per warp stay L2 resident (2.1 KB of
      // DRAM traffic per cell against 1.3 KB of state, profiles/r02r_dram_modes.txt; the streaming .cs form of round 1
      // measures the same speed); small ones keep the saved Jacobian L1 resident (drm19 -4 % if it bypasses L1)
      for (int e = lane; e < n * ld; e += 32) {
        if (L::BIG) __stcg(Jsave + e, A[e]);
        else Jsave[e] = A[e];
      }
      nje++;
      nstlj = nst;
    } else {
      jcur = 0;
      for (int e = lane; e < n * ld; e += 32) A[e] = L::BIG ? __ldcg(Jsave + e) : Jsave[e];
    }
    __syncwarp();
    const double c = -gamma;
    for (int e = lane; e < n * ld; e += 32) {
      const int i = e / ld, j = e - i * ld;
      double a = A[e] * c;
      if (i == j) a += 1.0;
      A[e] = a;
    }
    __syncwarp();
#if !defined(PLM_GJ2) || PLM_GJ2
    int ier;
    if constexpr (NSLOT == 2 && L::PAIRS) ier = gjinv2_warp<L>(A);
    else ier = gjinv_warp<L>(A);
#else
    const int ier = gjinv_warp<L>(A);
#endif
    if constexpr (L::POOL > 0) {
      // W to the warp's global buffer, transposed (rows beyond n: zero), and the matrix back to the pool
      for (int j = 0; j < n; ++j) {
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
          const int i = lane + 32 * s;
          __stcg(Wt + j * L::WLD + i, (i < n) ? A[i * ld + j] : 0.0);
        }
      }
      pool_release<L>(pool);
    }
    nsetups++;
    gamrat = 1.0;
    gammap = gamma;
    crate = 1.0;
    nstlp = nst;
    return ier > 0 ? cvc::NLS_CONV_RECVR : cvc::NLS_SUCCESS;
  }

  // cvNlsResidual: res = rl1*zn[1] + acor - gamma*f(tn, zn[0]+acor)   (into dl)
  __device__ __forceinline__ void residual()
  {
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) yv[s] = valid(s) ? zn(0, s) + acor[s] : 0.0;
    f(tn, yv, ft);
#pragma unroll
    for (int s = 0; s < NSLOT; ++s)
      if (valid(s)) {
        const double r = rl1 * zn(1, s) + acor[s];
        dl[s] = -gamma * ft[s] + r;
      } else {
        dl[s] = 0.0;
      }
  }

  __device__ __forceinline__ int newton(bool callSetup)
  {
    bool jbadflag = false;
    int nls_jcur = 0, retval;
#pragma unroll 1
    for (;;) {
      bool need_res = true;
      retval = cvc::NLS_SUCCESS;
      int mit = 0;
      bool setup_failed = false;
#pragma unroll 1
      for (;;) {
        if (need_res) residual();  // single call site of the residual / right-hand side inside the Newton loop
        if (mit == 0 && callSetup) {
          if (jbadflag) convfail = cvc::FAIL_BAD_J;
          retval = lsetup();
          nls_jcur = jcur;
          callSetup = false;
          if (retval != cvc::NLS_SUCCESS) {
            setup_failed = true;
            break;
          }
        }
        nni++;
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) wb[L::FB + lane + 32 * s] = -dl[s];
        solve_warp<L>(wb, lane, dl, Wt);
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) dl[s] = valid(s) ? dl[s] : 0.0;
        if (gamrat != 1.0) {
          const double sc = fdiv(2.0, 1.0 + gamrat);
#pragma unroll
          for (int s = 0; s < NSLOT; ++s) dl[s] *= sc;
        }
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) acor[s] = valid(s) ? acor[s] + dl[s] : 0.0;
        const double del = wrms(dl, ewt);
        if (mit > 0) crate = fmax(cvc::CRDOWN * crate, fdiv(del, delp));
        const double dcon = fdiv(del * fmin(1.0, crate), tq[4]);
        if (dcon <= 1.0) {
          acnrm = (mit == 0) ? del : wrms(acor, ewt);
          return cvc::NLS_SUCCESS;
        }
        if ((mit >= 1) && (del > cvc::RDIV * delp)) {
          retval = cvc::NLS_CONV_RECVR;
          break;
        }
        delp = del;
        mit++;
        if (mit >= o.maxcor) {
          retval = cvc::NLS_CONV_RECVR;
          break;
        }
        need_res = true;
      }
      if (setup_failed) return retval;
      if ((retval > 0) && !nls_jcur) {
        callSetup = true;
        jbadflag = true;
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) acor[s] = 0.0;
        continue;
      }
      break;
    }
    return retval;
  }

  __device__ __forceinline__ void rescale()
  {
    double fac = eta;
#pragma unroll
    for (int j = 1; j <= QM; j++) {
      if (j <= q) {
#pragma unroll
        for (int s = 0; s < NSLOT; ++s)
          if (valid(s)) zn(j, s) *= fac;
        fac = eta * fac;
      }
    }
    h = hscale * eta;
    hscale = h;
  }

  __device__ __forceinline__ void increase_bdf()
  {
    double alpha0, alpha1, prod, xi, xiold, hsum;
#pragma unroll
    for (int i = 0; i <= QM; i++) lc[i] = 0.0;
    lc[2] = alpha1 = prod = xiold = 1.0;
    alpha0 = -1.0;
    hsum = hscale;
#pragma unroll
    for (int j = 1; j < QM; j++) {
      if (j < q) {
        hsum += tau[j + 1];
        xi = hsum / hscale;
        prod *= xi;
        alpha0 -= 1.0 / (j + 1);
        alpha1 += 1.0 / xi;
#pragma unroll
        for (int i = j + 2; i >= 2; i--)
          if (i <= QM) lc[i] = lc[i] * xiold + lc[i - 1];
        xiold = xi;
      }
    }
    const double A1 = (-alpha0 - alpha1) / prod;
    // zn[L] = A1 * zn[indx_acor]; indx_acor == qmax always (cvode.c:3192)
#pragma unroll
    for (int jl = 2; jl <= QM; ++jl) {
      if (jl == L_) {
#pragma unroll
        for (int s = 0; s < NSLOT; ++s)
          if (valid(s)) zn(jl, s) = A1 * wb[L::ZN + indx_acor * NP + lane + 32 * s];
      }
    }
#pragma unroll
    for (int j = 2; j <= QM; j++) {
      if (j <= q) {
#pragma unroll
        for (int s = 0; s < NSLOT; ++s)
          if (valid(s)) zn(j, s) = lc[j] * wb[L::ZN + L_ * NP + lane + 32 * s] + zn(j, s);
      }
    }
  }
  __device__ __forceinline__ void decrease_bdf()
  {
    double hsum, xi;
#pragma unroll
    for (int i = 0; i <= QM; i++) lc[i] = 0.0;
    lc[2] = 1.0;
    hsum = 0.0;
#pragma unroll
    for (int j = 1; j <= QM - 2; j++) {
      if (j <= q - 2) {
        hsum += tau[j];
        xi = hsum / hscale;
#pragma unroll
        for (int i = j + 2; i >= 2; i--) lc[i] = lc[i] * xi + lc[i - 1];
      }
    }
#pragma unroll
    for (int j = 2; j < QM; j++) {
      if (j < q) {
#pragma unroll
        for (int s = 0; s < NSLOT; ++s)
          if (valid(s)) zn(j, s) = -lc[j] * wb[L::ZN + q * NP + lane + 32 * s] + zn(j, s);
      }
    }
  }
  __device__ __forceinline__ void adjust_order(int deltaq)
  {
    if ((q == 2) && (deltaq != 1)) return;
    if (deltaq == 1) increase_bdf();
    else if (deltaq == -1) decrease_bdf();
  }
  __device__ __forceinline__ void adjust_params()
  {
    if (qprime != q) {
      adjust_order(qprime - q);
      q = qprime;
      L_ = q + 1;
      qwait = L_;
    }
    rescale();
  }
  template <int SIGN>
  __device__ __forceinline__ void pascal()
  {
#pragma unroll
    for (int k = 1; k <= QM; k++) {
      if (k <= q) {
#pragma unroll
        for (int j = QM; j >= k; j--) {
          if (j <= q) {
#pragma unroll
            for (int s = 0; s < NSLOT; ++s)
              if (valid(s)) zn(j - 1, s) = (SIGN > 0) ? zn(j - 1, s) + zn(j, s) : zn(j - 1, s) - zn(j, s);
          }
        }
      }
    }
  }
  // cvPredict: the new tn is handed to set_bdf, which stores it with its own group of scalars (one pair of warp barriers
  // for both; nothing reads tn in between)
  __device__ __forceinline__ double predict()
  {
    const double tn_new = tn + h;
    pascal<1>();
    return tn_new;
  }
  __device__ __forceinline__ void restore(double saved_t)
  {
    tn = saved_t;
    pascal<-1>();
  }
  // cvSetBDF + cvSetTqBDF (coefficients formed in registers, stored once). Reusing the previous step's values when
  // (q, h, tau, qwait == 1) are unchanged was tried and measured slower (the key compare costs more than it saves).
  __device__ __forceinline__ void set_bdf(double tn_new)
  {
    const int q_ = q;
    const double h_ = h;
    const bool qw1 = (qwait == 1);
    {
      double l[QM + 1];
      double alpha0, alpha0_hat, xi_inv, xistar_inv, hsum;
      l[0] = l[1] = xi_inv = xistar_inv = 1.0;
#pragma unroll
      for (int i = 2; i <= QM; i++) l[i] = 0.0;
      alpha0 = alpha0_hat = -1.0;
      hsum = h_;
      if (q_ > 1) {
#pragma unroll
        for (int j = 2; j < QM; j++) {
          if (j < q_) {
            hsum += tau[j - 1];
            xi_inv = fdiv(h_, hsum);
            alpha0 -= 1.0 / j;
#pragma unroll
            for (int i = j; i >= 1; i--) l[i] += l[i - 1] * xi_inv;
          }
        }
        alpha0 -= inv_int(q_);
        xistar_inv = -l[1] - alpha0;
        hsum += tau[q_ - 1];
        xi_inv = fdiv(h_, hsum);
        alpha0_hat = -l[1] - xi_inv;
#pragma unroll
        for (int i = QM; i >= 1; i--)
          if (i <= q_) l[i] += l[i - 1] * xistar_inv;
      }
      const double lq = sel(l, q_);
      const double A1 = 1.0 - alpha0_hat + alpha0;
      const double A2 = 1.0 + q_ * A1;
      const double tq2 = fabs(fdiv(A1, alpha0 * A2));
      const double tq5 = fabs(fdiv(A2 * xistar_inv, lq * xi_inv));
      double tq1 = 0.0, tq3 = 0.0;
      if (qw1) {
        if (q_ > 1) {
          const double C = fdiv(xistar_inv, lq);
          const double A3 = alpha0 + inv_int(q_);
          const double A4 = alpha0_hat + xi_inv;
          const double Cpinv = fdiv(1.0 - A4 + A3, A3);
          tq1 = fabs(C * Cpinv);
        } else
          tq1 = 1.0;
        hsum += tau[q_];
        xi_inv = fdiv(h_, hsum);
        const double A5 = alpha0 - inv_int(q_ + 1);
        const double A6 = alpha0_hat - xi_inv;
        const double Cppinv = fdiv(1.0 - A6 + A5, A2);
        tq3 = fabs(fdiv(Cppinv, xi_inv * (q_ + 2) * A5));
      }
      const double rl1v = frcp(l[1]);
      const double g = h_ * rl1v;
      const bool first = (nst == 0);
      const double gp = first ? g : double(gammap);
      const double gr = first ? 1.0 : fdiv(g, gp);
      // everything above is in registers: one group of stores, one pair of warp barriers
      __syncwarp();
      tq[2].put(tq2);
      tq[5].put(tq5);
      if (qw1) {
        tq[1].put(tq1);
        tq[3].put(tq3);
      }
      tq[4].put(fdiv(cvc::CORTES, tq2));
#pragma unroll
      for (int i = 0; i <= QM; i++) lc[i].put(l[i]);
      rl1.put(rl1v);
      gamma.put(g);
      if (first) gammap.put(g);
      gamrat.put(gr);
      tn.put(tn_new);
      __syncwarp();
    }
  }
  __device__ __forceinline__ int nls(int nflag)
  {
    // cv_jcur = SUNFALSE belongs behind a successful solve (cvode_nls.c); it is only ever read right after a setup has set
    // it, so clearing it here, in one store group with convfail, is the same thing and one pair of warp barriers less
    __syncwarp();
    convfail.put(((nflag == cvc::FIRST_CALL) || (nflag == cvc::PREV_ERR_FAIL)) ? cvc::NO_FAILURES : cvc::FAIL_OTHER);
    jcur.put(0);
    __syncwarp();
    const bool callSetup = (nflag == cvc::PREV_CONV_FAIL) || (nflag == cvc::PREV_ERR_FAIL) || (nst == 0) ||
                           (nst >= nstlp + cvc::MSBP) || (fabs(gamrat - 1.0) > cvc::DGMAX_LSETUP);
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) acor[s] = 0.0;
    const int flag = newton(callSetup);
    if (flag != cvc::NLS_SUCCESS) return flag;
    return 0;
  }

  __device__ __forceinline__ int handle_nflag(int& nflag, double saved_t, int& ncf)
  {
    if (nflag == 0) return cvc::DO_ERROR_TEST;
    ncfn++;
    restore(saved_t);
    ncf++;
    etamax = 1.0;
    if (ncf == cvc::MXNCF) return cvc::CV_CONV_FAILURE;
    eta = cvc::ETA_CF;
    nflag = cvc::PREV_CONV_FAIL;
    rescale();
    return cvc::PREDICT_AGAIN;
  }

  __device__ __forceinline__ int do_error_test(int& nflag, double saved_t, int& nef, double& dsm)
  {
    dsm = acnrm * tq[2];
    if (dsm <= 1.0) return 0;
    nef++;
    netf++;
    nflag = cvc::PREV_ERR_FAIL;
    restore(saved_t);
    if (nef == o.maxnef) return cvc::CV_ERR_FAILURE;
    etamax = 1.0;
    if (nef <= cvc::MXNEF1) {
      eta = 1.0 / (rroot(cvc::BIAS2 * dsm, L_) + cvc::ADDON);
      eta = fmax(cvc::ETA_MIN_EF, eta);
      if (nef >= cvc::SMALL_NEF) eta = fmin(eta, cvc::ETA_MAX_EF);
      rescale();
      return cvc::TRY_AGAIN;
    }
    if (q > 1) {
      eta = cvc::ETA_MIN_EF;
      adjust_order(-1);
      L_ = q;
      q--;
      qwait = L_;
      rescale();
      return cvc::TRY_AGAIN;
    }
    eta = cvc::ETA_MIN_EF;
    h *= eta;
    hscale = h;
    qwait = cvc::LONG_WAIT;
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) yv[s] = valid(s) ? zn(0, s) : 0.0;
    f(tn, yv, tv);
#pragma unroll
    for (int s = 0; s < NSLOT; ++s)
      if (valid(s)) zn(1, s) = h * tv[s];
    return cvc::TRY_AGAIN;
  }

  // cvCompleteStep + cvPrepareNextStep + cvSetEta (+ the new etamax, which cvStep sets right after).  Every new scalar is
  // formed in registers -- the later decisions read the NEW nst / qwait from those registers, as the reference reads them
  // from memory -- and everything is stored in ONE group between one pair of warp barriers.  Measured one change at a time
  // (profiles/r02v_groups_ab.txt, drm19 2^18 cells): this merge 244.0 k -> 247.3 k cells*dt/s; moving convfail / jcur into
  // set_bdf's group 247.2 k, acnrm in a register 246.2 k, rescale / setup-tail groups 243.7 k -- but all four together
  // 235.7 k: at this point the kernel answers to code layout (its per-step code is 34-37 KB against a 32 KB instruction
  // cache), not to the barrier count.  On top of this merge (profiles/r02y_groups_ab2.txt): acnrm in
  // a register 249.9 k, convfail / jcur in set_bdf's group 245.7 k, both 246.3 k -- so this one and acnrm are in.
  __device__ __forceinline__ void complete_and_prepare(double dsm)
  {
    const int q_ = q, nst1 = nst + 1, qwait1 = qwait - 1, L0 = L_;
    const double h_ = h, etamax_ = etamax;
    double told[QM + 2];
#pragma unroll
    for (int i = 1; i <= QM; i++) told[i] = tau[i];
    // cvCompleteStep: zn += l * acor; the correction kept for a possible order increase
#pragma unroll
    for (int j = 0; j <= QM; j++) {
      if (j <= q_) {
        const double lj = lc[j];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s)
          if (valid(s)) zn(j, s) = lj * acor[s] + zn(j, s);
      }
    }
    const bool keep = (qwait1 == 1) && (q_ != o.qmax);
    const double tq5_ = tq[5];
    if (keep) {
#pragma unroll
      for (int s = 0; s < NSLOT; ++s)
        if (valid(s)) wb[L::ZN + o.qmax * NP + lane + 32 * s] = acor[s];
    }
    // cvPrepareNextStep / cvSetEta
    int qwait_ = qwait1, qprime_ = q_;
    double eta_ = 1.0, hprime_ = h_;
    if (etamax_ == 1.0) {
      qwait_ = max(qwait_, 2);
    } else {
      const double etaq = frcp(rroot(cvc::BIAS2 * dsm, L0) + cvc::ADDON);
      if (qwait_ != 0) {
        eta_ = etaq;
      } else {
        qwait_ = 2;
        // choose_order reads tau[2], saved_tq5 as they are after cvCompleteStep: the shifted tau[2] is told[1] (or h for q == 1)
        choose_order(etaq, eta_, qprime_, (q_ >= 2) ? told[1] : ((nst1 > 1) ? told[1] : tau[2]), keep ? tq5_ : double(saved_tq5));
      }
      if ((eta_ > cvc::ETA_MIN_FX) && (eta_ < cvc::ETA_MAX_FX)) {
        eta_ = 1.0;
      } else {
        if (eta_ >= cvc::ETA_MAX_FX) eta_ = fmin(eta_, etamax_);
        else eta_ = fmax(eta_, cvc::ETA_MIN);
        hprime_ = h_ * eta_;
      }
    }
    const double etamax_new = (nst1 <= cvc::SMALL_NST) ? cvc::ETA_MAX_ES : cvc::ETA_MAX_GS;
    __syncwarp();
    nst.put(nst1);
#pragma unroll
    for (int i = QM; i >= 2; i--)
      if (i <= q_) tau[i].put(told[i - 1]);
    if ((q_ == 1) && (nst1 > 1)) tau[2].put(told[1]);
    tau[1].put(h_);
    if (keep) {
      saved_tq5.put(tq5_);
      indx_acor.put(o.qmax);
    }
    eta.put(eta_);
    qprime.put(qprime_);
    hprime.put(hprime_);
    qwait.put(qwait_);
    etamax.put(etamax_new);
    __syncwarp();
  }

  // cvPrepareNextStep with qwait == 0: cvComputeEtaqm1 / cvComputeEtaqp1 / cvChooseEta (no scalar is stored here)
  // (tau2 / stq5: tau[2] and saved_tq5 as cvCompleteStep leaves them -- the caller has not stored them yet)
  __device__ __forceinline__ void choose_order(const double etaq, double& eta_out, int& qprime_out, const double tau2,
                                               const double stq5)
  {
    const int q_ = q;
    double etaqm1 = 0.0, etaqp1 = 0.0;
    if (q_ > 1) {
      double a = 0.0;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s)
        if (valid(s)) {
          const double p = wb[L::ZN + q_ * NP + lane + 32 * s] * ewt[s];
          a += p * p;
        }
      const double ddn = wrms_finish(warp_sum(a)) * tq[1];
      etaqm1 = frcp(rroot(cvc::BIAS1 * ddn, q_) + cvc::ADDON);
    }
    if (q_ != o.qmax && stq5 != 0.0) {
      const double cquot = fdiv(tq[5], stq5) * rpowI(fdiv(h, tau2), L_);
#pragma unroll
      for (int s = 0; s < NSLOT; ++s)
        tv[s] = valid(s) ? -cquot * wb[L::ZN + o.qmax * NP + lane + 32 * s] + acor[s] : 0.0;
      const double dup = wrms(tv, ewt) * tq[3];
      etaqp1 = frcp(rroot(cvc::BIAS3 * dup, L_ + 1) + cvc::ADDON);
    }
    const double etam = fmax(etaqm1, fmax(etaq, etaqp1));
    if ((etam > cvc::ETA_MIN_FX) && (etam < cvc::ETA_MAX_FX)) {
      eta_out = 1.0;
      qprime_out = q_;
    } else if (etam == etaq) {
      eta_out = etaq;
      qprime_out = q_;
    } else if (etam == etaqm1) {
      eta_out = etaqm1;
      qprime_out = q_ - 1;
    } else {
      eta_out = etaqp1;
      qprime_out = q_ + 1;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s)
        if (valid(s)) wb[L::ZN + o.qmax * NP + lane + 32 * s] = acor[s];
    }
  }

  __device__ __forceinline__ int step()
  {
    int ncf = 0, nef = 0, nflag, kflag, eflag;
    double dsm = 0.0;
    if ((nst > 0) && (hprime != h)) adjust_params();
    const double saved_t = tn;
    nflag = cvc::FIRST_CALL;
#pragma unroll 1
    for (;;) {
      set_bdf(predict());
      nflag = nls(nflag);
      kflag = handle_nflag(nflag, saved_t, ncf);
      if (kflag == cvc::PREDICT_AGAIN) continue;
      if (kflag != cvc::DO_ERROR_TEST) return kflag;
      eflag = do_error_test(nflag, saved_t, nef, dsm);
      if (eflag == cvc::TRY_AGAIN) continue;
      if (eflag != 0) return eflag;
      break;
    }
    complete_and_prepare(dsm);
    const double t2 = tq[2];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) acor[s] *= t2;
    return 0;
  }

  __device__ __forceinline__ int hin(double tout)
  {
    const double uround = DBL_EPSILON;
    const double tdiff = tout - tn;
    if (tdiff == 0.0) return cvc::CV_TOO_CLOSE;
    const int sign = (tdiff > 0.0) ? 1 : -1;
    const double tdist = fabs(tdiff);
    const double tround = uround * fmax(fabs(tn), fabs(tout));
    if (tdist < 2.0 * tround) return cvc::CV_TOO_CLOSE;
    const double hlb = cvc::HLB_FACTOR * tround;
    // cvUpperBoundH0
    double hub_inv_l = 0.0;
#pragma unroll
    for (int s = 0; s < NSLOT; ++s)
      if (valid(s)) {
        const double t1 = cvc::HUB_FACTOR * fabs(zn(0, s)) + 1.0 / ewt[s];
        hub_inv_l = fmax(hub_inv_l, fabs(zn(1, s)) / t1);
      }
    const double hub_inv = warp_max(hub_inv_l);
    double hub = cvc::HUB_FACTOR * tdist;
    if (hub * hub_inv > 1.0) hub = 1.0 / hub_inv;
    double hg = sqrt(hlb * hub);
    if (hub < hlb) {
      h = (sign == -1) ? -hg : hg;
      return 0;
    }
    double hnew = hg;
#pragma unroll 1
    for (int count1 = 1; count1 <= cvc::MAX_ITERS; count1++) {
      const double hgs = hg * sign;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) yv[s] = valid(s) ? hgs * zn(1, s) + zn(0, s) : 0.0;
      f(tn + hgs, yv, tv);
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) tv[s] = valid(s) ? (1.0 / hgs) * (tv[s] - zn(1, s)) : 0.0;
      const double yddnrm = wrms(tv, ewt);
      hnew = (yddnrm * hub * hub > 2.0) ? sqrt(2.0 / yddnrm) : sqrt(hg * hub);
      if (count1 == cvc::MAX_ITERS) break;
      const double hrat = hnew / hg;
      if ((hrat > 0.5) && (hrat < 2.0)) break;
      if ((count1 > 1) && (hrat > 2.0)) {
        hnew = hg;
        break;
      }
      hg = hnew;
    }
    double h0 = cvc::H_BIAS * hnew;
    if (h0 < hlb) h0 = hlb;
    if (h0 > hub) h0 = hub;
    if (sign == -1) h0 = -h0;
    h = h0;
    return 0;
  }

  // CVodeReInit + CVode(tout, CV_NORMAL). y0 in / result out through yio (lane components).
  __device__ __forceinline__ int integrate(double (&yio)[NSLOT], double t0, double tout, CellStats& st)
  {
    tn = t0;
    q = 1;
    L_ = 2;
    qwait = L_;
    etamax = cvc::ETA_MAX_FS;
    nst = nstlp = nstlj = 0;
    nfe.reset(); ncfn.reset(); netf.reset(); nni.reset(); nsetups.reset(); nje.reset();
    saved_tq5 = 0.0;
    indx_acor = o.qmax;
    crate = 1.0;
    delp = 0.0;
    jcur = 0;
    convfail = 0;
    eta = 1.0;
    qprime = 1;
    gammap = gamma = gamrat = rl1 = 1.0;
    acnrm = 0.0;
#pragma unroll
    for (int i = 0; i <= QM; ++i) tau[i] = lc[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) tq[i] = 0.0;
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      acor[s] = 0.0;
#pragma unroll
      for (int j = 0; j <= QM; ++j) zn(j, s) = 0.0;
      if (valid(s)) zn(0, s) = yio[s];
    }
    int istate = 0;
    if (!ewtset()) {
      istate = cvc::CV_ILL_INPUT;
    } else {
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) yv[s] = valid(s) ? zn(0, s) : 0.0;
      f(tn, yv, tv);
#pragma unroll
      for (int s = 0; s < NSLOT; ++s)
        if (valid(s)) zn(1, s) = tv[s];
      istate = hin(tout);
    }
    if (istate == 0) {
      hscale = h;
      hprime = h;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s)
        if (valid(s)) zn(1, s) *= h;
      int nstloc = 0;
#pragma unroll 1
      for (;;) {
        if (nst > 0 && !ewtset()) {
          istate = cvc::CV_ILL_INPUT;
          break;
        }
        if (nstloc >= o.mxstep) {
          istate = cvc::CV_TOO_MUCH_WORK;
          break;
        }
        // uround * wrms(zn[0], ewt) > 1 (cvode.c:1316-1321). Each term |y_i| / (rtol |y_i| + atol_i) is below 1 / rtol, so
        // the test cannot fire unless rtol is within a few ulp of uround: the reduction is skipped otherwise.
        if (o.rtol < 1e-15 && wss_zn0() > double(n) / (DBL_EPSILON * DBL_EPSILON)) {
          istate = cvc::CV_TOO_MUCH_ACC;
          break;
        }
        const int kflag = step();
        if (kflag != 0) {
          istate = kflag;
          break;
        }
        nstloc++;
        if ((tn - tout) * h >= 0.0) break;
      }
    }
    if (istate == 0) {
      // CVodeGetDky(tout, 0): sum_{j=q..0} s^j zn[j]
      const double sfac = (tout - tn) / h;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        if (valid(s)) {
          double z = 0.0;
#pragma unroll
          for (int j = QM; j >= 0; j--) {
            if (j <= q) {
              double c = 1.0;
              for (int i = 0; i < j; i++) c *= sfac;
              z = (j == q) ? c * zn(j, s) : z + c * zn(j, s);
            }
          }
          yio[s] = z;
        }
      }
    } else {
#pragma unroll
      for (int s = 0; s < NSLOT; ++s)
        if (valid(s)) yio[s] = zn(0, s);
    }
    __syncwarp();  // lane 0's counters
    st.nst = nst;
    st.nfe = nfe.final_value();
    st.nje = nje.final_value();
    st.nsetups = nsetups.final_value();
    st.nni = nni.final_value();
    st.ncfn = ncfn.final_value();
    st.netf = netf.final_value();
    st.flag = istate;
    return istate;
  }
};

}  // namespace plm
