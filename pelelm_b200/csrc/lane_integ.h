// Lane = cell integrators: one THREAD integrates one cell, a warp advances 32 cells in lock step through ROUNDS.
//
// A round is: (A) one right-hand side for every lane that needs one (straight-line generated code, 32 cells per
// instruction), (B) the per-lane bookkeeping that consumes it, (C) -- BDF only -- a linear-solver setup (analytic Jacobian,
// I - gamma J, dense LU) for the lanes that are waiting for one, run when enough of them wait, (D) -- BDF only -- one
// Newton correction (dense triangular solves, fully unrolled, the right-hand side vector in registers) plus everything
// CVODE does between two right-hand sides for the lanes that have a factorisation.  Every integrator decision stays per
// cell (own step size, order, Newton count, Jacobian age), so the arithmetic per cell is that of the reference's CPU
// reactor; what is shared by the 32 cells of a warp is only the instruction stream.  Each integrator is written as a
// state machine over the points where the reference's code calls the right-hand side.
//
//   LaneRK64  ReactorRK64::react        PelePhysics/Reactions/ReactorRK64.cpp:206-317 (coefficients ReactorRK64.H:11-42)
//   LaneBdf   CVodeReInit + CVode(CV_NORMAL) as ReactorCvode drives them (ReactorCvode.cpp:1372-1453), SUNDIALS 6.2.0:
//             CVode cvode/cvode.c:990-1476, cvHin :1945-2048, cvStep :2143-2249, cvAdjustParams/Order :2265-2451,
//             cvRescale :2457, cvPredict :2485, cvSetBDF/cvSetTqBDF :2691-2779, cvNls :2781, cvHandleNFlag :2954,
//             cvRestore :3008, cvDoErrorTest :3048, cvCompleteStep :3162, cvPrepareNextStep :3218, cvSetEta :3262,
//             cvComputeEtaqm1/qp1 :3299/:3318, cvChooseEta :3347, CVodeGetDky :1490; Newton iteration
//             sunnonlinsol/newton/sunnonlinsol_newton.c:187-335, cvNlsLSetup/ConvTest/Residual cvode/cvode_nls.c:251-388,
//             cvLsSetup/LinSys/Solve cvode/cvode_ls.c:1270-1720, denseGETRF/GETRS sundials/sundials_dense.c:135-232,
//             N_VWrmsNorm nvector/serial/nvector_serial.c:638-660 (serial sums, in index order, as there).
//
// Workspace: see struct Ws -- the vectors in shared memory, the matrices in global memory, both interleaved over the 32
// lanes of the warp.
#pragma once
#include "lane_chem.h"
#include "lane_tab.h"

namespace plm {
namespace lane {

constexpr int W = PLM_LANE_W;

// The warp's workspace; both parts interleave the 32 lanes (element e of lane l at base[e * 32 + l]) and the pointers
// are already offset by the lane:
//   d(e)  FAST space -- the vectors every round touches (Nordsieck array, weights, correction, state, right-hand side):
//         shared memory on the device (conflict-free for any per-lane element index)
//   g(e)  BIG space  -- LU factors, saved Jacobian, pivots, external source: global memory (one fully used 256-byte line
//         per access of the warp; the factors of the resident cells are what the L2 is for)
struct Ws {
  double* fast;
  double* big;
  PLM_LHD double& d(int e) const { return fast[size_t(e) * W]; }
  PLM_LHD double& g(int e) const { return big[size_t(e) * W]; }
};

struct LaneOpts {
  double rtol, atol;
  int qmax, maxcor, use_typ;
  int mxstep, maxnef, msbj;
  int rk64_guess, rk64_min, rk64_max;
  const double* typ_vals;  // [N], read when a cell is loaded (use_typ != 0)
};

struct LaneStats {
  int nst, nfe, nje, nsetups, nni, ncfn, netf, flag;
};

enum Phase : int {
  PH_IDLE = 0,     // no cell
  PH_LOAD,         // T-solve only: box_flatten's RHY2T on the loaded state
  PH_RHS0,         // BDF: f(t0, y0)         RK64: a stage
  PH_HIN,          // BDF: cvHin trial
  PH_NEWTON,       // BDF: residual of the Newton iteration
  PH_ERRF,         // BDF: f(tn, zn0) after repeated error-test failures at order 1
  PH_WAIT_SETUP,   // BDF: residual evaluated, linear-solver setup wanted
  PH_SOLVE,        // BDF: ready for the Newton correction (transient inside a round)
  PH_OUT,          // integration over: result in Y, caller prepares the store (transient)
  PH_FINISH,       // T-solve only: box_unflatten's RHY2T on the result
  PH_DONE          // caller stores the cell and takes the next one (transient)
};

// per-cell source terms read from the workspace by the right-hand side
struct WsSrc {
  Ws w;
  int off;
  double e;
  PLM_LHD double operator()(int k) const { return w.g(off + k); }
  PLM_LHD double energy() const { return e; }
};

template <class M, int QM>
struct LaneLay {
  static constexpr int N = M::N, NS = M::NS;
  // ---- FAST space (doubles per cell) ----
  static constexpr int ZN = 0;                        // Nordsieck array [(QM+1)][N]   (RK64: soln, carry, err)
  static constexpr int EWT = ZN + (QM + 1) * N;
  static constexpr int ACOR = EWT + N;
  static constexpr int Y = ACOR + N;                  // argument of the right-hand side / result of the cell
  // The straight-line right-hand side reads its argument into registers before it writes its result, and nothing reads the
  // argument afterwards (a Jacobian is only ever evaluated at the first residual of a Newton attempt, where y = zn0): the
  // result overwrites it.  The loop form reads y while it writes f.
  static constexpr int F = M::kTable ? Y + N : Y;     // last right-hand side; then the Newton right-hand side / correction,
  static constexpr int DEL = F;                       // Newton right-hand side / correction (in place)
  static constexpr int SC = F + N;                    // table-driven kernels: concentrations [NS] and g/RT [NS] of the
  static constexpr int G = SC + NS;                   // right-hand side / Jacobian being evaluated
  static constexpr int SCAL = M::kTable ? G + NS : F + N;   // cvode's tau[QM+3], tq[6], l[QM+3]
  static constexpr int FAST_BDF = SCAL + (QM + 1) + 5 + (QM + 2);
  static constexpr int RK_F = 3 * N, RK_SC = RK_F + N, RK_G = RK_SC + NS;   // RK64: soln, carry, err, (f, sc, g)
  static constexpr int FAST_RK = M::kTable ? RK_G + NS : 3 * N;
  // ---- BIG space ----
  static constexpr int SRC = 0;                       // external species source [NS]
  static constexpr int PERM = SRC + NS;               // pivot rows [N] (stored as doubles)
  static constexpr int MAT = PERM + N;                // LU factors of I - gamma J, column-major [N][N]
  static constexpr int SAV = MAT + N * N;             // saved Jacobian
  static constexpr int JSCR = SAV + N * N;            // table-driven Jacobian: h/RT, wdot, d wdot/dT, cp/R [4][NS]
  static constexpr int BIG_BDF = JSCR + (M::kTable ? 4 * NS : 0);
  static constexpr int BIG_RK = SRC + NS;
  // unroll factor of the vector loops: a divisor of N
  static constexpr int VU = (N % 11 == 0) ? 11 : ((N % 5 == 0) ? 5 : ((N % 6 == 0) ? 6 : 2));
};

// ---------------------------------------------------------------------------------------------------------------
// ReactorRK64
// ---------------------------------------------------------------------------------------------------------------
PLM_LCONST double c_rk64_lane[18] = {
    0.218150805229859,   0.256702469801519,  0.527402592007520,   0.0484864267224467, 1.24517071533530,   0.412366034843237,
    -0.113554138044166,  -0.215118587818400, -0.0510152146250577, -1.07992686223881,  -0.248664241213447, 0.0,
    -0.0554699315064507, 0.158481845574980,  -0.0905918835751907, -0.219084567203338, 0.164022338959433,  0.0426421977505659};

template <class M, int EINT>
struct LaneRK64 {
  using L = LaneLay<M, 2>;
  static constexpr int N = M::N, NS = M::NS;
  static constexpr int SOLN = L::ZN, CARRY = L::ZN + N, ERR = L::ZN + 2 * N;
  static constexpr bool kBdf = false;
  int phase = PH_IDLE;
  int stage, nsteps, nfe, flag, clamp;
  double time_init, time_out, current_time, dt_rk, dt_rk_min, dt_rk_max, atol;
  double rhoe_init, rhoesrc, hin_load;

  PLM_LHD bool rhs_active() const { return phase == PH_LOAD || phase == PH_RHS0 || phase == PH_FINISH; }
  PLM_LHD bool only_T() const { return phase == PH_LOAD || phase == PH_FINISH; }
  static constexpr int YV = SOLN;  // where the caller puts the loaded state and finds the result
  static constexpr int FV = L::RK_F, SCV = L::RK_SC, GV = L::RK_G;  // table-driven right-hand side: result and scratch
  PLM_LHD int rhs_vec() const { return SOLN; }
  PLM_LHD double hin_rho() const
  {
    return phase == PH_RHS0 ? rhoe_init + rhoesrc * (current_time - time_init) : hin_load;
  }
  // the caller has put the state into Y (species, T or T guess), SRC, rhoe_init, rhoesrc
  PLM_LHD void begin(const Ws& w, const LaneOpts& o, double time, double dt, bool flat)
  {
    time_init = time;
    time_out = time + dt;
    current_time = time;
    dt_rk = dt / double(o.rk64_guess);
    dt_rk_min = dt / double(o.rk64_max);
    dt_rk_max = dt / double(o.rk64_min);
    atol = o.atol;
    nsteps = 0;
    nfe = 0;
    flag = 0;
    stage = 0;
    hin_load = rhoe_init;
    // react(box) limits every new step to the time left (ReactorRK64.cpp:307, inside the loop); the 1dArray react applies
    // that limit only after its loop (:141-142), so its last step runs past time_out: reproduced
    clamp = flat ? 0 : 1;
    if (flat) start(w);
    else phase = PH_LOAD;
  }
  PLM_LHD void start(const Ws& w)
  {
#pragma unroll(L::VU)
    for (int i = 0; i < N; ++i) {
      w.d(CARRY + i) = w.d(SOLN + i);
      w.d(ERR + i) = 0.0;
    }
    phase = (current_time < time_out) ? PH_RHS0 : PH_OUT;
  }
  PLM_LHD void post_rhs(const Ws& w, const double (&f)[N], double T)
  {
    if (phase == PH_LOAD) {
      w.d(SOLN + NS) = T;
      start(w);
      return;
    }
    if (phase == PH_FINISH) {
      w.d(SOLN + NS) = T;
      phase = PH_DONE;
      return;
    }
    nfe++;
    const double ca = c_rk64_lane[stage], cb = c_rk64_lane[6 + stage], ce = c_rk64_lane[12 + stage];
    double me = 1e-50;  // tinyval
    const bool last = (stage == 5);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double e = w.d(ERR + i) + ce * dt_rk * f[i];
      const double s = w.d(CARRY + i) + ca * dt_rk * f[i];
      w.d(SOLN + i) = s;
      w.d(CARRY + i) = s + cb * dt_rk * f[i];
      w.d(ERR + i) = last ? 0.0 : e;
      me = lane_fabs(e) > me ? lane_fabs(e) : me;
    }
    if (!last) {
      stage++;
      return;
    }
    stage = 0;
    current_time += dt_rk;
    nsteps++;
    const double max_err = me;
    if (max_err < atol) {
      const double cf = lane_powc(atol / max_err, 0.25);
      dt_rk = (dt_rk_max < dt_rk * cf) ? dt_rk_max : dt_rk * cf;
    } else {
      const double cf = lane_powc(atol / max_err, 0.2);
      dt_rk = (dt_rk_min > dt_rk * cf) ? dt_rk_min : dt_rk * cf;
    }
    if (clamp) dt_rk = (dt_rk < time_out - current_time) ? dt_rk : time_out - current_time;
    if (!(current_time < time_out)) phase = PH_OUT;
  }
  PLM_LHD void stats(LaneStats& s) const
  {
    s = LaneStats{};
    s.nst = nsteps;
    s.nfe = nfe;
    s.flag = flag;
  }
  PLM_LHD int fc() const { return nsteps; }
  // nothing to set up or solve
  PLM_LHD bool wants_setup() const { return false; }
};

// ---------------------------------------------------------------------------------------------------------------
// ReactorCvode (BDF, Newton, dense LU, analytic Jacobian)
// ---------------------------------------------------------------------------------------------------------------
namespace cv {
constexpr double ETA_MIN_FX = 0.0, ETA_MAX_FX = 1.5, ETA_MAX_FS = 10000.0, ETA_MAX_ES = 10.0, ETA_MAX_GS = 10.0,
                 ETA_MIN = 0.1, ETA_MAX_EF = 0.2, ETA_MIN_EF = 0.1, ETA_CF = 0.25, ADDON = 0.000001, BIAS1 = 6.0,
                 BIAS2 = 6.0, BIAS3 = 10.0, HLB_FACTOR = 100.0, HUB_FACTOR = 0.1, H_BIAS = 0.5, CORTES = 0.1,
                 CRDOWN = 0.3, RDIV = 2.0, DGMAX_LSETUP = 0.3, DGMAX_JBAD = 0.2, UROUND = 2.220446049250313e-16;
constexpr int SMALL_NST = 10, SMALL_NEF = 2, LONG_WAIT = 10, MXNCF = 10, MXNEF1 = 3, MAX_ITERS = 4, MSBP = 20;
constexpr int CV_TOO_MUCH_WORK = -1, CV_TOO_MUCH_ACC = -2, CV_ERR_FAILURE = -3, CV_CONV_FAILURE = -4, CV_ILL_INPUT = -22,
              CV_TOO_CLOSE = -27;
constexpr int FIRST_CALL = 6, PREV_CONV_FAIL = 7, PREV_ERR_FAIL = 8;
constexpr int CV_NO_FAILURES = 0, CV_FAIL_BAD_J = 1, CV_FAIL_OTHER = 2;
}  // namespace cv

#ifdef LANE_MODEL_COUNT_SWAPS
static long lane_model_swaps = 0, lane_model_pivots = 0, lane_model_lus = 0, lane_model_lus_swapped = 0;
#endif

template <class M, int QM, int EINT>
struct LaneBdf {
  using L = LaneLay<M, QM>;
  static constexpr int N = M::N, NS = M::NS, VU = L::VU;
  static constexpr int NT = QM + 3;
  static constexpr bool kBdf = true;

  int phase = PH_IDLE;
  // ---- per-cell scalars (cvode_impl.h CVodeMemRec) ----
  double tn, h, hprime, hscale, eta, etamax, tout, t0;
  // cvode's tau[], tq[], l[] live in the FAST space (SCAL): as member arrays they are indexed dynamically, i.e. local
  // memory, and with the SM's shared memory given to the vectors local memory is served by the L2
  // (tau[1 .. QM+1], tq[1 .. 5] and l[0 .. QM+1] are the entries ever touched: the bases are shifted so that tau[1] and tq[1]
  // are the first doubles of their blocks)
  static constexpr int T_TAU = L::SCAL - 1, T_TQ = T_TAU + (QM + 1), T_L = T_TQ + 5 + 1;
  double rl1, gamma, gammap, gamrat, crate, delp, acnrm, saved_tq5, saved_t, dsm;
  double hg, hub, hlb, hsign;  // cvHin
  double rhoe_init, rhoesrc, hin_load, t_eval;
  double rtol, atol;
  int q, qprime, Lq, qwait, qmax, maxcor, mxstep, maxnef, msbj;
  int nst, nfe, ncfn, netf, nni, nsetups, nje, nstlp, nstlj, nstloc;
  int jcur, convfail, nflag, ncf, nef, mnewt, call_setup, jbad, nls_jcur, first_res, count1, flag;
  const double* typ;  // typical values (ReactorBase.cpp:105-123) or nullptr

  PLM_LHD bool rhs_active() const { return phase >= PH_LOAD && phase <= PH_ERRF || phase == PH_FINISH; }
  PLM_LHD bool only_T() const { return phase == PH_LOAD || phase == PH_FINISH; }
  static constexpr int YV = L::Y;  // where the caller puts the loaded state and finds the result
  static constexpr int FV = L::F, SCV = L::SC, GV = L::G;  // table-driven right-hand side: result and scratch
  PLM_LHD int rhs_vec() const { return L::Y; }
  PLM_LHD double hin_rho() const { return only_T() ? hin_load : rhoe_init + rhoesrc * t_eval; }
  PLM_LHD double atol_of(int i) const { return typ ? typ[i] * atol : atol; }
  PLM_LHD bool wants_setup() const { return phase == PH_WAIT_SETUP; }
  PLM_LHD int fc() const { return nfe; }
  PLM_LHD void stats(LaneStats& s) const
  {
    s.nst = nst; s.nfe = nfe; s.nje = nje; s.nsetups = nsetups; s.nni = nni; s.ncfn = ncfn; s.netf = netf; s.flag = flag;
  }

  PLM_LHD double& zn(const Ws& w, int j, int i) const { return w.d(L::ZN + j * N + i); }

  // N_VWrmsNorm of vector at offset a with the error weights
  PLM_LHD double wrms(const Ws& w, int a) const
  {
    double sum = 0.0;
#pragma unroll(VU)
    for (int i = 0; i < N; ++i) {
      const double p = w.d(a + i) * w.d(L::EWT + i);
      sum += p * p;
    }
    return sqrt(sum / N);
  }
  // cvEwtSet (scalar rtol, per-component atol: ReactorBase.cpp:105-123)
  PLM_LHD bool ewtset(const Ws& w) const
  {
    bool bad = false;
#pragma unroll(VU)
    for (int i = 0; i < N; ++i) {
      const double t = rtol * lane_fabs(zn(w, 0, i)) + atol_of(i);
      bad = bad || (t <= 0.0);
      w.d(L::EWT + i) = lane_rcp(t);
    }
    return !bad;
  }

  // the caller has put the state into Y (species, T or T guess), SRC, rhoe_init, rhoesrc
  PLM_LHD void begin(const Ws& w, const LaneOpts& o, double time, double dt, bool flat)
  {
    rtol = o.rtol; atol = o.atol; qmax = o.qmax; maxcor = o.maxcor; mxstep = o.mxstep; maxnef = o.maxnef; msbj = o.msbj;
    typ = o.use_typ ? o.typ_vals : nullptr;
    t0 = time;
    tout = time + dt;
    hin_load = rhoe_init;
    flag = 0;
    nst = nfe = ncfn = netf = nni = nsetups = nstlp = nje = nstlj = 0;
    if (flat) init(w);
    else phase = PH_LOAD;
  }
  // CVodeReInit + the first-call block of CVode up to f(t0, y0)
  PLM_LHD void init(const Ws& w)
  {
    tn = t0;
    q = 1; Lq = 2; qwait = Lq;
    etamax = cv::ETA_MAX_FS;
    saved_tq5 = 0.0;
    crate = 1.0; delp = 0.0; jcur = 0; eta = 1.0; qprime = 1;
#pragma unroll
    for (int i = 1; i <= QM + 1; ++i) w.d(T_TAU + i) = 0.0;
#pragma unroll
    for (int i = 0; i <= QM + 1; ++i) w.d(T_L + i) = 0.0;
#pragma unroll
    for (int i = 1; i < 6; ++i) w.d(T_TQ + i) = 0.0;
#pragma unroll(VU)
    for (int i = 0; i < N; ++i) zn(w, 0, i) = w.d(L::Y + i);
    if (!ewtset(w)) {  // CVode returns CV_ILL_INPUT before any work: the state is the input
      flag = cv::CV_ILL_INPUT;
      phase = PH_OUT;
      return;
    }
    t_eval = tn;
    phase = PH_RHS0;
  }
  // Deferred actions: the pieces of cvStep that several paths share (restore, order change, rescale, predict + cvSetBDF,
  // the first residual) exist ONCE, in tail(), and the paths only say which of them they need. Straight-line code that
  // 32 cells stream through is the whole point of the layout; five inlined copies of the predictor would undo it.
  enum : int { P_RESTORE = 1, P_FAIL = 2, P_ERRF = 4, P_ADJ = 8, P_RESCALE = 16, P_ATTEMPT = 32, P_BEGIN = 64, P_FIRSTRES = 128 };
  int pend = 0, adj = 0;

  // integration ends with an error: result = zn0 (after a pending restore)
  PLM_LHD void fail(int code)
  {
    flag = code;
    pend |= P_FAIL;
  }

  // ---- cvHin ----
  PLM_LHD void hin_trial(const Ws& w)
  {
    const double hgs = hg * hsign;
#pragma unroll(VU)
    for (int i = 0; i < N; ++i) w.d(L::Y + i) = hgs * zn(w, 1, i) + zn(w, 0, i);
    t_eval = tn + hgs;
    phase = PH_HIN;
  }
  // returns true when the trial loop is needed
  PLM_LHD bool hin_begin(const Ws& w)
  {
    const double tdiff = tout - tn;
    if (tdiff == 0.0) { fail(cv::CV_TOO_CLOSE); return false; }
    hsign = (tdiff > 0.0) ? 1.0 : -1.0;
    const double tdist = lane_fabs(tdiff);
    const double tround = cv::UROUND * ((lane_fabs(tn) > lane_fabs(tout)) ? lane_fabs(tn) : lane_fabs(tout));
    if (tdist < 2.0 * tround) { fail(cv::CV_TOO_CLOSE); return false; }
    hlb = cv::HLB_FACTOR * tround;
    // cvUpperBoundH0
    double hub_inv = 0.0;
#pragma unroll(VU)
    for (int i = 0; i < N; ++i) {
      const double t2 = lane_fabs(zn(w, 0, i));
      const double t1 = cv::HUB_FACTOR * t2 + 1.0 / w.d(L::EWT + i);
      const double r = lane_fabs(zn(w, 1, i)) / t1;
      if (r > hub_inv) hub_inv = r;
    }
    hub = cv::HUB_FACTOR * tdist;
    if (hub * hub_inv > 1.0) hub = 1.0 / hub_inv;
    hg = sqrt(hlb * hub);
    if (hub < hlb) {
      h = (hsign < 0.0) ? -hg : hg;
      start_steps(w);
      return false;
    }
    count1 = 1;
    return true;
  }
  // returns true when another trial is needed
  PLM_LHD bool hin_post(const Ws& w, const double (&f)[N])
  {
    const double hgs = hg * hsign;
    double sum = 0.0;
    const double hinv = 1.0 / hgs;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double tv = hinv * (f[i] - zn(w, 1, i));  // N_VLinearSum(1/hg, tempv, -1/hg, zn1) -> VScaleDiff
      const double p = tv * w.d(L::EWT + i);
      sum += p * p;
    }
    const double yddnrm = sqrt(sum / N);
    double hnew = (yddnrm * hub * hub > 2.0) ? sqrt(2.0 / yddnrm) : sqrt(hg * hub);
    bool done = (count1 == cv::MAX_ITERS);
    if (!done) {
      const double hrat = hnew / hg;
      if ((hrat > 0.5) && (hrat < 2.0)) done = true;
      else if ((count1 > 1) && (hrat > 2.0)) {
        hnew = hg;
        done = true;
      }
    }
    if (!done) {
      hg = hnew;
      count1++;
      return true;
    }
    double h0 = cv::H_BIAS * hnew;
    if (h0 < hlb) h0 = hlb;
    if (h0 > hub) h0 = hub;
    if (hsign < 0.0) h0 = -h0;
    h = h0;
    start_steps(w);
    return false;
  }
  PLM_LHD void start_steps(const Ws& w)
  {
    hscale = h;
    hprime = h;
#pragma unroll(VU)
    for (int i = 0; i < N; ++i) zn(w, 1, i) *= h;
    nstloc = 0;
    pend |= P_BEGIN;
  }

  PLM_LHD void rescale(const Ws& w)
  {
    double fac = eta;
    for (int j = 1; j <= q; j++) {
#pragma unroll(VU)
      for (int i = 0; i < N; ++i) zn(w, j, i) *= fac;
      fac = eta * fac;
    }
    h = hscale * eta;
    hscale = h;
  }
  PLM_LHD void increase_bdf(const Ws& w)
  {
    double alpha0, alpha1, prod, xi, xiold, hsum, A1;
    for (int i = 0; i <= qmax; i++) w.d(T_L + (i)) = 0.0;
    w.d(T_L + (2)) = alpha1 = prod = xiold = 1.0;
    alpha0 = -1.0;
    hsum = hscale;
    if (q > 1) {
      for (int j = 1; j < q; j++) {
        hsum += w.d(T_TAU + (j + 1));
        xi = hsum / hscale;
        prod *= xi;
        alpha0 -= 1.0 / (j + 1);
        alpha1 += 1.0 / xi;
        for (int i = j + 2; i >= 2; i--) w.d(T_L + (i)) = w.d(T_L + (i)) * xiold + w.d(T_L + (i - 1));
        xiold = xi;
      }
    }
    A1 = (-alpha0 - alpha1) / prod;
    // indx_acor == qmax always (cvCompleteStep stores acor there)
#pragma unroll(VU)
    for (int i = 0; i < N; ++i) zn(w, Lq, i) = A1 * zn(w, qmax, i);
    if (q > 1)
      for (int j = 2; j <= q; j++) {
#pragma unroll(VU)
        for (int i = 0; i < N; ++i) zn(w, j, i) = w.d(T_L + (j)) * zn(w, Lq, i) + zn(w, j, i);
      }
  }
  PLM_LHD void decrease_bdf(const Ws& w)
  {
    double hsum, xi;
    for (int i = 0; i <= qmax; i++) w.d(T_L + (i)) = 0.0;
    w.d(T_L + (2)) = 1.0;
    hsum = 0.0;
    for (int j = 1; j <= q - 2; j++) {
      hsum += w.d(T_TAU + (j));
      xi = hsum / hscale;
      for (int i = j + 2; i >= 2; i--) w.d(T_L + (i)) = w.d(T_L + (i)) * xi + w.d(T_L + (i - 1));
    }
    if (q > 2)
      for (int j = 2; j < q; j++) {
#pragma unroll(VU)
        for (int i = 0; i < N; ++i) zn(w, j, i) = -w.d(T_L + (j)) * zn(w, q, i) + zn(w, j, i);
      }
  }
  PLM_LHD void predict(const Ws& w)
  {
    tn += h;
    for (int k = 1; k <= q; k++)
      for (int j = q; j >= k; j--) {
#pragma unroll(VU)
        for (int i = 0; i < N; ++i) zn(w, j - 1, i) = zn(w, j - 1, i) + zn(w, j, i);
      }
  }
  PLM_LHD void restore(const Ws& w)
  {
    tn = saved_t;
    for (int k = 1; k <= q; k++)
      for (int j = q; j >= k; j--) {
#pragma unroll(VU)
        for (int i = 0; i < N; ++i) zn(w, j - 1, i) = zn(w, j - 1, i) - zn(w, j, i);
      }
  }
  PLM_LHD void set_bdf(const Ws& w)
  {
    double alpha0, alpha0_hat, xi_inv, xistar_inv, hsum;
    w.d(T_L + (0)) = w.d(T_L + (1)) = xi_inv = xistar_inv = 1.0;
    for (int i = 2; i <= q; i++) w.d(T_L + (i)) = 0.0;
    alpha0 = alpha0_hat = -1.0;
    hsum = h;
    if (q > 1) {
      for (int j = 2; j < q; j++) {
        hsum += w.d(T_TAU + (j - 1));
        xi_inv = h / hsum;
        alpha0 -= 1.0 / j;
        for (int i = j; i >= 1; i--) w.d(T_L + (i)) += w.d(T_L + (i - 1)) * xi_inv;
      }
      alpha0 -= 1.0 / q;
      xistar_inv = -w.d(T_L + (1)) - alpha0;
      hsum += w.d(T_TAU + (q - 1));
      xi_inv = h / hsum;
      alpha0_hat = -w.d(T_L + (1)) - xi_inv;
      for (int i = q; i >= 1; i--) w.d(T_L + (i)) += w.d(T_L + (i - 1)) * xistar_inv;
    }
    // cvSetTqBDF
    double A1, A2, A3, A4, A5, A6, C, Cpinv, Cppinv;
    A1 = 1.0 - alpha0_hat + alpha0;
    A2 = 1.0 + q * A1;
    w.d(T_TQ + (2)) = lane_fabs(A1 / (alpha0 * A2));
    w.d(T_TQ + (5)) = lane_fabs(A2 * xistar_inv / (w.d(T_L + (q)) * xi_inv));
    if (qwait == 1) {
      if (q > 1) {
        C = xistar_inv / w.d(T_L + (q));
        A3 = alpha0 + 1.0 / q;
        A4 = alpha0_hat + xi_inv;
        Cpinv = (1.0 - A4 + A3) / A3;
        w.d(T_TQ + (1)) = lane_fabs(C * Cpinv);
      } else
        w.d(T_TQ + (1)) = 1.0;
      hsum += w.d(T_TAU + (q));
      xi_inv = h / hsum;
      A5 = alpha0 - (1.0 / (q + 1));
      A6 = alpha0_hat - xi_inv;
      Cppinv = (1.0 - A6 + A5) / A2;
      w.d(T_TQ + (3)) = lane_fabs(Cppinv / (xi_inv * (q + 2) * A5));
    }
    w.d(T_TQ + (4)) = cv::CORTES / w.d(T_TQ + (2));
    rl1 = 1.0 / w.d(T_L + (1));
    gamma = h * rl1;
    if (nst == 0) gammap = gamma;
    gamrat = (nst > 0) ? gamma / gammap : 1.0;
  }

  // ---- (E) the shared tail of a round: whatever (B) and (D) left pending, in the order cvStep runs it ----
  PLM_LHD void tail(const Ws& w)
  {
    if (pend == 0) return;
    if (pend & P_RESTORE) restore(w);
    if (pend & P_BEGIN) {  // the step loop of CVode + the head of cvStep
      bool ok = true;
      if (nst > 0) ok = ewtset(w);
      if (!ok) fail(cv::CV_ILL_INPUT);
      else if (nstloc >= mxstep) fail(cv::CV_TOO_MUCH_WORK);
      else if (cv::UROUND * wrms(w, L::ZN) > 1.0) fail(cv::CV_TOO_MUCH_ACC);
      else {
        if ((nst > 0) && (hprime != h)) {  // cvAdjustParams
          if (qprime != q) {
            adj = qprime - q;
            pend |= P_ADJ;
          }
          pend |= P_RESCALE;
        }
        saved_t = tn;
        nflag = cv::FIRST_CALL;
        ncf = nef = 0;
        pend |= P_ATTEMPT;
      }
    }
    if (pend & P_FAIL) {
#pragma unroll(VU)
      for (int i = 0; i < N; ++i) w.d(L::Y + i) = zn(w, 0, i);
      phase = PH_OUT;
      pend = 0;
      return;
    }
    if (pend & P_ERRF) {  // order 1 and more than MXNEF1 error-test failures: reload zn1 = h f(tn, zn0)
#pragma unroll(VU)
      for (int i = 0; i < N; ++i) w.d(L::Y + i) = zn(w, 0, i);
      t_eval = tn;
      phase = PH_ERRF;
      pend = 0;
      return;
    }
    if (pend & P_ADJ) {  // cvAdjustOrder + the bookkeeping of its two callers (same result in both)
      if (!((q == 2) && (adj != 1))) {
        if (adj == 1) increase_bdf(w);
        else if (adj == -1) decrease_bdf(w);
      }
      q += adj;
      Lq = q + 1;
      qwait = Lq;
    }
    if (pend & P_RESCALE) rescale(w);
    if (pend & P_ATTEMPT) {  // top of cvStep's retry loop up to the first residual of the Newton iteration
      predict(w);
      set_bdf(w);
      convfail = ((nflag == cv::FIRST_CALL) || (nflag == cv::PREV_ERR_FAIL)) ? cv::CV_NO_FAILURES : cv::CV_FAIL_OTHER;
      call_setup = (nflag == cv::PREV_CONV_FAIL) || (nflag == cv::PREV_ERR_FAIL) || (nst == 0) ||
                   (nst >= nstlp + cv::MSBP) || (lane_fabs(gamrat - 1.0) > cv::DGMAX_LSETUP);
      jbad = 0;
      nls_jcur = 0;
      pend |= P_FIRSTRES;
    }
    if (pend & P_FIRSTRES) {  // acor = 0, y = zn0 + acor, residual wanted
#pragma unroll(VU)
      for (int i = 0; i < N; ++i) {
        w.d(L::ACOR + i) = 0.0;
        w.d(L::Y + i) = zn(w, 0, i) + 0.0;
      }
      first_res = 1;
      t_eval = tn;
      phase = PH_NEWTON;
    }
    pend = 0;
  }

  // ---- (B) consume a right-hand side ----
  PLM_LHD void post_rhs(const Ws& w, const double (&f)[N], double T)
  {
    bool trial = false;
    switch (phase) {
      case PH_LOAD:
        w.d(L::Y + NS) = T;
        init(w);
        break;
      case PH_FINISH:
        w.d(L::Y + NS) = T;
        phase = PH_DONE;
        break;
      case PH_RHS0:
        nfe++;
#pragma unroll
        for (int i = 0; i < N; ++i) zn(w, 1, i) = f[i];
        phase = PH_IDLE;  // set by whatever follows
        trial = hin_begin(w);
        break;
      case PH_HIN:
        nfe++;
        phase = PH_IDLE;
        trial = hin_post(w, f);
        break;
      case PH_ERRF:
        nfe++;
#pragma unroll
        for (int i = 0; i < N; ++i) zn(w, 1, i) = h * f[i];
        phase = PH_IDLE;
        pend |= P_ATTEMPT;
        break;
      case PH_NEWTON:
        nfe++;
#pragma unroll
        for (int i = 0; i < N; ++i) w.d(L::F + i) = f[i];
        if (first_res && call_setup) {
          if (jbad) convfail = cv::CV_FAIL_BAD_J;
          setup_decide();
          phase = PH_WAIT_SETUP;
        } else {
          if (first_res) mnewt = 0;
          phase = PH_SOLVE;
        }
        first_res = 0;
        break;
      default:
        break;
    }
    if (trial) hin_trial(w);
  }

  // ---- (C) cvNlsLSetup + cvLsSetup + cvLsLinSys + SUNLinSolSetup_Dense ----
  struct JAcc {
    Ws w;
    int off;
    PLM_LHD double& operator()(int r, int c) const { return w.g(off + c * N + r); }
  };
  // The setup in three pieces, so that the kernel can run them at different granularities: setup_decide when the lane starts
  // to wait (does it need a fresh Jacobian?), setup_jac for the lanes that do (lane-parallel straight-line code, batched),
  // then the factorisation -- getrf() below per lane, or the warp-cooperative one of lane_kernels.cuh per cell -- and
  // setup_finish.  setup() is all of it for one lane (host model; fallback).
  int need_jac = 0;
  PLM_LHD void setup_decide()
  {
    const double dgamma = lane_fabs((gamma / gammap) - 1.0);
    const bool jb = (nst == 0) || (nst >= nstlj + msbj) || ((convfail == cv::CV_FAIL_BAD_J) && (dgamma < cv::DGMAX_JBAD)) ||
                    (convfail == cv::CV_FAIL_OTHER);
    jcur = jb ? 1 : 0;
    need_jac = jcur;
  }
  PLM_LHD void setup_jac(const Ws& w)
  {
    JAcc ja{w, L::SAV};
    if constexpr (M::kTable) {
      lt_cjac<M>(LV{&w.d(L::ZN)}, ja, LV{&w.d(L::SC)}, LV{&w.d(L::G)}, LV{&w.g(L::JSCR)}, LV{&w.g(L::JSCR + NS)},
                 LV{&w.g(L::JSCR + 2 * NS)}, LV{&w.g(L::JSCR + 3 * NS)});
    } else {
      double y[N];
#pragma unroll
      for (int i = 0; i < N; ++i) y[i] = zn(w, 0, i);  // = y of the residual just evaluated (acor = 0)
      lane_cjac<M>(y, ja);
    }
    nje++;
    nstlj = nst;
    need_jac = 0;
  }
  PLM_LHD void setup_finish(int ier)
  {
    nsetups++;
    gamrat = 1.0;
    gammap = gamma;
    crate = 1.0;
    nstlp = nst;
    nls_jcur = jcur;
    if (ier > 0) {
      phase = PH_IDLE;
      nls_failed();  // "LSetup failure leaves the setup loop directly": recoverable failure of the nonlinear solve
      return;
    }
    mnewt = 0;
    phase = PH_SOLVE;
  }
  PLM_LHD void setup(const Ws& w)
  {
    setup_decide();
    if (need_jac) setup_jac(w);
    setup_finish(getrf(w));
  }
  // A = I - gamma J (cvLsLinSys) and denseGETRF (partial pivoting, first maximum) in ONE pass over the saved Jacobian,
  // column by column ("left-looking"): column j is formed in the FAST space, receives the updates of the finished columns
  // k < j in the order k = 0, 1, ... (the order, operands and roundings of denseGETRF's right-looking sweeps: the factors
  // are bit-identical), is searched for its pivot, scaled, and stored once.  Global traffic: every factor is written once
  // and read n/3 times on average; nothing is read back right after it was written.  Pivot rows kept as doubles in PERM.
  PLM_LHD int getrf(const Ws& w)
  {
    // the accumulated correction is zero whenever a setup runs (first residual of a Newton attempt): its vector is the
    // scratch column, and is zeroed again afterwards
    constexpr int COL = L::ACOR;
    const double c = -gamma;
    int ier = 0;
#ifdef LANE_MODEL_COUNT_SWAPS
    const long swaps0 = lane_model_swaps;
#endif
    for (int j = 0; j < N; ++j) {
      // column j of I - gamma J with the row interchanges of the columns before it
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double a = w.g(L::SAV + j * N + i) * c;
        if (i == j) a += 1.0;
        w.d(COL + i) = a;
      }
      for (int k = 0; k < j; ++k) {
        const int pk = int(w.g(L::PERM + k));
        if (pk != k) {
          const double t = w.d(COL + k);
          w.d(COL + k) = w.d(COL + pk);
          w.d(COL + pk) = t;
        }
      }
      for (int k = 0; k < j; ++k) {
        const double a_kj = w.d(COL + k);
        if (a_kj != 0.0) {
#pragma unroll
          for (int i = 0; i < N; ++i)
            if (i > k) w.d(COL + i) -= a_kj * w.g(L::MAT + k * N + i);
        }
      }
      // pivot of column j
      int lp = j;
      double amax = lane_fabs(w.d(COL + j));
      for (int i = j + 1; i < N; ++i) {
        const double v = lane_fabs(w.d(COL + i));
        if (v > amax) {
          amax = v;
          lp = i;
        }
      }
      w.g(L::PERM + j) = double(lp);
#ifdef LANE_MODEL_COUNT_SWAPS
      lane_model_swaps += (lp != j);
      lane_model_pivots += 1;
#endif
      if (w.d(COL + lp) == 0.0) {
        ier = j + 1;
        break;
      }
      if (lp != j) {  // swap rows j and lp: in this column, and in the multipliers already stored
        const double t = w.d(COL + lp);
        w.d(COL + lp) = w.d(COL + j);
        w.d(COL + j) = t;
        for (int k = 0; k < j; ++k) {
          const double u = w.g(L::MAT + k * N + lp);
          w.g(L::MAT + k * N + lp) = w.g(L::MAT + k * N + j);
          w.g(L::MAT + k * N + j) = u;
        }
      }
      const double mult = 1.0 / w.d(COL + j);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const double v = w.d(COL + i);
        w.g(L::MAT + j * N + i) = (i > j) ? v * mult : v;
      }
    }
#ifdef LANE_MODEL_COUNT_SWAPS
    lane_model_lus += 1;
    lane_model_lus_swapped += (lane_model_swaps != swaps0);
#endif
#pragma unroll
    for (int i = 0; i < N; ++i) w.d(COL + i) = 0.0;
    return ier;
  }

  // ---- (D) one Newton correction: cvLsSolve (denseGETRS), cvNlsConvTest, and whatever follows up to the next residual ----
  PLM_LHD void solve(const Ws& w)
  {
    nni++;
    phase = PH_IDLE;  // set by whatever follows
    // b = -(rl1 * zn1 + acor - gamma * f), pivot swaps applied in memory, then the triangular solves in registers
#pragma unroll(VU)
    for (int i = 0; i < N; ++i) {
      double res = rl1 * zn(w, 1, i) + w.d(L::ACOR + i);
      res = -gamma * w.d(L::F + i) + res;
      w.d(L::DEL + i) = -1.0 * res;
    }
    {
      int pv[N];
#pragma unroll
      for (int k = 0; k < N; ++k) pv[k] = int(w.g(L::PERM + k));
#pragma unroll
      for (int k = 0; k < N; ++k) {
        if (pv[k] != k) {
          const double t = w.d(L::DEL + k);
          w.d(L::DEL + k) = w.d(L::DEL + pv[k]);
          w.d(L::DEL + pv[k]) = t;
        }
      }
    }
    double b[N];
#pragma unroll
    for (int i = 0; i < N; ++i) b[i] = w.d(L::DEL + i);
#pragma unroll
    for (int k = 0; k < N - 1; ++k) {
#pragma unroll
      for (int i = k + 1; i < N; ++i) b[i] -= w.g(L::MAT + k * N + i) * b[k];
    }
#pragma unroll
    for (int k = N - 1; k > 0; --k) {
      b[k] = lane_div(b[k], w.g(L::MAT + k * N + k));
#pragma unroll
      for (int i = 0; i < k; ++i) b[i] -= w.g(L::MAT + k * N + i) * b[k];
    }
    b[0] = lane_div(b[0], w.g(L::MAT));
    if (gamrat != 1.0) {
      const double s = 2.0 / (1.0 + gamrat);
#pragma unroll
      for (int i = 0; i < N; ++i) b[i] *= s;
    }
    double sum = 0.0, suma = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double ew = w.d(L::EWT + i);
      const double ac = w.d(L::ACOR + i) + b[i];
      w.d(L::ACOR + i) = ac;
      const double p = b[i] * ew, pa = ac * ew;
      sum += p * p;
      suma += pa * pa;
    }
    const double del = sqrt(sum / N);
    if (mnewt > 0) {
      const double r = del / delp;
      crate = (cv::CRDOWN * crate > r) ? cv::CRDOWN * crate : r;
    }
    const double dcon = del * ((1.0 < crate) ? 1.0 : crate) / w.d(T_TQ + (4));
    if (dcon <= 1.0) {
      acnrm = (mnewt == 0) ? del : sqrt(suma / N);
      nls_success(w);
      return;
    }
    bool failed = (mnewt >= 1) && (del > cv::RDIV * delp);
    if (!failed) {
      delp = del;
      mnewt++;
      failed = (mnewt >= maxcor);
    }
    if (!failed) {  // next iteration: y = zn0 + acor
#pragma unroll(VU)
      for (int i = 0; i < N; ++i) w.d(L::Y + i) = zn(w, 0, i) + w.d(L::ACOR + i);
      t_eval = tn;
      phase = PH_NEWTON;
      return;
    }
    if (!nls_jcur) {  // stale Jacobian: redo the iteration with a fresh one
      call_setup = 1;
      jbad = 1;
      pend |= P_FIRSTRES;
      return;
    }
    nls_failed();
  }
  // cvHandleNFlag for a recoverable failure of the nonlinear solver
  PLM_LHD void nls_failed()
  {
    ncfn++;
    pend |= P_RESTORE;
    ncf++;
    etamax = 1.0;
    if (ncf == cv::MXNCF) { fail(cv::CV_CONV_FAILURE); return; }
    eta = cv::ETA_CF;
    nflag = cv::PREV_CONV_FAIL;
    pend |= P_RESCALE | P_ATTEMPT;
  }
  PLM_LHD static double rpowR(double base, int k) { return base <= 0.0 ? 0.0 : lane_rroot(base, k); }
  // cvDoErrorTest onwards
  PLM_LHD void nls_success(const Ws& w)
  {
    jcur = 0;
    dsm = acnrm * w.d(T_TQ + (2));
    if (dsm > 1.0) {
      nef++;
      netf++;
      nflag = cv::PREV_ERR_FAIL;
      pend |= P_RESTORE;
      if (nef == maxnef) { fail(cv::CV_ERR_FAILURE); return; }
      etamax = 1.0;
      if (nef <= cv::MXNEF1) {
        eta = 1.0 / (rpowR(cv::BIAS2 * dsm, Lq) + cv::ADDON);
        eta = (cv::ETA_MIN_EF > eta) ? cv::ETA_MIN_EF : eta;
        if (nef >= cv::SMALL_NEF) eta = (eta < cv::ETA_MAX_EF) ? eta : cv::ETA_MAX_EF;
        pend |= P_RESCALE | P_ATTEMPT;
        return;
      }
      if (q > 1) {
        eta = cv::ETA_MIN_EF;
        adj = -1;
        pend |= P_ADJ | P_RESCALE | P_ATTEMPT;
        return;
      }
      eta = cv::ETA_MIN_EF;
      h *= eta;
      hscale = h;
      qwait = cv::LONG_WAIT;
      pend |= P_ERRF;
      return;
    }
    complete_step(w);
    prepare_next_step(w);
    etamax = (nst <= cv::SMALL_NST) ? cv::ETA_MAX_ES : cv::ETA_MAX_GS;
#pragma unroll(VU)
    for (int i = 0; i < N; ++i) w.d(L::ACOR + i) *= w.d(T_TQ + (2));
    nstloc++;
    if ((tn - tout) * h >= 0.0) {
      // CVodeGetDky(tout, 0): N_VLinearCombination over zn[q..0] with s^j
      const double s = (tout - tn) / h;
      double cvals[QM + 1];
      for (int j = q; j >= 0; j--) {
        cvals[j] = 1.0;
        for (int i = 0; i < j; i++) cvals[j] *= s;
      }
#pragma unroll(VU)
      for (int i = 0; i < N; ++i) {
        double z = cvals[q] * zn(w, q, i);
        for (int j = q - 1; j >= 0; j--) z += cvals[j] * zn(w, j, i);
        w.d(L::Y + i) = z;
      }
      flag = 0;
      phase = PH_OUT;
      return;
    }
    pend |= P_BEGIN;
  }
  PLM_LHD void complete_step(const Ws& w)
  {
    nst++;
    for (int i = q; i >= 2; i--) w.d(T_TAU + (i)) = w.d(T_TAU + (i - 1));
    if ((q == 1) && (nst > 1)) w.d(T_TAU + (2)) = w.d(T_TAU + (1));
    w.d(T_TAU + (1)) = h;
    for (int j = 0; j <= q; j++) {
#pragma unroll(VU)
      for (int i = 0; i < N; ++i) zn(w, j, i) = w.d(T_L + (j)) * w.d(L::ACOR + i) + zn(w, j, i);
    }
    qwait--;
    if ((qwait == 1) && (q != qmax)) {
#pragma unroll(VU)
      for (int i = 0; i < N; ++i) zn(w, qmax, i) = w.d(L::ACOR + i);
      saved_tq5 = w.d(T_TQ + (5));
    }
  }
  PLM_LHD void set_eta()
  {
    if ((eta > cv::ETA_MIN_FX) && (eta < cv::ETA_MAX_FX)) {
      eta = 1.0;
      hprime = h;
    } else {
      if (eta >= cv::ETA_MAX_FX) eta = (eta < etamax) ? eta : etamax;
      else eta = (eta > cv::ETA_MIN) ? eta : cv::ETA_MIN;
      hprime = h * eta;
    }
  }
  PLM_LHD void prepare_next_step(const Ws& w)
  {
    if (etamax == 1.0) {
      qwait = (qwait > 2) ? qwait : 2;
      qprime = q;
      hprime = h;
      eta = 1.0;
      return;
    }
    const double etaq = 1.0 / (rpowR(cv::BIAS2 * dsm, Lq) + cv::ADDON);
    if (qwait != 0) {
      eta = etaq;
      qprime = q;
      set_eta();
      return;
    }
    qwait = 2;
    double etaqm1 = 0.0;
    if (q > 1) {
      const double ddn = wrms(w, L::ZN + q * N) * w.d(T_TQ + (1));
      etaqm1 = 1.0 / (rpowR(cv::BIAS1 * ddn, q) + cv::ADDON);
    }
    double etaqp1 = 0.0;
    if (q != qmax && saved_tq5 != 0.0) {
      double rp = 1.0;  // SUNRpowerI(h / w.d(T_TAU + (2)), L)
      const double base = h / w.d(T_TAU + (2));
      for (int i = 1; i <= Lq; i++) rp *= base;
      const double cquot = (w.d(T_TQ + (5)) / saved_tq5) * rp;
      double sum = 0.0;
#pragma unroll(VU)
      for (int i = 0; i < N; ++i) {
        const double tv = -cquot * zn(w, qmax, i) + w.d(L::ACOR + i);
        const double p = tv * w.d(L::EWT + i);
        sum += p * p;
      }
      const double dup = sqrt(sum / N) * w.d(T_TQ + (3));
      etaqp1 = 1.0 / (rpowR(cv::BIAS3 * dup, Lq + 1) + cv::ADDON);
    }
    double etam = (etaq > etaqp1) ? etaq : etaqp1;
    etam = (etaqm1 > etam) ? etaqm1 : etam;
    if ((etam > cv::ETA_MIN_FX) && (etam < cv::ETA_MAX_FX)) {
      eta = 1.0;
      qprime = q;
    } else if (etam == etaq) {
      eta = etaq;
      qprime = q;
    } else if (etam == etaqm1) {
      eta = etaqm1;
      qprime = q - 1;
    } else {
      eta = etaqp1;
      qprime = q + 1;
#pragma unroll(VU)
      for (int i = 0; i < N; ++i) zn(w, qmax, i) = w.d(L::ACOR + i);
    }
    set_eta();
  }
};

}  // namespace lane
}  // namespace plm
