// sm_100a kernel templates of the lane = cell path (one THREAD per cell; lane_integ.h explains the round structure).
//
//   lane_react_kernel<M, I, LAY, EINT>   persistent warps; every LANE pulls its own cells from the global queue, so a warp
//                                         always carries 32 cells that are each somewhere in their own integration
//   lane_rhs_kernel / lane_jac_kernel     one fKernelSpec / cJac per thread (parity entry points and the A/B timing of the
//                                         straight-line mechanism code against the warp-per-cell kernels)
//
// Used for the mechanisms whose per-cell state is small enough to keep 32 cells per warp in flight (LiDryer, drm19) with
// ReactorRK64 and with ReactorCvode + analytic Jacobian on the enthalpy reactor -- the configuration PeleLM runs; every
// other flavour (dodecane_lu, difference-quotient Jacobian, internal-energy BDF) goes through the warp-per-cell kernels
// (kernels.cu).  launch_lane_react returns cudaErrorNotSupported when the request is not covered.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "kernels.h"
#include "lane_cell.h"

namespace plm {

using namespace lane;

namespace lanek {

// CTA shape: one CTA per SM with as many warps (<= 8) as the shared-memory vectors of 32 cells per warp allow
template <class I, class LAY>
struct LaneShape {
  static constexpr int FAST = I::kBdf ? LAY::FAST_BDF : LAY::FAST_RK;
  static constexpr int BIG = I::kBdf ? LAY::BIG_BDF : LAY::BIG_RK;
#ifdef PLM_LANE_COOP_BUILD   // the warp-cooperative LU (measured and rejected, see coop_getrf) needs one matrix of scratch per warp
  static constexpr int LUA = I::kBdf ? LAY::N * (LAY::N + 1) : 0;
#else
  static constexpr int LUA = 0;
#endif
  static constexpr int WARP_BYTES = (FAST * 32 + LUA) * 8;
  static constexpr int FIT = (227 * 1024 - 1024) / WARP_BYTES;
  // 8 warps of 255 registers fill the register file; the explicit integrator of the smallest mechanism needs 144 registers
  // and is bound by FP64 dependency latency (`wait` stalls), so it runs 12
  static constexpr int MAXW = (!I::kBdf && LAY::N <= 12) ? 12 : 8;
  static constexpr int WARPS = FIT >= MAXW ? MAXW : FIT;
  static constexpr int THREADS = WARPS * 32;
  static constexpr int SMEM = WARPS * WARP_BYTES + kMaxN * 8;
  static_assert(WARPS >= 1, "the per-cell vectors of one warp do not fit in shared memory");
};
constexpr int kLaneThreads = 128;   // rhs / jac parity kernels

__device__ __forceinline__ int64_t lane_a4_off(const plm_a4d& a, int i, int j, int k)
{
  return int64_t(i - a.lo[0]) + int64_t(j - a.lo[1]) * a.jstride + int64_t(k - a.lo[2]) * a.kstride;
}

// the cell's addresses; returns the mask value (1 when there is no mask)
template <int NS>
__device__ __forceinline__ int lane_view(const BoxArgs& a0, int64_t c, CellView& v)
{
  const BoxArgs* bx = &a0;
  if (a0.nboxes > 0) {  // the box that holds cell c: off[b] <= c < off[b + 1]
    int lo = 0, hi = a0.nboxes;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(a0.off + mid) <= c) lo = mid;
      else hi = mid;
    }
    bx = a0.tab + lo;
    c -= __ldg(a0.off + lo);
  }
  const BoxArgs& a = *bx;
  if (a0.flat) {
    v.rY = a.rY.p + c * (NS + 1); v.sY = 1;
    v.rYsrc = a.rYsrc.p + c * NS; v.sS = 1;
    v.T = nullptr; v.rE = a.rE.p + c; v.rEsrc = a.rEsrc.p + c; v.FC = a.FC.p + c; v.IR = nullptr; v.sI = 0;
    return 1;
  }
  const int nx = a.hi[0] - a.lo[0] + 1, ny = a.hi[1] - a.lo[1] + 1;
  const int i = a.lo[0] + int(c % nx), j = a.lo[1] + int((c / nx) % ny), k = a.lo[2] + int(c / (int64_t(nx) * ny));
  v.rY = a.rY.p + lane_a4_off(a.rY, i, j, k); v.sY = a.rY.nstride;
  v.rYsrc = a.rYsrc.p + lane_a4_off(a.rYsrc, i, j, k); v.sS = a.rYsrc.nstride;
  v.T = a.T.p + lane_a4_off(a.T, i, j, k);
  v.rE = a.rE.p + lane_a4_off(a.rE, i, j, k);
  v.rEsrc = a.rEsrc.p + lane_a4_off(a.rEsrc, i, j, k);
  v.FC = a.FC.p + lane_a4_off(a.FC, i, j, k);
  v.IR = a.IR.p ? a.IR.p + lane_a4_off(a.IR, i, j, k) : nullptr; v.sI = a.IR.nstride;
  if (a.mask.p)
    return a.mask.p[int64_t(i - a.mask.lo[0]) + int64_t(j - a.mask.lo[1]) * a.mask.jstride +
                    int64_t(k - a.mask.lo[2]) * a.mask.kstride];
  return 1;
}

// Warp-cooperative factorisation of ONE cell's Newton matrix: A = I - gamma J (cvLsLinSys) and denseGETRF (partial pivoting,
// first maximum: sundials_dense.c:135-196) with lane = row, the matrix in shared memory (leading dimension odd: rows are
// swept conflict-free).  Every element receives the updates A[i][j] -= A[k][j] * L[i][k] in the order k = 0, 1, ... with the
// operands of the serial routine, so the factors are bit-identical to LaneBdf::getrf.  `big` = the BIG space of the cell
// (lane offset included): saved Jacobian in, LU factors and pivot rows out.  Returns denseGETRF's flag (0 or k + 1).
// A lane-parallel LU costs the same instructions whether 2 or 32 lanes of the warp need one; a linear-solver setup is
// wanted by ~6 % of the lanes per round, so the warp serves them one cell at a time, 22 lanes wide, as soon as they ask.
template <class LAY>
__device__ __forceinline__ int coop_getrf(double* __restrict__ A, double* __restrict__ big, double gamma, int lane)
{
  constexpr int N = LAY::N, LD = (N % 2 == 0) ? N + 1 : N;
  const double c = -gamma;
  for (int e = lane; e < N * N; e += 32) {
    const int col = e / N, row = e - col * N;
    double a = big[size_t(LAY::SAV + e) * 32] * c;
    if (row == col) a += 1.0;
    A[row * LD + col] = a;
  }
  __syncwarp();
  int ier = 0;
#pragma unroll 1
  for (int k = 0; k < N; ++k) {
    // pivot: first row >= k with the largest |a_ik| (non-negative doubles order like their bit patterns)
    const bool mine = lane >= k && lane < N;
    const double av = mine ? fabs(A[lane * LD + k]) : 0.0;
    const unsigned hi = mine ? unsigned(__double2hiint(av)) : 0u, lo = unsigned(__double2loint(av));
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, (mine && hi == mhi) ? lo : 0u);
    const unsigned cand = __ballot_sync(0xffffffffu, mine && hi == mhi && lo == mlo);
    const int p = cand ? __ffs(cand) - 1 : k;
    if (lane == 0) big[size_t(LAY::PERM + k) * 32] = double(p);
    if (A[p * LD + k] == 0.0) {
      ier = k + 1;
      break;
    }
    if (p != k && lane < N) {  // lane = column
      const double t = A[k * LD + lane];
      A[k * LD + lane] = A[p * LD + lane];
      A[p * LD + lane] = t;
    }
    __syncwarp();
    const double akk = A[k * LD + k];
    const double mult = (fabs(akk) > 1e-280 && fabs(akk) < 1e280) ? frcp(akk) : 1.0 / akk;  // ONE / col_k[k]
    if (lane > k && lane < N) {
      double* row = A + lane * LD;
      const double lik = row[k] * mult;
      row[k] = lik;
      const double* rk = A + k * LD;
#pragma unroll 2
      for (int j = k + 1; j < N; ++j) {
        const double a_kj = rk[j];
        if (a_kj != 0.0) row[j] -= a_kj * lik;
      }
    }
    __syncwarp();
  }
  for (int e = lane; e < N * N; e += 32) {
    const int col = e / N, row = e - col * N;
    big[size_t(LAY::MAT + e) * 32] = A[row * LD + col];
  }
  __syncwarp();
  return ier;
}

template <class M, class I, class LAY, int EINT>
__global__ void __launch_bounds__(LaneShape<I, LAY>::THREADS, 1)
    lane_react_kernel(const ReactOpts ro, const BoxArgs a, unsigned long long* __restrict__ counter,
                      unsigned long long* __restrict__ gstats, int* __restrict__ status, double* __restrict__ ws, int theta)
{
  constexpr int NS = M::NS;
  constexpr bool BDF = I::kBdf;
  using SH = LaneShape<I, LAY>;
  extern __shared__ double lane_smem[];
  double* s_typ = lane_smem + SH::WARPS * (SH::FAST * 32 + SH::LUA);
  double* lu_scratch = lane_smem + SH::WARPS * SH::FAST * 32 + size_t(threadIdx.x >> 5) * SH::LUA;
  for (int i = threadIdx.x; i < kMaxN; i += blockDim.x) s_typ[i] = ro.typ_vals[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const Ws w{lane_smem + size_t(threadIdx.x >> 5) * SH::FAST * 32 + lane, ws + size_t(gwarp) * SH::BIG * 32 + lane};
  LaneOpts o;
  o.rtol = ro.rtol; o.atol = ro.atol; o.qmax = ro.qmax; o.maxcor = ro.maxcor; o.use_typ = ro.use_typ;
  o.mxstep = 10000; o.maxnef = 100; o.msbj = 100;
  o.rk64_guess = ro.rk64_guess; o.rk64_min = ro.rk64_min; o.rk64_max = ro.rk64_max;
  o.typ_vals = s_typ;
  CellCfg g;
  g.time = a.time; g.dt = a.dt; g.dt_inv = a.dt_inv; g.flat = a.flat; g.mks = a.mks;
  g.clean = (BDF && ro.clean_init_massfrac && !a.flat) ? 1 : 0;

  I it;
  it.phase = PH_IDLE;
  int64_t mycell = -1;
  bool exhausted = false;
  int wait_acc = 0;
  const int64_t ncells = a.ncells;
  // theta bit 30: phase-aligned rounds -- the warps of the CTA start every round together (one barrier per round), so that
  // they stream through the straight-line code as ONE instruction stream the instruction cache can prefetch
  const bool align = (theta >> 30) & 1;
  // theta bit 29 (development, PLM_LANE_COOP=1): factorisations warp-cooperative, one cell at a time, as soon as a lane asks,
  // instead of lane-parallel when enough lanes wait.  Measured on B200 (profiles/r02f_*): 19 % fewer instructions, but the
  // warps of a CTA then carry different numbers of factorisations per round and wait for each other: drm19 -12 %, LiDryer -25 %.
#ifdef PLM_LANE_COOP_BUILD
  const bool coop = (theta >> 29) & 1;
#else
  const bool coop = false;
#endif
  theta &= 0x1fffffff;

#pragma unroll 1
  for (;;) {
    // ---- every idle lane takes the next cell of the queue (one atomic per warp and round) ----
    const bool want = (it.phase == PH_IDLE) && !exhausted;
    const unsigned wm = __ballot_sync(0xffffffffu, want);
    if (wm) {
      const int leader = __ffs(wm) - 1;
      unsigned long long base = 0;
      if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(wm));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (want) {
        const int64_t c = int64_t(base) + __popc(wm & ((1u << lane) - 1u));
        if (c >= ncells) {
          exhausted = true;
        } else {
          CellView v;
          const int mk = lane_view<NS>(a, c, v);
          if (BDF && mk == -1) {  // ReactorCvode.cpp:1395, :1439-1441 (ReactorRK64 ignores the mask)
            *v.FC = 0.0;
            if (status) status[c] = 0;
            atomicAdd(gstats + 8, 1ull);
            if (a.mks) {  // the reference still scales a skipped cell to CGS and back (PeleLM.cpp:5737-5772): same roundings
              for (int k = 0; k < NS; ++k) {
                double* dst = v.rY + k * v.sY;
                const double vold = *dst, vnew = __dmul_rn(__dmul_rn(vold, 1.0e-3), 1.0e3);
                if (v.IR) v.IR[k * v.sI] = __dsub_rn(__dmul_rn(-__dsub_rn(vold, vnew), a.dt_inv), v.rYsrc[k * v.sS]);
                *dst = vnew;
              }
              *v.rE = __dmul_rn(__dmul_rn(*v.rE, 10.0), 0.1);
            }
          } else {
            mycell = c;
            lane_cell_load<M, I, LAY>(it, w, v, g, o);
          }
        }
      }
    }
    const unsigned active = __ballot_sync(0xffffffffu, it.phase != PH_IDLE);
    if (align) {  // every warp of the CTA runs every round (and every barrier of it) until all of them are out of cells
      const bool warp_done = (active == 0) && (__ballot_sync(0xffffffffu, !exhausted) == 0);
      if (__syncthreads_and(warp_done)) break;
    } else if (active == 0) {
      if (__ballot_sync(0xffffffffu, !exhausted) == 0) break;
      continue;
    }
    // ---- (A) + (B) ----
    if (active) lane_round_rhs<M, EINT, I, LAY>(it, w);

    if constexpr (BDF) {
      // ---- (C) linear-solver setups.  Fresh Jacobians (lane-parallel straight-line code) when enough lanes want one, or
      //      nothing else can run; factorisations (warp-cooperative, one cell at a time) as soon as a lane asks ----
      bool go;
      const bool wants = it.phase == PH_WAIT_SETUP && (!coop || it.need_jac);
      if (align) {  // decided for the whole CTA: its warps then run the setup code together, as they run everything else
        const int nwait = __syncthreads_count(wants);
        const int nbusy = __syncthreads_count(it.phase != PH_IDLE && !wants);
        wait_acc += nwait;
        go = nwait > 0 && (wait_acc >= theta * int(blockDim.x >> 5) || nbusy == 0);
      } else {
        const unsigned waitm = __ballot_sync(0xffffffffu, wants);
        wait_acc += __popc(waitm);
        const unsigned busy = __ballot_sync(0xffffffffu, it.phase != PH_IDLE && !wants);
        go = waitm != 0 && (wait_acc >= theta || busy == 0);
      }
      if (go) {
        wait_acc = 0;
        if (wants) {
          if (it.need_jac) it.setup_jac(w);
          if (!coop) it.setup_finish(it.getrf(w));
        }
      }
#ifdef PLM_LANE_COOP_BUILD
      if (coop) {
        unsigned lum = __ballot_sync(0xffffffffu, it.phase == PH_WAIT_SETUP && !it.need_jac);
        int my_ier = 0;
        while (lum) {
          const int l = __ffs(lum) - 1;
          lum &= lum - 1;
          const double gam = __shfl_sync(0xffffffffu, it.gamma, l);
          const int ier = coop_getrf<LAY>(lu_scratch, w.big - lane + l, gam, lane);
          if (lane == l) my_ier = ier;
        }
        if (it.phase == PH_WAIT_SETUP && !it.need_jac) it.setup_finish(my_ier);
      }
#endif
      // ---- (D) Newton correction and the step logic behind it, (E) the shared tail ----
      if (it.phase == PH_SOLVE) it.solve(w);
      it.tail(w);
    }

    // ---- cells that finished: final species, then (next round) the temperature; or the store ----
    if (it.phase == PH_OUT || it.phase == PH_DONE) {
      CellView v;
      lane_view<NS>(a, mycell, v);
      if (it.phase == PH_OUT) lane_cell_out<M, I, LAY>(it, w, v, g);
      if (it.phase == PH_DONE) {
        LaneStats st;
        it.stats(st);
        const int fc = it.fc();
        lane_cell_store<M, I, LAY>(it, w, v, g);
        if (status) status[mycell] = st.flag;
        atomicAdd(gstats + 0, (unsigned long long)st.nst);
        atomicAdd(gstats + 1, (unsigned long long)st.nfe);
        if (BDF) {
          if (st.nje) atomicAdd(gstats + 2, (unsigned long long)st.nje);
          if (st.nsetups) atomicAdd(gstats + 3, (unsigned long long)st.nsetups);
          if (st.nni) atomicAdd(gstats + 4, (unsigned long long)st.nni);
          if (st.ncfn) atomicAdd(gstats + 5, (unsigned long long)st.ncfn);
          if (st.netf) atomicAdd(gstats + 6, (unsigned long long)st.netf);
          if (st.flag != 0) atomicAdd(gstats + 7, 1ull);
        }
        atomicAdd(gstats + 9, 1ull);
        atomicMax(gstats + 10, (unsigned long long)fc);
        atomicMax(gstats + 11, 0xffffffffull - (unsigned long long)fc);
      }
    }
  }
}

// one fKernelSpec per thread: y [c][n], ydot [c][n] (the layout of plm_eval_rhs)
struct PtrSrc {
  const double* s;
  double e;
  __device__ __forceinline__ double operator()(int k) const { return s[k]; }
  __device__ __forceinline__ double energy() const { return e; }
};
template <class M, int EINT>
__global__ void __launch_bounds__(kLaneThreads, 2)
    lane_rhs_kernel(int ncells, double t, const double* __restrict__ y, const double* __restrict__ rhoe,
                    const double* __restrict__ rhoesrc, const double* __restrict__ rYsrc, double* __restrict__ ydot)
{
  constexpr int N = M::N, NS = M::NS;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += gridDim.x * blockDim.x) {
    double yy[N], f[N];
#pragma unroll
    for (int i = 0; i < N; ++i) yy[i] = y[size_t(c) * N + i];
    PtrSrc src{rYsrc + size_t(c) * NS, rhoesrc[c]};
    if constexpr (M::kTable) {  // per-thread local arrays (this kernel is a parity entry point, not the integration kernel)
      double sc[NS], g[NS];
      lt_rhs<M, EINT>(LVt<1>{yy}, LVt<1>{f}, LVt<1>{sc}, LVt<1>{g}, rhoe[c] + rhoesrc[c] * t, src, false);
    } else {
      lane_rhs<M, EINT>(yy, f, rhoe[c] + rhoesrc[c] * t, src, false);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) ydot[size_t(c) * N + i] = f[i];
  }
}

template <int N>
struct JOut {
  double* p;
  __device__ __forceinline__ double& operator()(int r, int c) const { return p[c * N + r]; }
};
template <class M>
__global__ void __launch_bounds__(kLaneThreads, 2) lane_jac_kernel(int ncells, const double* __restrict__ y, double* __restrict__ J)
{
  constexpr int N = M::N;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += gridDim.x * blockDim.x) {
    double yy[N];
#pragma unroll
    for (int i = 0; i < N; ++i) yy[i] = y[size_t(c) * N + i];
    JOut<N> ja{J + size_t(c) * N * N};
    if constexpr (M::kTable) {
      double sc[M::NS], g[M::NS], h[M::NS], wd[M::NS], jT[M::NS], cR[M::NS];
      lt_cjac<M>(LVt<1>{yy}, ja, LVt<1>{sc}, LVt<1>{g}, LVt<1>{h}, LVt<1>{wd}, LVt<1>{jT}, LVt<1>{cR});
    } else {
      lane_cjac<M>(yy, ja);
    }
  }
}

}  // namespace lanek
}  // namespace plm
