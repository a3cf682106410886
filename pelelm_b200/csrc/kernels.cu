// sm_100a kernels behind the C ABI (include/plm_chem.h).
//
//   react_kernel<Lay, INTEG>    persistent warps pull cells from a global queue (per-cell cost varies
//                               10x inside a box, SURVEY.md section 7 "Divergence"): gather the cell from
//                               the FAB (box_flatten, ReactorUtils.H:76-136), invert h -> T, integrate
//                               (BDF: bdf_warp.cuh, RK64: below), scatter back (box_unflatten :138-207)
//   rhs_kernel / jac_kernel     one warp per cell, for the fKernelSpec / cJac parity tests
//   box_cost_kernel             per-box sum of FC (PeleLM::getFuncCountDM PeleLM.cpp:5557-5571)
//   dfma_peak_kernel            measured FP64 roofline denominator
// Kernels are instantiated per mechanism (compile-time species / reaction counts fix the shared-memory layout and
// let small loops unroll) and per maximum BDF order class (<= 2, <= 5).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "bdf_warp.cuh"
#include "kernels.h"

namespace plm {

static_assert(sizeof(DevMech) <= kHdr * sizeof(double) - 64, "DevMech + the per-warp slot words must fit in the shared-memory header");

// ---- ReactorRK64 (ReactorRK64.cpp:206-317; coefficients ReactorRK64.H:11-42), one warp per cell ----
__constant__ double c_alpha_rk64[6] = {0.218150805229859, 0.256702469801519, 0.527402592007520,
                                       0.0484864267224467, 1.24517071533530, 0.412366034843237};
__constant__ double c_beta_rk64[6] = {-0.113554138044166, -0.215118587818400, -0.0510152146250577,
                                      -1.07992686223881, -0.248664241213447, 0.0};
__constant__ double c_err_rk64[6] = {-0.0554699315064507, 0.158481845574980, -0.0905918835751907,
                                     -0.219084567203338, 0.164022338959433, 0.0426421977505659};

template <class L>
__device__ __forceinline__ int rk64_warp(double* wb, int lane, const ReactOpts& o, double (&soln)[L::NSLOT],
                                         double time_init, double dt_react, int& nfe, bool clamp_inside)
{
  constexpr int NSLOT = L::NSLOT, n = L::N;
  const double exp1 = 0.25, exp2 = 0.2, betaerr = 1.0, tinyval = 1e-50;
  double carry[NSLOT], err[NSLOT], rhs[NSLOT];
  const double time_out = time_init + dt_react;
  double current_time = time_init;
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) carry[s] = soln[s];
  double dt_rk = dt_react / double(o.rk64_guess);
  const double dt_rk_min = dt_react / double(o.rk64_max);
  const double dt_rk_max = dt_react / double(o.rk64_min);
  int nsteps = 0;
#pragma unroll 1
  while (current_time < time_out) {
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) err[s] = 0.0;
#pragma unroll 1
    for (int stage = 0; stage < 6; stage++) {
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) wb[L::YB + lane + 32 * s] = soln[s];
      __syncwarp();
      rhs_warp<L>(current_time - time_init);
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) rhs[s] = wb[L::FB + lane + 32 * s];
      nfe++;
      const double ce = c_err_rk64[stage], ca = c_alpha_rk64[stage], cb = c_beta_rk64[stage];
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        err[s] += ce * dt_rk * rhs[s];
        soln[s] = carry[s] + ca * dt_rk * rhs[s];
        carry[s] = soln[s] + cb * dt_rk * rhs[s];
      }
    }
    current_time += dt_rk;
    nsteps++;
    double me = tinyval;
#pragma unroll
    for (int s = 0; s < NSLOT; ++s)
      if (lane + 32 * s < n) me = fabs(err[s]) > me ? fabs(err[s]) : me;
    // NaN-propagating max is not needed: the reference's comparison chain also drops NaNs
    const double max_err = warp_max(me);
    if (max_err < o.atol) {
      const double cf = betaerr * dexp(exp1 * dlog(o.atol / max_err));
      dt_rk = fmin(dt_rk_max, dt_rk * cf);
    } else {
      const double cf = betaerr * dexp(exp2 * dlog(o.atol / max_err));
      dt_rk = fmax(dt_rk_min, dt_rk * cf);
    }
    // react(box) limits every new step to the time left (ReactorRK64.cpp:307); the 1dArray react applies that limit only
    // after its loop (:141-142), so its last step runs past time_out: reproduced
    if (clamp_inside) dt_rk = fmin(dt_rk, time_out - current_time);
  }
  return nsteps;
}

// copy the mechanism descriptor into the shared-memory header and set the constant pad entries of the warp's area
template <class L>
__device__ __forceinline__ double* block_prologue(const DevMech* __restrict__ dm)
{
  extern __shared__ double smem_dyn[];
  {
    const int words = int(sizeof(DevMech) / sizeof(int));
    const int* src = reinterpret_cast<const int*>(dm);
    int* dst = reinterpret_cast<int*>(smem_dyn);
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
  }
  const int lane = threadIdx.x & 31;
  if (lane == 0) *slot_word() = 0;
#ifdef PLM_SOFTALIGN
  if (threadIdx.x == 0) *softbar_word() = unsigned(blockDim.x >> 5);  // every warp a member, nobody arrived, generation 0
#endif
  if constexpr (L::TSM) {  // stage the tables of the right-hand side behind the warp areas (once per CTA)
    __syncthreads();
    const DevMech& m = smem_mech();
    double* tb = const_cast<double*>(Tab<L>::base());
    auto stage = [&](int off, const void* src, int ndoubles) {
      const double* sp = static_cast<const double*>(src);
      for (int i = threadIdx.x; i < ndoubles; i += blockDim.x) tb[off + i] = sp[i];
    };
    using T = Tab<L>;
    stage(T::IMW, m.imw, L::NP); stage(T::MW, m.mw, L::NP); stage(T::TMID, m.tmid, L::NP);
    stage(T::H, m.h, 12 * L::NP); stage(T::CP, m.cp, 10 * L::NP); stage(T::G, m.g, 14 * L::NP);
    stage(T::RP0, m.rp0, 2 * L::NRP); stage(T::RP1, m.rp1, 2 * L::NRP);
    stage(T::FO0, m.fo0, 64); stage(T::FO1, m.fo1, 64); stage(T::FO2, m.fo2, 64); stage(T::FO3, m.fo3, 64);
    stage(T::FO4, m.fo4, 32); stage(T::SEG, m.seg, (kMaxSeg * L::NP) / 2);
    stage(T::EFF, m.eff, m.emax * 64); stage(T::EFF + m.emax * 64, m.gat, m.glen * L::NP);
  }
  if constexpr (L::POOL > 0) {
    if (threadIdx.x < 32) reinterpret_cast<int*>(pool_base<L>() + L::POOL * L::MATD)[threadIdx.x] = 0;  // the locks: all free
  }
  for (int sl = 0; sl < L::SLOTS; ++sl) {
    double* w = warp_slot<L>(sl);
    for (int e = lane; e < L::TOTAL; e += 32) w[e] = 0.0;
    __syncwarp();
    if (lane == 0) w[L::CG + 2 * L::NS] = 1.0;  // unit concentration for unused reactant/product slots; g = h = 0
  }
  __syncthreads();
  return warp_slot<L>(0);
}

__device__ __forceinline__ int64_t a4_off(const plm_a4d& a, int i, int j, int k)
{
  return int64_t(i - a.lo[0]) + int64_t(j - a.lo[1]) * a.jstride + int64_t(k - a.lo[2]) * a.kstride;
}

// where cell c of the call lives in the caller's arrays (FAB layout, or the flat AoS layout of the 1dArray API)
struct CellAddr {
  int64_t oY, oS, oT, oE, oEs, oF, oI, sY, sS;
  int mk;
  int box;  // index into the box table of a multi-box call (plm_react_boxes), -1: the call's own box
};
// the box of a multi-box call that holds cell c (off[b] <= c < off[b + 1]); c becomes the cell's index inside it
__device__ __forceinline__ int find_box(const BoxArgs& a0, int64_t& c)
{
  if (a0.nboxes <= 0) return -1;
  int lo = 0, hi = a0.nboxes;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(a0.off + mid) <= c) lo = mid;
    else hi = mid;
  }
  c -= __ldg(a0.off + lo);
  return lo;
}
template <class L>
__device__ __noinline__ CellAddr cell_addr(const BoxArgs& a0, int64_t c)
{
  constexpr int ns = L::NS, n = L::N;
  CellAddr r;
  r.box = find_box(a0, c);
  const BoxArgs& a = (r.box >= 0) ? a0.tab[r.box] : a0;
  r.mk = 1;
  if (a.flat) {
    r.oY = c * n; r.oS = c * ns; r.oT = c * n + ns; r.oE = c; r.oEs = c; r.oF = c; r.oI = 0; r.sY = 1; r.sS = 1;
  } else {
    const int nx = a.hi[0] - a.lo[0] + 1, ny = a.hi[1] - a.lo[1] + 1;
    const int i = a.lo[0] + int(c % nx), j = a.lo[1] + int((c / nx) % ny), k = a.lo[2] + int(c / (int64_t(nx) * ny));
    r.oY = a4_off(a.rY, i, j, k); r.sY = a.rY.nstride;
    r.oS = a4_off(a.rYsrc, i, j, k); r.sS = a.rYsrc.nstride;
    r.oT = a4_off(a.T, i, j, k);
    r.oE = a4_off(a.rE, i, j, k);
    r.oEs = a4_off(a.rEsrc, i, j, k);
    r.oF = a4_off(a.FC, i, j, k);
    r.oI = a.IR.p ? a4_off(a.IR, i, j, k) : 0;
    if (a.mask.p) r.mk = a.mask.p[int64_t(i - a.mask.lo[0]) + int64_t(j - a.mask.lo[1]) * a.mask.jstride +
                                  int64_t(k - a.mask.lo[2]) * a.mask.kstride];
  }
  return r;
}

template <class L, int INTEG>
__global__ void __launch_bounds__(L::WARPS * 32, L::MINB)
    react_kernel(const DevMech* __restrict__ dm, const ReactOpts o, const __grid_constant__ BoxArgs a, unsigned long long* __restrict__ counter,
                 unsigned long long* __restrict__ gstats, int* __restrict__ status, double* __restrict__ jglobal)
{
  constexpr int NSLOT = L::NSLOT, ns = L::NS, n = L::N;
  double* wb = block_prologue<L>(dm);
  const DevMech& m = smem_mech();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * (blockDim.x >> 5) + warp;

  BdfOpts bo;
  bo.rtol = o.rtol; bo.atol = o.atol; bo.qmax = o.qmax; bo.maxcor = o.maxcor; bo.ajac = o.ajac;
  bo.mxstep = 10000; bo.maxnef = 100; bo.msbj = 100;
  BdfWarp<L> bdf(bo, lane, wb, jglobal + size_t(gwarp) * (n * L::LD + (L::POOL > 0 ? n * L::WLD : 0)));
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) {
    const int i = lane + 32 * s;
    bdf.atolv[s] = (i < n) ? (o.use_typ ? o.typ_vals[i] * o.atol : o.atol) : 1.0;
  }

  // statistics: a few atomics per cell (a cell is millions of instructions) instead of accumulators held in registers
  // across the integration
  const int64_t ncells = a.ncells;
  const int nwarps = gridDim.x * (blockDim.x >> 5);

#pragma unroll 1
  for (;;) {
    // kCellsPerGrab consecutive cells per grab (full 32-byte sectors of the state arrays) while plenty remain, single
    // cells near the end of the queue: cell cost varies 10x, and the tail of the call is set by the last grabs
    unsigned long long c0 = 0;
    int grab = kCellsPerGrab;
    if (lane == 0) {
      const unsigned long long seen = *reinterpret_cast<volatile unsigned long long*>(counter);
      if (int64_t(seen) + int64_t(nwarps) * (2 * kCellsPerGrab) >= ncells) grab = 1;
      c0 = atomicAdd(counter, (unsigned long long)grab);
    }
    c0 = __shfl_sync(kFull, c0, 0);
    grab = __shfl_sync(kFull, grab, 0);
    if (int64_t(c0) >= ncells) break;
#pragma unroll 1
    for (int cc = 0; cc < grab; ++cc) {
      const int64_t c = int64_t(c0) + cc;
      if (c >= ncells) break;
      CellAddr ad = cell_addr<L>(a, c);
      // the arrays of the cell's box: the call's own, or an entry of the box table of a multi-box call (plm_react_boxes)
      const BoxArgs* bxp = (ad.box >= 0) ? a.tab + ad.box : &a;
      if (INTEG == kIntegCvode && ad.mk == -1) {  // ReactorCvode.cpp:1395, :1439-1441 (RK64 ignores the mask)
        if (lane == 0) {
          bxp->FC.p[ad.oF] = 0.0;
          if (status) status[c] = 0;
          atomicAdd(gstats + 8, 1ull);
        }
        if (a.mks) {  // the reference still scales a skipped cell to CGS and back (PeleLM.cpp:5737-5772): same roundings
#pragma unroll
          for (int s = 0; s < NSLOT; ++s) {
            const int k = lane + 32 * s;
            if (k < ns) {
              double* dst = bxp->rY.p + ad.oY + int64_t(k) * ad.sY;
              const double vold = *dst, vnew = __dmul_rn(__dmul_rn(vold, 1.0e-3), 1.0e3);
              if (bxp->IR.p)
                bxp->IR.p[ad.oI + int64_t(k) * bxp->IR.nstride] =
                    __dsub_rn(__dmul_rn(-__dsub_rn(vold, vnew), a.dt_inv), bxp->rYsrc.p[ad.oS + int64_t(k) * ad.sS]);
              *dst = vnew;
            }
          }
          if (lane == 0) bxp->rE.p[ad.oE] = __dmul_rn(__dmul_rn(bxp->rE.p[ad.oE], 10.0), 0.1);
        }
        continue;
      }
      double y[NSLOT];
      const double Tin = bxp->T.p[ad.oT];
      const bool mks = a.mks != 0;
      // __dmul_rn / __dadd_rn: the scalings are separate, rounded operations in the reference (no contraction into FMAs)
      const double rhoe_init = mks ? __dmul_rn(bxp->rE.p[ad.oE], 10.0) : bxp->rE.p[ad.oE];
      const double rhoesrc = mks ? __dmul_rn(bxp->rEsrc.p[ad.oEs], 10.0) : bxp->rEsrc.p[ad.oEs];
      __syncwarp();
      const bool clean = (INTEG == kIntegCvode) && o.clean_init_massfrac && !a.flat;
      double rho_old = 0.0;
      {
        double ry0[NSLOT];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
          const int k = lane + 32 * s;
          ry0[s] = (k < ns) ? bxp->rY.p[ad.oY + int64_t(k) * ad.sY] : 0.0;
          double sv = (k < ns) ? bxp->rYsrc.p[ad.oS + int64_t(k) * ad.sS] : 0.0;
          if (mks) {  // PeleLM.cpp:5737-5746
            ry0[s] = __dmul_rn(ry0[s], 1.0e-3);
            sv = __dmul_rn(sv, 1.0e-3);
          }
          wb[L::SRC + k] = sv;
        }
        if (lane == 0) {
          wb[L::SM + 24] = rhoe_init;
          wb[L::SM + 25] = rhoesrc;
        }
        if (clean) {
          double p = 0.0;
#pragma unroll
          for (int s = 0; s < NSLOT; ++s) p += ry0[s];
          rho_old = warp_sum(p);
        }
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
          const int k = lane + 32 * s;
          y[s] = clean ? fmin(fmax(0.0, ry0[s]), rho_old) : ry0[s];
          if (k >= ns) y[s] = 0.0;
        }
      }
      double T0 = Tin, rho_oldclipped = 0.0;
      if (!a.flat) {  // box_flatten: T <- RHY2T(rho, rhoH/rho, Y, T_guess)
        double part = 0.0;
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) part += y[s];
        const double rho = warp_sum(part);
        rho_oldclipped = rho;
        const double rho_inv = 1.0 / rho;
        double Yw[NSLOT], cp;
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) Yw[s] = (lane + 32 * s < ns) ? (y[s] * rho_inv) * ldg(m.imw + lane + 32 * s) : 0.0;
        T0 = get_T_given_hY<L>(rhoe_init * rho_inv, Yw, Tin, cp);
      }
#pragma unroll
      for (int s = 0; s < NSLOT; ++s)
        if (lane + 32 * s == ns) y[s] = T0;
      if (lane == 0) {  // kept for box_unflatten in the scalar area (beyond the integrator's entries)
        wb[L::SM + 56] = rho_old;
        wb[L::SM + 57] = rho_oldclipped;
      }
      __syncwarp();

      CellStats st{};
      int fc;
      if (INTEG == kIntegCvode) {
        bdf.integrate(y, a.time, a.time + a.dt, st);
        fc = st.nfe;
      } else {
        int nfe = 0;
        fc = rk64_warp<L>(wb, lane, o, y, a.time, a.dt, nfe, !a.flat);
        st.nst = fc;
        st.nfe = nfe;
      }
      // box_unflatten (addresses recomputed: nothing of the cell's bookkeeping stays live across the integration)
      ad = cell_addr<L>(a, c);
      bxp = (ad.box >= 0) ? a.tab + ad.box : &a;
      const double rhoe0 = wb[L::SM + 24], rhoesrc0 = wb[L::SM + 25];
      double outv[NSLOT], p2 = 0.0;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        const int k = lane + 32 * s;
        outv[s] = y[s];
        if (clean) {  // ReactorUtils.H:160-166: the loaded and the clipped state, recomputed from the value still in the FAB
          double r0 = (k < ns) ? bxp->rY.p[ad.oY + int64_t(k) * ad.sY] : 0.0;
          if (mks) r0 = __dmul_rn(r0, 1.0e-3);
          const double cl = (k < ns) ? fmin(fmax(0.0, r0), wb[L::SM + 56]) : 0.0;
          outv[s] = r0 + (y[s] - cl) * wb[L::SM + 56] / wb[L::SM + 57];
        }
        if (k < ns) p2 += outv[s];
      }
      double Tout = 0.0;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s)
        if (lane + 32 * s == ns) Tout = y[s];
      Tout = __shfl_sync(kFull, Tout, ns & 31);
      const double rE_new = __dadd_rn(rhoe0, __dmul_rn(a.dt, rhoesrc0));  // unfused, as the host compiler evaluates ReactorUtils.H:180
      if (!a.flat) {
        const double rho = warp_sum(p2);
        const double rho_inv = 1.0 / rho;
        double Yw[NSLOT], cp;
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) Yw[s] = (lane + 32 * s < ns) ? (outv[s] * rho_inv) * ldg(m.imw + lane + 32 * s) : 0.0;
        Tout = get_T_given_hY<L>(rE_new * rho_inv, Yw, Tout, cp);
      }
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        const int k = lane + 32 * s;
        if (k < ns) {
          double* dst = bxp->rY.p + ad.oY + int64_t(k) * ad.sY;
          if (mks) {  // CGS -> MKS (PeleLM.cpp:5765-5772), then I_R (PeleLM.cpp:5795-5799) from the old value still in place
            const double vnew = __dmul_rn(outv[s], 1.0e3);
            if (bxp->IR.p) {
              const double vold = *dst;
              const double frc = bxp->rYsrc.p[ad.oS + int64_t(k) * ad.sS];
              bxp->IR.p[ad.oI + int64_t(k) * bxp->IR.nstride] = __dsub_rn(__dmul_rn(-__dsub_rn(vold, vnew), a.dt_inv), frc);
            }
            *dst = vnew;
          } else {
            *dst = outv[s];
          }
        }
      }
      if (lane == 0) {
        bxp->T.p[ad.oT] = Tout;
        bxp->rE.p[ad.oE] = mks ? __dmul_rn(rE_new, 0.1) : rE_new;
        bxp->FC.p[ad.oF] = double(fc);
        if (status) status[c] = st.flag;
        atomicAdd(gstats + 0, (unsigned long long)st.nst);
        atomicAdd(gstats + 1, (unsigned long long)st.nfe);
        if (st.nje) atomicAdd(gstats + 2, (unsigned long long)st.nje);
        if (st.nsetups) atomicAdd(gstats + 3, (unsigned long long)st.nsetups);
        if (st.nni) atomicAdd(gstats + 4, (unsigned long long)st.nni);
        if (st.ncfn) atomicAdd(gstats + 5, (unsigned long long)st.ncfn);
        if (st.netf) atomicAdd(gstats + 6, (unsigned long long)st.netf);
        if (st.flag != 0) atomicAdd(gstats + 7, 1ull);
        atomicAdd(gstats + 9, 1ull);
        atomicMax(gstats + 10, (unsigned long long)fc);
        atomicMax(gstats + 11, 0xffffffffull - (unsigned long long)fc);
      }
    }
  }
#ifdef PLM_ALIGN
  // phase-aligned experiment: a warp that is out of cells keeps meeting the barrier of BdfWarp::f until all warps are out
  if (INTEG == kIntegCvode)
    while (!__syncthreads_and(1)) {}
#elif defined(PLM_SOFTALIGN)
  if (INTEG == kIntegCvode) softbar_leave();  // out of cells: the others stop waiting for this warp
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// ReactorRK64, two cells per warp. The explicit integrator has no per-cell control flow beyond its step count, so a warp
// keeps two cells in flight (slots 0 and 1 of its shared-memory area) and advances both by one Runge-Kutta stage per
// trip: one call of rhs_warp2 evaluates both right-hand sides with interleaved dependency chains and shared table loads.
// A slot whose cell reaches the end of the interval stores it and takes the next cell from the queue. Per cell the
// arithmetic (and therefore the result) is exactly that of rk64_warp: ReactorRK64.cpp:206-317.
// ---------------------------------------------------------------------------------------------------------------
template <class L>
__global__ void __launch_bounds__(L::WARPS * 32, L::MINB)
    react_rk64x2_kernel(const DevMech* __restrict__ dm, const ReactOpts o, const BoxArgs a, unsigned long long* __restrict__ counter,
                        unsigned long long* __restrict__ gstats, int* __restrict__ status)
{
  static_assert(L::SLOTS == 2, "two cell areas per warp");
  constexpr int NSLOT = L::NSLOT, ns = L::NS, n = L::N;
  // per-slot integrator state: vectors in the (otherwise unused) BDF vector areas, scalars in the scalar area
  constexpr int SOLN = L::EWT, CARRY = L::ACOR, ERR = L::ATOL;
  enum { S_T = 26, S_DT = 27, S_I = 40 };  // doubles; ints at S_I: nsteps, stage, nfe, active, cell (2 ints)
  block_prologue<L>(dm);
  const DevMech& m = smem_mech();
  const int lane = threadIdx.x & 31;
  const int64_t ncells = a.ncells;
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  const double exp1 = 0.25, exp2 = 0.2, betaerr = 1.0, tinyval = 1e-50;
  const double time_init = a.time, time_out = a.time + a.dt;
  const double dt_rk_min = a.dt / double(o.rk64_max), dt_rk_max = a.dt / double(o.rk64_min);
  int64_t q_next = 0, q_end = 0;  // cells of the current grab not yet started
  bool exhausted = false;
#pragma unroll 1
  for (;;) {
    // ---- refill empty slots ----
#pragma unroll 1
    for (int sl = 0; sl < 2; ++sl) {
      double* w = warp_slot<L>(sl);
      int* wi = reinterpret_cast<int*>(w + L::SM + S_I);
      if (wi[3]) continue;  // active
      if (q_next >= q_end && !exhausted) {
        unsigned long long c0 = 0;
        int grab = kCellsPerGrab;
        if (lane == 0) {
          const unsigned long long seen = *reinterpret_cast<volatile unsigned long long*>(counter);
          if (int64_t(seen) + int64_t(nwarps) * (4 * kCellsPerGrab) >= ncells) grab = 1;
          c0 = atomicAdd(counter, (unsigned long long)grab);
        }
        c0 = __shfl_sync(kFull, c0, 0);
        grab = __shfl_sync(kFull, grab, 0);
        q_next = int64_t(c0);
        q_end = q_next + grab < ncells ? q_next + grab : ncells;
        if (q_next >= ncells) exhausted = true;
      }
      if (q_next >= q_end) continue;
      const int64_t c = q_next++;
      // box_flatten (ReactorUtils.H:76-136); ReactorRK64 ignores the mask (ReactorRK64.cpp:176)
      set_slot(sl);
      const CellAddr ad = cell_addr<L>(a, c);
      const bool mks = a.mks != 0;
      const double Tin = a.T.p[ad.oT];
      const double rhoe_init = mks ? __dmul_rn(a.rE.p[ad.oE], 10.0) : a.rE.p[ad.oE];
      const double rhoesrc = mks ? __dmul_rn(a.rEsrc.p[ad.oEs], 10.0) : a.rEsrc.p[ad.oEs];
      double y[NSLOT], part = 0.0;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        const int k = lane + 32 * s;
        y[s] = (k < ns) ? a.rY.p[ad.oY + int64_t(k) * ad.sY] : 0.0;
        double sv = (k < ns) ? a.rYsrc.p[ad.oS + int64_t(k) * ad.sS] : 0.0;
        if (mks) {
          y[s] = __dmul_rn(y[s], 1.0e-3);
          sv = __dmul_rn(sv, 1.0e-3);
        }
        w[L::SRC + k] = sv;
        part += y[s];
      }
      w[L::SM + 24] = rhoe_init;
      w[L::SM + 25] = rhoesrc;
      double T0 = Tin;
      if (!a.flat) {
        const double rho = warp_sum(part);
        const double rho_inv = 1.0 / rho;
        double Yw[NSLOT], cp;
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) Yw[s] = (lane + 32 * s < ns) ? (y[s] * rho_inv) * ldg(m.imw + lane + 32 * s) : 0.0;
        T0 = get_T_given_hY<L>(rhoe_init * rho_inv, Yw, Tin, cp);
      }
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        const int k = lane + 32 * s;
        if (k == ns) y[s] = T0;
        w[SOLN + k] = y[s];
        w[CARRY + k] = y[s];
        w[ERR + k] = 0.0;
      }
      w[L::SM + S_T] = time_init;
      w[L::SM + S_DT] = a.dt / double(o.rk64_guess);
      wi[0] = 0;  // nsteps
      wi[1] = 0;  // stage
      wi[2] = 0;  // nfe
      wi[3] = 1;  // active
      wi[4] = int(c & 0xffffffff);
      wi[5] = int(c >> 32);
      __syncwarp();
      if (!(time_init < time_out)) {  // empty interval: nothing to integrate (loop condition of ReactorRK64.cpp:232)
        wi[1] = 6;
      }
    }
    double* w0 = warp_slot<L>(0);
    double* w1 = warp_slot<L>(1);
    int* i0 = reinterpret_cast<int*>(w0 + L::SM + S_I);
    int* i1 = reinterpret_cast<int*>(w1 + L::SM + S_I);
    const bool act0 = i0[3] != 0, act1 = i1[3] != 0;
    if (!act0 && !act1) break;
    // ---- one stage of every active cell ----
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      if (act0) w0[L::YB + lane + 32 * s] = w0[SOLN + lane + 32 * s];
      if (act1) w1[L::YB + lane + 32 * s] = w1[SOLN + lane + 32 * s];
    }
    __syncwarp();
    const double tc0 = w0[L::SM + S_T], tc1 = w1[L::SM + S_T];
    bool both = false;
    if (act0 && act1) both = rhs_warp2<L>(tc0 - time_init, tc1 - time_init);
    if (!both) {
      if (act0) {
        set_slot(0);
        rhs_warp<L>(tc0 - time_init);
      }
      if (act1) {
        set_slot(1);
        rhs_warp<L>(tc1 - time_init);
      }
    }
    __syncwarp();
#pragma unroll 1
    for (int sl = 0; sl < 2; ++sl) {
      double* w = sl ? w1 : w0;
      int* wi = sl ? i1 : i0;
      if (!wi[3]) continue;
      const int stage = wi[1];
      double dt_rk = w[L::SM + S_DT];
      double soln[NSLOT];
      {
        const double ce = c_err_rk64[stage], ca = c_alpha_rk64[stage], cb = c_beta_rk64[stage];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
          const int k = lane + 32 * s;
          const double rhs = w[L::FB + k];
          w[ERR + k] += ce * dt_rk * rhs;
          soln[s] = w[CARRY + k] + ca * dt_rk * rhs;
          w[SOLN + k] = soln[s];
          w[CARRY + k] = soln[s] + cb * dt_rk * rhs;
        }
      }
      wi[2] += 1;
      if (stage < 5) {
        wi[1] = stage + 1;
        continue;
      }
      // ---- end of a step: error control (ReactorRK64.cpp:283-307) ----
      wi[1] = 0;
      double current_time = w[L::SM + S_T] + dt_rk;
      wi[0] += 1;
      double me = tinyval;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        const int k = lane + 32 * s;
        const double e = w[ERR + k];
        if (k < n) me = fabs(e) > me ? fabs(e) : me;
        w[ERR + k] = 0.0;
      }
      const double max_err = warp_max(me);
      if (max_err < o.atol) {
        const double cf = betaerr * dexp(exp1 * dlog(o.atol / max_err));
        dt_rk = fmin(dt_rk_max, dt_rk * cf);
      } else {
        const double cf = betaerr * dexp(exp2 * dlog(o.atol / max_err));
        dt_rk = fmax(dt_rk_min, dt_rk * cf);
      }
      if (!a.flat) dt_rk = fmin(dt_rk, time_out - current_time);
      w[L::SM + S_T] = current_time;
      w[L::SM + S_DT] = dt_rk;
      if (current_time < time_out) continue;
      // ---- cell finished: box_unflatten (ReactorUtils.H:138-207) ----
      set_slot(sl);
      const int64_t c = (int64_t(wi[5]) << 32) | int64_t(unsigned(wi[4]));
      const CellAddr ad = cell_addr<L>(a, c);
      const bool mks = a.mks != 0;
      const int fc = wi[0], nfe = wi[2];
      const double rhoe0 = w[L::SM + 24], rhoesrc0 = w[L::SM + 25];
      double p2 = 0.0;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s)
        if (lane + 32 * s < ns) p2 += soln[s];
      double Tout = 0.0;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s)
        if (lane + 32 * s == ns) Tout = soln[s];
      Tout = __shfl_sync(kFull, Tout, ns & 31);
      const double rE_new = __dadd_rn(rhoe0, __dmul_rn(a.dt, rhoesrc0));
      if (!a.flat) {
        const double rho = warp_sum(p2);
        const double rho_inv = 1.0 / rho;
        double Yw[NSLOT], cp;
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) Yw[s] = (lane + 32 * s < ns) ? (soln[s] * rho_inv) * ldg(m.imw + lane + 32 * s) : 0.0;
        Tout = get_T_given_hY<L>(rE_new * rho_inv, Yw, Tout, cp);
      }
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        const int k = lane + 32 * s;
        if (k < ns) {
          double* dst = a.rY.p + ad.oY + int64_t(k) * ad.sY;
          if (mks) {
            const double vnew = __dmul_rn(soln[s], 1.0e3);
            if (a.IR.p) {
              const double vold = *dst;
              const double frc = a.rYsrc.p[ad.oS + int64_t(k) * ad.sS];
              a.IR.p[ad.oI + int64_t(k) * a.IR.nstride] = __dsub_rn(__dmul_rn(-__dsub_rn(vold, vnew), a.dt_inv), frc);
            }
            *dst = vnew;
          } else {
            *dst = soln[s];
          }
        }
      }
      if (lane == 0) {
        a.T.p[ad.oT] = Tout;
        a.rE.p[ad.oE] = mks ? __dmul_rn(rE_new, 0.1) : rE_new;
        a.FC.p[ad.oF] = double(fc);
        if (status) status[c] = 0;
        atomicAdd(gstats + 0, (unsigned long long)fc);
        atomicAdd(gstats + 1, (unsigned long long)nfe);
        atomicAdd(gstats + 9, 1ull);
        atomicMax(gstats + 10, (unsigned long long)fc);
        atomicMax(gstats + 11, 0xffffffffull - (unsigned long long)fc);
      }
      wi[3] = 0;
      __syncwarp();
    }
  }
}

template <class L>
__global__ void __launch_bounds__(L::WARPS * 32) rhs_kernel(const DevMech* __restrict__ dm, int ncells, double t,
                                                            const double* __restrict__ y,
                                                            const double* __restrict__ rhoe,
                                                            const double* __restrict__ rhoesrc,
                                                            const double* __restrict__ rYsrc, double* __restrict__ ydot)
{
  constexpr int NSLOT = L::NSLOT, ns = L::NS, n = L::N;
  double* wb = block_prologue<L>(dm);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = gridDim.x * (blockDim.x >> 5);
  for (int c = blockIdx.x * (blockDim.x >> 5) + warp; c < ncells; c += nw) {
    __syncwarp();
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      const int i = lane + 32 * s;
      wb[L::YB + i] = (i < n) ? y[size_t(c) * n + i] : 0.0;
      wb[L::SRC + i] = (i < ns) ? rYsrc[size_t(c) * ns + i] : 0.0;
    }
    if (lane == 0) {
      wb[L::SM + 24] = rhoe[c];
      wb[L::SM + 25] = rhoesrc[c];
    }
    __syncwarp();
    rhs_warp<L>(t);
#pragma unroll
    for (int s = 0; s < NSLOT; ++s)
      if (lane + 32 * s < n) ydot[size_t(c) * n + lane + 32 * s] = wb[L::FB + lane + 32 * s];
  }
}

template <class L>
__global__ void __launch_bounds__(L::WARPS * 32) jac_kernel(const DevMech* __restrict__ dm, int ncells,
                                                            const double* __restrict__ y, double* __restrict__ Jout)
{
  constexpr int NSLOT = L::NSLOT, n = L::N, ld = L::LD;
  double* wb = block_prologue<L>(dm);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = gridDim.x * (blockDim.x >> 5);
  for (int c = blockIdx.x * (blockDim.x >> 5) + warp; c < ncells; c += nw) {
    __syncwarp();
#pragma unroll
    for (int s = 0; s < NSLOT; ++s)
      wb[L::YB + lane + 32 * s] = (lane + 32 * s < n) ? y[size_t(c) * n + lane + 32 * s] : 0.0;
    __syncwarp();
    int pool = 0;
    double* A;
    if constexpr (L::POOL > 0) {
      pool = pool_acquire<L>();
      A = pool_matrix<L>(pool);
    } else {
      A = wb + L::A;
    }
    jac_warp<L>(A);
    // column-major output like the SUNDenseMatrix handed to CVODE
    for (int e = lane; e < n * n; e += 32) {
      const int col = e / n, row = e - col * n;
      Jout[size_t(c) * n * n + e] = A[row * ld + col];
    }
    __syncwarp();
    if constexpr (L::POOL > 0) pool_release<L>(pool);
  }
}

// PeleLM_K.H:57-101 reactionRateRhoY, one warp per cell: rho = sum rhoY (MKS), Y, T <- HY2T(rhoH/rho * 1e4, Y; guess T),
// wdot*W = RTY2WDOT(rho * 1e-3, T, Y) (Fuego.H:168-183: C_k = rho Y_k / W_k, CKWC, x W_k), rhoYdot = x 1000.
template <class L>
__global__ void __launch_bounds__(L::WARPS * 32) reaction_rate_kernel(const DevMech* __restrict__ dm, const BoxArgs a,
                                                                      const plm_a4d out)
{
  constexpr int NSLOT = L::NSLOT, ns = L::NS;
  double* wb = block_prologue<L>(dm);
  const DevMech& m = smem_mech();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = gridDim.x * (blockDim.x >> 5);
  const int nx = a.hi[0] - a.lo[0] + 1, ny = a.hi[1] - a.lo[1] + 1;
  for (int64_t c = blockIdx.x * (blockDim.x >> 5) + warp; c < a.ncells; c += nw) {
    const int i = a.lo[0] + int(c % nx), j = a.lo[1] + int((c / nx) % ny), k3 = a.lo[2] + int(c / (int64_t(nx) * ny));
    const int64_t oY = a4_off(a.rY, i, j, k3), oO = a4_off(out, i, j, k3);
    int mk = 1;
    if (a.mask.p) mk = a.mask.p[int64_t(i - a.mask.lo[0]) + int64_t(j - a.mask.lo[1]) * a.mask.jstride +
                                int64_t(k3 - a.mask.lo[2]) * a.mask.kstride];
    __syncwarp();
    if (mk < 0) {  // EB covered cell: no reactions
#pragma unroll
      for (int s = 0; s < NSLOT; ++s)
        if (lane + 32 * s < ns) out.p[oO + int64_t(lane + 32 * s) * out.nstride] = 0.0;
      continue;
    }
    double ry[NSLOT], part = 0.0;
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      const int k = lane + 32 * s;
      ry[s] = (k < ns) ? a.rY.p[oY + int64_t(k) * a.rY.nstride] : 0.0;
      part += ry[s];
    }
    double rho = warp_sum(part);
    const double rhoinv = 1.0 / rho;
    double Y[NSLOT], Yw[NSLOT], imw[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      const int k = lane + 32 * s;
      imw[s] = ldg(m.imw + k);
      Y[s] = ry[s] * rhoinv;
      Yw[s] = (k < ns) ? Y[s] * imw[s] : 0.0;
    }
    const double H = a.rE.p[a4_off(a.rE, i, j, k3)] * rhoinv * 1.0e4;
    double cp;
    const double T = get_T_given_hY<L>(H, Yw, a.T.p[a4_off(a.T, i, j, k3)], cp);
    rho = rho * 0.001;
    const double logT = dlog(T), invT = 1.0 / T;
    double cpart = 0.0;
    double2* cg = reinterpret_cast<double2*>(wb + L::CG);
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      const int k = lane + 32 * s;
      if (k < ns) {
        const double cc = ((rho * Y[s]) * imw[s]) * 1.0e6;  // CKYTCR (mechanism.H:2403) then CKWC's SI conversion (:4911)
        cg[k] = make_double2(cc, gibbs_k<L>(m, k, T, logT, invT));
        cpart += cc;
      }
    }
    const double mixture = warp_sum(cpart);
    __syncwarp();
    eval_rates<L>(m, wb, lane, T, logT, invT, mixture);
    __syncwarp();
    double wdot[NSLOT];
    gather_wdot<L>(m, wb, lane, wdot);
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      const int k = lane + 32 * s;
      if (k < ns) out.p[oO + int64_t(k) * out.nstride] = ((wdot[s] * 1.0e-6) * ldg(m.mw + k)) * 1000.0;
    }
  }
}

__global__ void box_cost_kernel(const double* __restrict__ FC, const int64_t* __restrict__ off, int nboxes,
                                long long* __restrict__ costs)
{
  const int b = blockIdx.x;
  if (b >= nboxes) return;
  double s = 0.0;
  for (int64_t c = off[b] + threadIdx.x; c < off[b + 1]; c += blockDim.x) s += FC[c];
  __shared__ double red[32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) costs[b] = (long long)(v + 0.5);
  }
}

// Register-resident DFMA throughput probe: the measured FP64 SIMT peak that the roofline fractions are
// quoted against (MEASURED_PEAKS.json has no FP64 entry). 16 independent chains per thread.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* __restrict__ out, int iters, double a, double b)
{
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = double(threadIdx.x + i) * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // keeps the chains live, never true in practice
}

cudaError_t launch_dfma_peak(double* scratch, int blocks, int iters, cudaStream_t stream)
{
  dfma_peak_kernel<<<blocks, 256, 0, stream>>>(scratch, iters, 0.999999, 1e-9);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// launchers: dispatch on (mechanism, order class, integrator)
// ------------------------------------------------------------------------------------------------
using LLi2 = Lay<9, 21, 2>;    using LLi5 = Lay<9, 21, 5>;    using LLi0 = Lay<9, 21, 0>;    using LLiD = Lay<9, 21, 0, 2>;
using LDr2 = Lay<21, 84, 2>;   using LDr5 = Lay<21, 84, 5>;   using LDr0 = Lay<21, 84, 0>;   using LDrD = Lay<21, 84, 0, 2>;
using LDo2 = Lay<53, 268, 2>;  using LDo5 = Lay<53, 268, 5>;  using LDo0 = Lay<53, 268, 0>;  using LDoD = Lay<53, 268, 0, 2>;

// ReactorRK64 has two kernels: one cell per warp (react_kernel, the default) and two cells per warp (react_rk64x2_kernel
// on the SLOTS == 2 layouts; PLM_RK64_DUAL=1). The second executes 17 % fewer instructions per cell and gives every warp
// two independent dependency chains, and produces bit-identical results (tests/test_gpu_parity.py), but its larger
// loop body does not fit the instruction cache (hit rate 77 %, GPC instruction cache at 98 % of its bandwidth) and it
// measures 15-20 % slower on B200: kept as the prototype of the two-cell right-hand side the BDF kernel needs next.
static bool rk64_dual()
{
  static const bool v = [] {
    const char* e = std::getenv("PLM_RK64_DUAL");
    return e != nullptr && e[0] == '1';
  }();
  return v;
}

// calls F.template run<Lay, INTEG>() for the matching instantiation
template <class F>
static cudaError_t dispatch(const DevMech& m, int integ, int qmax, F&& f)
{
  const int qc = (integ == kIntegRK64) ? 0 : (qmax <= 2 ? 2 : 5);
#define PLM_CASE(NS, NR, L0, LD, L2, L5)                                                   \
  if (m.ns == NS && m.nr == NR) {                                                          \
    if (qc == 0 && rk64_dual()) return f.template run<LD, kIntegRK64>();                   \
    if (qc == 0) return f.template run<L0, kIntegRK64>();                                  \
    if (qc == 2) return f.template run<L2, kIntegCvode>();                                 \
    return f.template run<L5, kIntegCvode>();                                              \
  }
  PLM_CASE(9, 21, LLi0, LLiD, LLi2, LLi5)
  PLM_CASE(21, 84, LDr0, LDrD, LDr2, LDr5)
  PLM_CASE(53, 268, LDo0, LDoD, LDo2, LDo5)
#undef PLM_CASE
  return cudaErrorInvalidValue;
}

// dynamic shared memory of every kernel instantiated on layout L: header + warp areas + (TSM) the staged tables
template <class L>
static size_t smem_bytes_of(const DevMech& m)
{
  return (size_t(kHdr) + size_t(L::WARPS) * L::SLOTS * L::TOTAL + size_t(L::POOLD) + size_t(Tab<L>::doubles(m.emax, m.glen))) *
         sizeof(double);
}

struct SmemQuery {
  const DevMech* mech;
  size_t* out;
  template <class L, int IG>
  cudaError_t run()
  {
    *out = smem_bytes_of<L>(*mech);
    return cudaSuccess;
  }
};
size_t react_smem_bytes(const DevMech& m, int integ, int qmax, int warps_per_block)
{
  size_t v = 0;
  (void)warps_per_block;
  SmemQuery q{&m, &v};
  dispatch(m, integ, qmax, q);
  return v;
}

struct ThreadsQuery {
  int* out;
  template <class L, int IG>
  cudaError_t run()
  {
    *out = L::WARPS * 32;
    return cudaSuccess;
  }
};
int react_threads(const DevMech& m, int integ, int qmax)
{
  int v = 0;
  ThreadsQuery q{&v};
  dispatch(m, integ, qmax, q);
  return v;
}

struct JQuery {
  size_t* out;
  template <class L, int IG>
  cudaError_t run()
  {
    *out = size_t(L::N) * L::LD + (L::POOL > 0 ? size_t(L::N) * L::WLD : 0);  // saved Jacobian (+ the transposed inverse)
    return cudaSuccess;
  }
};
size_t react_jsave_doubles(const DevMech& m)
{
  size_t v = 0;
  JQuery q{&v};
  dispatch(m, kIntegCvode, 2, q);
  return v;
}

struct OccQuery {
  int threads;
  size_t smem;
  int* nb;
  template <class L, int IG>
  cudaError_t run()
  {
    if constexpr (IG == kIntegRK64 && L::SLOTS == 2) {
      auto k = react_rk64x2_kernel<L>;
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      if (e != cudaSuccess) return e;
      return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, k, threads, smem);
    } else {
      auto k = react_kernel<L, IG>;
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      if (e != cudaSuccess) return e;
      return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, k, threads, smem);
    }
  }
};
int react_occupancy(const DevMech& m, int integ, int qmax, int threads, size_t smem)
{
  int nb = 0;
  OccQuery q{threads, smem, &nb};
  return dispatch(m, integ, qmax, q) == cudaSuccess ? nb : -1;
}

struct ReactLaunch {
  const DevMech* dm;
  const ReactOpts& o;
  const BoxArgs& a;
  const LaunchCfg& cfg;
  unsigned long long *counter, *gstats;
  int* status;
  double* jglobal;
  cudaStream_t stream;
  template <class L, int IG>
  cudaError_t run()
  {
    if constexpr (IG == kIntegRK64 && L::SLOTS == 2) {
      auto kern = react_rk64x2_kernel<L>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(cfg.smem_bytes));
      if (e != cudaSuccess) return e;
      kern<<<cfg.blocks, cfg.threads, cfg.smem_bytes, stream>>>(dm, o, a, counter, gstats, status);
    } else {
      auto kern = react_kernel<L, IG>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(cfg.smem_bytes));
      if (e != cudaSuccess) return e;
      kern<<<cfg.blocks, cfg.threads, cfg.smem_bytes, stream>>>(dm, o, a, counter, gstats, status, jglobal);
    }
    return cudaGetLastError();
  }
};
cudaError_t launch_react(const DevMech& m, const DevMech* dm, const ReactOpts& o, const BoxArgs& a, const LaunchCfg& cfg,
                         int integ, unsigned long long* counter, unsigned long long* gstats, int* status,
                         double* jglobal, cudaStream_t stream)
{
  ReactLaunch l{dm, o, a, cfg, counter, gstats, status, jglobal, stream};
  return dispatch(m, integ, o.qmax, l);
}

struct RhsLaunch {
  const DevMech* mech;  // host copy (sizes)
  const DevMech* dm;
  int ncells;
  double t;
  const double *y, *rhoe, *rhoesrc, *rYsrc;
  double* ydot;
  cudaStream_t stream;
  template <class L, int IG>
  cudaError_t run()
  {
    const int wpb = L::WARPS;
    const size_t smem = smem_bytes_of<L>(*mech);
    const int blocks = std::max(1, std::min((ncells + wpb - 1) / wpb, 148 * 4));
    cudaError_t e = cudaFuncSetAttribute(rhs_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    rhs_kernel<L><<<blocks, L::WARPS * 32, smem, stream>>>(dm, ncells, t, y, rhoe, rhoesrc, rYsrc, ydot);
    return cudaGetLastError();
  }
};
cudaError_t launch_rhs(const DevMech& m, const DevMech* dm, int ncells, double t, const double* y, const double* rhoe,
                       const double* rhoesrc, const double* rYsrc, double* ydot, cudaStream_t stream)
{
  RhsLaunch l{&m, dm, ncells, t, y, rhoe, rhoesrc, rYsrc, ydot, stream};
  return dispatch(m, kIntegCvode, 2, l);
}

struct JacLaunch {
  const DevMech* mech;
  const DevMech* dm;
  int ncells;
  const double* y;
  double* J;
  cudaStream_t stream;
  template <class L, int IG>
  cudaError_t run()
  {
    const int wpb = L::WARPS;
    const size_t smem = smem_bytes_of<L>(*mech);
    const int blocks = std::max(1, std::min((ncells + wpb - 1) / wpb, 148 * 4));
    cudaError_t e = cudaFuncSetAttribute(jac_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    jac_kernel<L><<<blocks, L::WARPS * 32, smem, stream>>>(dm, ncells, y, J);
    return cudaGetLastError();
  }
};
cudaError_t launch_jac(const DevMech& m, const DevMech* dm, int ncells, const double* y, double* J, cudaStream_t stream)
{
  JacLaunch l{&m, dm, ncells, y, J, stream};
  return dispatch(m, kIntegCvode, 2, l);
}

struct RateLaunch {
  const DevMech* mech;
  const DevMech* dm;
  const BoxArgs& a;
  plm_a4d out;
  cudaStream_t stream;
  template <class L, int IG>
  cudaError_t run()
  {
    const int wpb = L::WARPS;
    const size_t smem = smem_bytes_of<L>(*mech);
    const int64_t want = (a.ncells + wpb - 1) / wpb;
    const int blocks = int(std::max<int64_t>(1, std::min<int64_t>(want, 148 * 8)));
    cudaError_t e = cudaFuncSetAttribute(reaction_rate_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    reaction_rate_kernel<L><<<blocks, L::WARPS * 32, smem, stream>>>(dm, a, out);
    return cudaGetLastError();
  }
};
cudaError_t launch_reaction_rate(const DevMech& m, const DevMech* dm, const BoxArgs& a, plm_a4d rhoYdot, cudaStream_t stream)
{
  RateLaunch l{&m, dm, a, rhoYdot, stream};
  return dispatch(m, kIntegCvode, 2, l);
}

cudaError_t launch_box_costs(const double* FC, const int64_t* off, int nboxes, long long* costs, cudaStream_t stream)
{
  box_cost_kernel<<<nboxes, 256, 0, stream>>>(FC, off, nboxes, costs);
  return cudaGetLastError();
}

}  // namespace plm
