// sm_100a kernels behind the C ABI (include/plm_chem.h).
//
//   react_kernel<NSLOT, INTEG>  persistent warps pull cells from a global queue (per-cell cost varies
//                               10x inside a box, SURVEY.md section 7 "Divergence"): gather the cell from
//                               the FAB (box_flatten, ReactorUtils.H:76-136), invert h -> T, integrate
//                               (BDF: bdf_warp.cuh, RK64: below), scatter back (box_unflatten :138-207)
//   rhs_kernel / jac_kernel     one warp per cell, for the fKernelSpec / cJac parity tests
//   box_cost_kernel             per-box sum of FC (PeleLM::getFuncCountDM PeleLM.cpp:5557-5571)
#include <cuda_runtime.h>

#include <cstdint>

#include "bdf_warp.cuh"
#include "kernels.h"

namespace plm {

// ---- ReactorRK64 (ReactorRK64.cpp:206-317; coefficients ReactorRK64.H:11-42), one warp per cell ----
__constant__ double c_alpha_rk64[6] = {0.218150805229859, 0.256702469801519, 0.527402592007520,
                                       0.0484864267224467, 1.24517071533530, 0.412366034843237};
__constant__ double c_beta_rk64[6] = {-0.113554138044166, -0.215118587818400, -0.0510152146250577,
                                      -1.07992686223881, -0.248664241213447, 0.0};
__constant__ double c_err_rk64[6] = {-0.0554699315064507, 0.158481845574980, -0.0905918835751907,
                                     -0.219084567203338, 0.164022338959433, 0.0426421977505659};

template <int NSLOT>
__device__ __forceinline__ int rk64_warp(const DevMech& m, const WarpMem& wm, int lane, const ReactOpts& o,
                                         double (&soln)[NSLOT], const double (&src)[NSLOT], double rhoe_init,
                                         double rhoesrc, double time_init, double dt_react, int& nfe)
{
  const double exp1 = 0.25, exp2 = 0.2, betaerr = 1.0, tinyval = 1e-50;
  const int n = m.n;
  double carry[NSLOT], err[NSLOT], rhs[NSLOT];
  const double time_out = time_init + dt_react;
  double current_time = time_init;
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) carry[s] = soln[s];
  double dt_rk = dt_react / double(o.rk64_guess);
  const double dt_rk_min = dt_react / double(o.rk64_max);
  const double dt_rk_max = dt_react / double(o.rk64_min);
  int nsteps = 0;
  while (current_time < time_out) {
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) err[s] = 0.0;
    for (int stage = 0; stage < 6; stage++) {
      rhs_warp<NSLOT>(m, wm, lane, current_time - time_init, soln, src, rhoe_init, rhoesrc, rhs);
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
      const double cf = betaerr * exp(exp1 * log(o.atol / max_err));
      dt_rk = fmin(dt_rk_max, dt_rk * cf);
    } else {
      const double cf = betaerr * exp(exp2 * log(o.atol / max_err));
      dt_rk = fmax(dt_rk_min, dt_rk * cf);
    }
    dt_rk = fmin(dt_rk, time_out - current_time);
  }
  return nsteps;
}

// carve the warp's shared-memory area
__device__ __forceinline__ WarpMem carve(double* base, const DevMech& m, int qmax, int ld, double* jglobal)
{
  WarpMem wm;
  double* p = base;
  wm.sc = p; p += m.nsp_pad;
  wm.gk = p; p += m.nsp_pad;
  wm.hk = p; p += m.nsp_pad;
  wm.q = p; p += m.nr_pad;
  wm.A = p; p += size_t(m.n) * ld;
  if (jglobal) {
    wm.J = jglobal;
  } else {
    wm.J = p; p += size_t(m.n) * ld;
  }
  wm.zn = p; p += size_t(qmax + 1) * m.nsp_pad;
  wm.sm = p; p += 32;
  wm.perm = reinterpret_cast<int*>(p);
  wm.ld = ld;
  return wm;
}

__device__ __forceinline__ void init_pads(const DevMech& m, const WarpMem& wm, int lane)
{
  for (int k = m.ns + lane; k < m.nsp_pad; k += 32) {
    wm.sc[k] = 1.0;  // unit concentration for unused reactant/product slots
    wm.gk[k] = 0.0;
    wm.hk[k] = 0.0;
  }
  __syncwarp();
}

__device__ __forceinline__ int64_t a4_off(const plm_a4d& a, int i, int j, int k)
{
  return int64_t(i - a.lo[0]) + int64_t(j - a.lo[1]) * a.jstride + int64_t(k - a.lo[2]) * a.kstride;
}

template <int NSLOT, int INTEG>
__global__ void __launch_bounds__(kReactThreads) react_kernel(const DevMech m, const ReactOpts o, const BoxArgs a,
                                                              unsigned long long* __restrict__ counter,
                                                              unsigned long long* __restrict__ gstats,
                                                              int* __restrict__ status, double* __restrict__ jglobal)
{
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ns = m.ns, n = m.n;
  const size_t wdoubles = warp_mem_doubles(n, m.nsp_pad, m.nr_pad, o.qmax, a.ld, jglobal == nullptr);
  const int gwarp = blockIdx.x * (blockDim.x >> 5) + warp;
  double* jg = jglobal ? jglobal + size_t(gwarp) * n * a.ld : nullptr;
  const WarpMem wm = carve(smem + warp * wdoubles, m, o.qmax, a.ld, jg);
  init_pads(m, wm, lane);

  BdfOpts bo;
  bo.rtol = o.rtol; bo.atol = o.atol; bo.qmax = o.qmax; bo.maxcor = o.maxcor; bo.ajac = o.ajac;
  bo.use_typ = o.use_typ; bo.mxstep = 10000; bo.maxnef = 100; bo.msbj = 100;
  BdfWarp<NSLOT> bdf(m, wm, bo, lane);
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) {
    const int i = lane + 32 * s;
    bdf.atolv[s] = (i < n) ? (o.use_typ ? o.typ_vals[i] * o.atol : o.atol) : 1.0;
  }

  unsigned long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // nst nfe nje nsetups nni ncfn netf nfailed
  unsigned long long nskipped = 0, ndone = 0;
  int maxfc = 0, minfc = 0x7fffffff;
  const int nx = a.hi[0] - a.lo[0] + 1, ny = a.hi[1] - a.lo[1] + 1;
  const int64_t ncells = a.ncells;

  for (;;) {
    unsigned long long c0 = 0;
    if (lane == 0) c0 = atomicAdd(counter, (unsigned long long)kCellsPerGrab);
    c0 = __shfl_sync(kFull, c0, 0);
    if (int64_t(c0) >= ncells) break;
    for (int cc = 0; cc < kCellsPerGrab; ++cc) {
      const int64_t c = int64_t(c0) + cc;
      if (c >= ncells) break;
      // cell -> (i,j,k) or flat index
      int64_t oY, oS, oT, oE, oEs, oF, sY, sS;
      int mk = 1;
      if (a.flat) {
        oY = c * n; oS = c * ns; oT = c * n + ns; oE = c; oEs = c; oF = c; sY = 1; sS = 1;
      } else {
        const int i = a.lo[0] + int(c % nx), j = a.lo[1] + int((c / nx) % ny), k = a.lo[2] + int(c / (int64_t(nx) * ny));
        oY = a4_off(a.rY, i, j, k); sY = a.rY.nstride;
        oS = a4_off(a.rYsrc, i, j, k); sS = a.rYsrc.nstride;
        oT = a4_off(a.T, i, j, k);
        oE = a4_off(a.rE, i, j, k);
        oEs = a4_off(a.rEsrc, i, j, k);
        oF = a4_off(a.FC, i, j, k);
        if (a.mask.p) mk = a.mask.p[int64_t(i - a.mask.lo[0]) + int64_t(j - a.mask.lo[1]) * a.mask.jstride +
                                    int64_t(k - a.mask.lo[2]) * a.mask.kstride];
      }
      if (INTEG == kIntegCvode && mk == -1) {  // ReactorCvode.cpp:1395, :1439-1441 (RK64 ignores the mask)
        if (lane == 0) {
          a.FC.p[oF] = 0.0;
          if (status) status[c] = 0;
        }
        nskipped++;
        continue;
      }
      double ry0[NSLOT], y[NSLOT];
      double Tin = a.T.p[oT];
      const double rhoe_init = a.rE.p[oE], rhoesrc = a.rEsrc.p[oEs];
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        const int k = lane + 32 * s;
        ry0[s] = (k < ns) ? a.rY.p[oY + int64_t(k) * sY] : 0.0;
        bdf.src[s] = (k < ns) ? a.rYsrc.p[oS + int64_t(k) * sS] : 0.0;
      }
      const bool clean = (INTEG == kIntegCvode) && o.clean_init_massfrac && !a.flat;
      double rho_old = 0.0, rho_oldclipped = 0.0, clipped[NSLOT];
      if (clean) {
        double p = 0.0;
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) p += ry0[s];
        rho_old = warp_sum(p);
      }
      double part = 0.0;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        const int k = lane + 32 * s;
        y[s] = clean ? fmin(fmax(0.0, ry0[s]), rho_old) : ry0[s];
        if (k >= ns) y[s] = 0.0;
        clipped[s] = y[s];
        part += y[s];
      }
      double T0 = Tin;
      if (!a.flat) {  // box_flatten: T <- RHY2T(rho, rhoH/rho, Y, T_guess)
        const double rho = warp_sum(part);
        rho_oldclipped = rho;
        const double rho_inv = 1.0 / rho;
        double Yw[NSLOT], hrt[NSLOT], cpr[NSLOT], cp;
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) Yw[s] = (lane + 32 * s < ns) ? (y[s] * rho_inv) * m.imw[lane + 32 * s] : 0.0;
        T0 = get_T_given_hY<NSLOT>(m, lane, rhoe_init * rho_inv, Yw, Tin, hrt, cpr, cp);
      }
#pragma unroll
      for (int s = 0; s < NSLOT; ++s)
        if (lane + 32 * s == ns) y[s] = T0;

      CellStats st{};
      int fc;
      if (INTEG == kIntegCvode) {
        bdf.rhoe_init = rhoe_init;
        bdf.rhoesrc = rhoesrc;
        bdf.integrate(y, a.time, a.time + a.dt, st);
        fc = st.nfe;
      } else {
        int nfe = 0;
        fc = rk64_warp<NSLOT>(m, wm, lane, o, y, bdf.src, rhoe_init, rhoesrc, a.time, a.dt, nfe);
        st.nst = fc;
        st.nfe = nfe;
      }
      // box_unflatten
      double outv[NSLOT], p2 = 0.0;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        const int k = lane + 32 * s;
        outv[s] = clean ? ry0[s] + (y[s] - clipped[s]) * rho_old / rho_oldclipped : y[s];
        if (k < ns) p2 += outv[s];
      }
      double Tout = 0.0;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s)
        if (lane + 32 * s == ns) Tout = y[s];
      Tout = __shfl_sync(kFull, Tout, ns & 31);
      const double rE_new = rhoe_init + a.dt * rhoesrc;
      if (!a.flat) {
        const double rho = warp_sum(p2);
        const double rho_inv = 1.0 / rho;
        double Yw[NSLOT], hrt[NSLOT], cpr[NSLOT], cp;
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) Yw[s] = (lane + 32 * s < ns) ? (outv[s] * rho_inv) * m.imw[lane + 32 * s] : 0.0;
        Tout = get_T_given_hY<NSLOT>(m, lane, rE_new * rho_inv, Yw, Tout, hrt, cpr, cp);
      }
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        const int k = lane + 32 * s;
        if (k < ns) a.rY.p[oY + int64_t(k) * sY] = outv[s];
      }
      if (lane == 0) {
        a.T.p[oT] = Tout;
        a.rE.p[oE] = rE_new;
        a.FC.p[oF] = double(fc);
        if (status) status[c] = st.flag;
      }
      acc[0] += st.nst; acc[1] += st.nfe; acc[2] += st.nje; acc[3] += st.nsetups;
      acc[4] += st.nni; acc[5] += st.ncfn; acc[6] += st.netf; acc[7] += (st.flag != 0);
      maxfc = max(maxfc, fc);
      minfc = min(minfc, fc);
      ndone++;
    }
  }
  if (lane == 0) {
    for (int i = 0; i < 8; ++i)
      if (acc[i]) atomicAdd(gstats + i, acc[i]);
    if (nskipped) atomicAdd(gstats + 8, nskipped);
    if (ndone) {
      atomicAdd(gstats + 9, ndone);
      atomicMax(gstats + 10, (unsigned long long)maxfc);
      atomicMax(gstats + 11, 0xffffffffull - (unsigned long long)minfc);
    }
  }
}

template <int NSLOT>
__global__ void __launch_bounds__(kReactThreads) rhs_kernel(const DevMech m, int ncells, double t,
                                                            const double* __restrict__ y,
                                                            const double* __restrict__ rhoe,
                                                            const double* __restrict__ rhoesrc,
                                                            const double* __restrict__ rYsrc, double* __restrict__ ydot,
                                                            int ld)
{
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t wdoubles = warp_mem_doubles(m.n, m.nsp_pad, m.nr_pad, 0, ld, true);
  const WarpMem wm = carve(smem + warp * wdoubles, m, 0, ld, nullptr);
  init_pads(m, wm, lane);
  const int nw = gridDim.x * (blockDim.x >> 5);
  for (int c = blockIdx.x * (blockDim.x >> 5) + warp; c < ncells; c += nw) {
    double yv[NSLOT], src[NSLOT], yd[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      const int i = lane + 32 * s;
      yv[s] = (i < m.n) ? y[size_t(c) * m.n + i] : 0.0;
      src[s] = (i < m.ns) ? rYsrc[size_t(c) * m.ns + i] : 0.0;
    }
    rhs_warp<NSLOT>(m, wm, lane, t, yv, src, rhoe[c], rhoesrc[c], yd);
#pragma unroll
    for (int s = 0; s < NSLOT; ++s)
      if (lane + 32 * s < m.n) ydot[size_t(c) * m.n + lane + 32 * s] = yd[s];
  }
}

template <int NSLOT>
__global__ void __launch_bounds__(kReactThreads) jac_kernel(const DevMech m, int ncells, const double* __restrict__ y,
                                                            double* __restrict__ Jout, int ld)
{
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t wdoubles = warp_mem_doubles(m.n, m.nsp_pad, m.nr_pad, 0, ld, true);
  const WarpMem wm = carve(smem + warp * wdoubles, m, 0, ld, nullptr);
  init_pads(m, wm, lane);
  const int nw = gridDim.x * (blockDim.x >> 5);
  const int n = m.n;
  for (int c = blockIdx.x * (blockDim.x >> 5) + warp; c < ncells; c += nw) {
    double yv[NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) yv[s] = (lane + 32 * s < n) ? y[size_t(c) * n + lane + 32 * s] : 0.0;
    jac_warp<NSLOT>(m, wm, lane, yv, wm.A);
    // column-major output like the SUNDenseMatrix handed to CVODE
    for (int e = lane; e < n * n; e += 32) {
      const int col = e / n, row = e - col * n;
      Jout[size_t(c) * n * n + e] = wm.A[row * ld + col];
    }
    __syncwarp();
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
// launchers
// ------------------------------------------------------------------------------------------------
static int nslot_of(const DevMech& m) { return (m.n + 31) / 32; }

int ld_for(const DevMech& m) { return (m.n % 2 == 0) ? m.n + 1 : m.n; }

size_t react_smem_bytes(const DevMech& m, int qmax, bool j_in_smem, int warps_per_block)
{
  return warp_mem_doubles(m.n, m.nsp_pad, m.nr_pad, qmax, ld_for(m), j_in_smem) * sizeof(double) * warps_per_block;
}

template <int NSLOT, int INTEG>
static cudaError_t launch_react_t(const DevMech& m, const ReactOpts& o, const BoxArgs& a, const LaunchCfg& cfg,
                                  unsigned long long* counter, unsigned long long* gstats, int* status,
                                  double* jglobal, cudaStream_t stream)
{
  auto kern = react_kernel<NSLOT, INTEG>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(cfg.smem_bytes));
  if (e != cudaSuccess) return e;
  kern<<<cfg.blocks, cfg.threads, cfg.smem_bytes, stream>>>(m, o, a, counter, gstats, status, jglobal);
  return cudaGetLastError();
}

cudaError_t launch_react(const DevMech& m, const ReactOpts& o, const BoxArgs& a, const LaunchCfg& cfg, int integ,
                         unsigned long long* counter, unsigned long long* gstats, int* status, double* jglobal,
                         cudaStream_t stream)
{
  const int ns = nslot_of(m);
  if (ns == 1 && integ == kIntegCvode) return launch_react_t<1, kIntegCvode>(m, o, a, cfg, counter, gstats, status, jglobal, stream);
  if (ns == 1 && integ == kIntegRK64) return launch_react_t<1, kIntegRK64>(m, o, a, cfg, counter, gstats, status, jglobal, stream);
  if (ns == 2 && integ == kIntegCvode) return launch_react_t<2, kIntegCvode>(m, o, a, cfg, counter, gstats, status, jglobal, stream);
  if (ns == 2 && integ == kIntegRK64) return launch_react_t<2, kIntegRK64>(m, o, a, cfg, counter, gstats, status, jglobal, stream);
  return cudaErrorInvalidValue;
}

int react_occupancy(const DevMech& m, int integ, int threads, size_t smem)
{
  int nb = 0;
  const int ns = nslot_of(m);
  cudaError_t e = cudaErrorInvalidValue;
#define PLM_OCC(NS, IG)                                                                                         \
  {                                                                                                             \
    auto k = react_kernel<NS, IG>;                                                                              \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));                            \
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, threads, smem);                                   \
  }
  if (ns == 1 && integ == kIntegCvode) PLM_OCC(1, kIntegCvode)
  else if (ns == 1) PLM_OCC(1, kIntegRK64)
  else if (integ == kIntegCvode) PLM_OCC(2, kIntegCvode)
  else PLM_OCC(2, kIntegRK64)
#undef PLM_OCC
  return e == cudaSuccess ? nb : -1;
}

cudaError_t launch_rhs(const DevMech& m, int ncells, double t, const double* y, const double* rhoe,
                       const double* rhoesrc, const double* rYsrc, double* ydot, cudaStream_t stream)
{
  const int ld = ld_for(m);
  const int wpb = 4;
  const size_t smem = warp_mem_doubles(m.n, m.nsp_pad, m.nr_pad, 0, ld, true) * sizeof(double) * wpb;
  const int blocks = std::max(1, std::min((ncells + wpb - 1) / wpb, 148 * 8));
  if (nslot_of(m) == 1) {
    cudaFuncSetAttribute(rhs_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    rhs_kernel<1><<<blocks, wpb * 32, smem, stream>>>(m, ncells, t, y, rhoe, rhoesrc, rYsrc, ydot, ld);
  } else {
    cudaFuncSetAttribute(rhs_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    rhs_kernel<2><<<blocks, wpb * 32, smem, stream>>>(m, ncells, t, y, rhoe, rhoesrc, rYsrc, ydot, ld);
  }
  return cudaGetLastError();
}

cudaError_t launch_jac(const DevMech& m, int ncells, const double* y, double* J, cudaStream_t stream)
{
  const int ld = ld_for(m);
  const int wpb = 2;
  const size_t smem = warp_mem_doubles(m.n, m.nsp_pad, m.nr_pad, 0, ld, true) * sizeof(double) * wpb;
  const int blocks = std::max(1, std::min((ncells + wpb - 1) / wpb, 148 * 8));
  if (nslot_of(m) == 1) {
    cudaFuncSetAttribute(jac_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    jac_kernel<1><<<blocks, wpb * 32, smem, stream>>>(m, ncells, y, J, ld);
  } else {
    cudaFuncSetAttribute(jac_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    jac_kernel<2><<<blocks, wpb * 32, smem, stream>>>(m, ncells, y, J, ld);
  }
  return cudaGetLastError();
}

cudaError_t launch_box_costs(const double* FC, const int64_t* off, int nboxes, long long* costs, cudaStream_t stream)
{
  box_cost_kernel<<<nboxes, 256, 0, stream>>>(FC, off, nboxes, costs);
  return cudaGetLastError();
}

}  // namespace plm
