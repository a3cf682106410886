// One translation unit per (mechanism, kernel family) of the lane = cell path, so that the straight-line kernels compile in
// parallel: nvcc ... -DLANE_MECH=<LiDryer|drm19> -DLANE_KIND=<0|2|5> lane_inst.cu
//   LANE_KIND 0: ReactorRK64 (both reactor types) + the fKernelSpec / cJac parity kernels
//   LANE_KIND 2: ReactorCvode, cvode.max_order <= 2      LANE_KIND 5: ReactorCvode, cvode.max_order <= 5
#include <algorithm>

#include "lane_kernels.cuh"

#define PLM_CAT2(a, b) a##b
#define PLM_CAT(a, b) PLM_CAT2(a, b)
#define PLM_STR2(x) #x
#define PLM_STR(x) PLM_STR2(x)
// LANE_TAB = 1: the mechanism as loops over constant-memory tables (lane_tab.h, gen/lanetab_<mech>.h);
// LANE_TAB = 0: as straight-line generated code (lane_chem.h, gen/lane_<mech>.h)
#if LANE_TAB
#include PLM_STR(PLM_CAT(gen/lanetab_, LANE_MECH).h)
#else
#include PLM_STR(PLM_CAT(gen/lane_, LANE_MECH).h)
#endif

namespace plm {
using namespace lane;
using namespace lanek;
#if LANE_TAB
using M = PLM_CAT(Tab_, LANE_MECH);
#else
using M = PLM_CAT(Mech_, LANE_MECH);
#endif

template <class I, class LAY, int EINT>
static cudaError_t launch_kernel(const ReactOpts& o, const BoxArgs& a, int blocks, unsigned long long* counter,
                                 unsigned long long* gstats, int* status, double* ws, int theta, cudaStream_t stream)
{
  using SH = LaneShape<I, LAY>;
  auto kern = lane_react_kernel<M, I, LAY, EINT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SH::SMEM);
  if (e != cudaSuccess) return e;
  kern<<<blocks, SH::THREADS, SH::SMEM, stream>>>(o, a, counter, gstats, status, ws, theta);
  return cudaGetLastError();
}

#if LANE_KIND == 0
cudaError_t PLM_CAT(lane_launch_rk64_, LANE_MECH)(const ReactOpts& o, const BoxArgs& a, int eint, int blocks,
                                                  unsigned long long* counter, unsigned long long* gstats, int* status,
                                                  double* ws, int theta, cudaStream_t stream)
{
  if (eint) return launch_kernel<LaneRK64<M, 1>, LaneLay<M, 2>, 1>(o, a, blocks, counter, gstats, status, ws, theta, stream);
  return launch_kernel<LaneRK64<M, 0>, LaneLay<M, 2>, 0>(o, a, blocks, counter, gstats, status, ws, theta, stream);
}
cudaError_t PLM_CAT(lane_launch_rhs_, LANE_MECH)(int eint, int ncells, double t, const double* y, const double* rhoe,
                                                 const double* rhoesrc, const double* rYsrc, double* ydot, cudaStream_t stream)
{
  const int blocks = std::max(1, std::min((ncells + kLaneThreads - 1) / kLaneThreads, 148 * 2));
  if (eint) lane_rhs_kernel<M, 1><<<blocks, kLaneThreads, 0, stream>>>(ncells, t, y, rhoe, rhoesrc, rYsrc, ydot);
  else lane_rhs_kernel<M, 0><<<blocks, kLaneThreads, 0, stream>>>(ncells, t, y, rhoe, rhoesrc, rYsrc, ydot);
  return cudaGetLastError();
}
cudaError_t PLM_CAT(lane_launch_jac_, LANE_MECH)(int ncells, const double* y, double* J, cudaStream_t stream)
{
  const int blocks = std::max(1, std::min((ncells + kLaneThreads - 1) / kLaneThreads, 148 * 2));
  lane_jac_kernel<M><<<blocks, kLaneThreads, 0, stream>>>(ncells, y, J);
  return cudaGetLastError();
}
size_t PLM_CAT(lane_ws_rk64_, LANE_MECH)() { return LaneLay<M, 2>::BIG_RK; }
int PLM_CAT(lane_threads_rk64_, LANE_MECH)() { return LaneShape<LaneRK64<M, 0>, LaneLay<M, 2>>::THREADS; }
#else
#define PLM_BDF_NAME PLM_CAT(PLM_CAT(PLM_CAT(lane_launch_bdf, LANE_KIND), _), LANE_MECH)
#define PLM_WS_NAME PLM_CAT(PLM_CAT(PLM_CAT(lane_ws_bdf, LANE_KIND), _), LANE_MECH)
cudaError_t PLM_BDF_NAME(const ReactOpts& o, const BoxArgs& a, int blocks, unsigned long long* counter,
                         unsigned long long* gstats, int* status, double* ws, int theta, cudaStream_t stream)
{
  return launch_kernel<LaneBdf<M, LANE_KIND, 0>, LaneLay<M, LANE_KIND>, 0>(o, a, blocks, counter, gstats, status, ws, theta, stream);
}
size_t PLM_WS_NAME() { return LaneLay<M, LANE_KIND>::BIG_BDF; }
#define PLM_TH_NAME PLM_CAT(PLM_CAT(PLM_CAT(lane_threads_bdf, LANE_KIND), _), LANE_MECH)
int PLM_TH_NAME() { return LaneShape<LaneBdf<M, LANE_KIND, 0>, LaneLay<M, LANE_KIND>>::THREADS; }
#endif

}  // namespace plm
