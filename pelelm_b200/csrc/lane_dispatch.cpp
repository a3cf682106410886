// Host-side dispatch of the lane = cell path (lane_kernels.cuh, lane_inst.cu): which requests it covers, its workspace,
// and the launch.  Used for the mechanisms whose per-cell state is small enough to keep 32 cells per warp in flight
// (LiDryer, drm19) with ReactorRK64 and with ReactorCvode + analytic Jacobian on the enthalpy reactor -- the
// configuration PeleLM runs; every other flavour (dodecane_lu, difference-quotient Jacobian, internal-energy BDF) goes
// through the warp-per-cell kernels (kernels.cu).
#include <cuda_runtime.h>

#include <cstdlib>

#include "kernels.h"

namespace plm {

#define PLM_LANE_DECL(MECH)                                                                                               \
  cudaError_t lane_launch_rk64_##MECH(const ReactOpts&, const BoxArgs&, int, int, unsigned long long*, unsigned long long*, \
                                      int*, double*, int, cudaStream_t);                                                  \
  cudaError_t lane_launch_rhs_##MECH(int, int, double, const double*, const double*, const double*, const double*, double*, \
                                     cudaStream_t);                                                                       \
  cudaError_t lane_launch_jac_##MECH(int, const double*, double*, cudaStream_t);                                          \
  cudaError_t lane_launch_bdf2_##MECH(const ReactOpts&, const BoxArgs&, int, unsigned long long*, unsigned long long*, int*, \
                                      double*, int, cudaStream_t);                                                        \
  cudaError_t lane_launch_bdf5_##MECH(const ReactOpts&, const BoxArgs&, int, unsigned long long*, unsigned long long*, int*, \
                                      double*, int, cudaStream_t);                                                        \
  size_t lane_ws_rk64_##MECH();                                                                                           \
  size_t lane_ws_bdf2_##MECH();                                                                                           \
  size_t lane_ws_bdf5_##MECH();                                                                                           \
  int lane_threads_rk64_##MECH();                                                                                         \
  int lane_threads_bdf2_##MECH();                                                                                         \
  int lane_threads_bdf5_##MECH();
PLM_LANE_DECL(LiDryer)
PLM_LANE_DECL(drm19)
#undef PLM_LANE_DECL

namespace {
bool lane_enabled()
{
  static const bool v = [] {
    const char* e = std::getenv("PLM_LANE");
    return !(e != nullptr && e[0] == '0');
  }();
  return v;
}
// lane-rounds of waiting that trigger a linear-solver setup for the lanes of a warp that want one (lane_kernels.cuh (C))
int lane_theta()
{
  static const int v = [] {
    const char* e = std::getenv("PLM_LANE_THETA");
    const int t = e ? std::atoi(e) : 0;
    const char* al = std::getenv("PLM_LANE_ALIGN");  // development: phase-aligned rounds (one CTA barrier per round)
    const char* co = std::getenv("PLM_LANE_COOP");   // development: warp-cooperative factorisations (lane_kernels.cuh)
    // phase-aligned rounds are the default: LiDryer ReactorCvode 295 k -> 636 k cells*dt/s, drm19 62 k -> 165 k (profiles/r02d, r02e)
    return (t > 0 ? t : 64) | ((al != nullptr && al[0] == '0') ? 0 : (1 << 30)) | ((co != nullptr && co[0] == '1') ? (1 << 29) : 0);
  }();
  return v;
}
int mech_id(const DevMech& m)
{
  if (m.ns == 9 && m.nr == 21) return 0;
  if (m.ns == 21 && m.nr == 84) return 1;
  return -1;
}
}  // namespace

bool lane_supported(const DevMech& m, int integ, const ReactOpts& o)
{
  if (!lane_enabled() || mech_id(m) < 0) return false;
  if (integ == kIntegRK64) return true;
  return o.ajac == 1 && m.eint == 0;
}

// What the automatic choice (plm_reactor_set_path 0) runs through the lane = cell kernels: the requests where they are the
// faster family on B200 (profiles/r02c_lane_ab.json: LiDryer ReactorRK64 5.6x, ReactorCvode 1.3x the warp-per-cell kernels;
// drm19's straight-line code is instruction-fetch bound and stays with the warp-per-cell kernels).  PLM_LANE_AUTO=all|none
// overrides (development).
bool lane_auto(const DevMech& m, int integ, const ReactOpts& o)
{
  if (!lane_supported(m, integ, o)) return false;
  static const int mode = [] {
    const char* e = std::getenv("PLM_LANE_AUTO");
    if (e != nullptr && e[0] == 'a') return 2;
    if (e != nullptr && e[0] == 'n') return 0;
    return 1;
  }();
  if (mode == 2) return true;
  if (mode == 0) return false;
  return mech_id(m) == 0;
}

int lane_threads(const DevMech& m, int integ, int qmax)
{
  const int id = mech_id(m);
  if (integ == kIntegRK64) return id == 0 ? lane_threads_rk64_LiDryer() : lane_threads_rk64_drm19();
  if (qmax <= 2) return id == 0 ? lane_threads_bdf2_LiDryer() : lane_threads_bdf2_drm19();
  return id == 0 ? lane_threads_bdf5_LiDryer() : lane_threads_bdf5_drm19();
}

size_t lane_ws_bytes(const DevMech& m, int integ, int qmax, int blocks)
{
  const int id = mech_id(m);
  size_t per_cell = 0;
  if (integ == kIntegRK64) per_cell = id == 0 ? lane_ws_rk64_LiDryer() : lane_ws_rk64_drm19();
  else if (qmax <= 2) per_cell = id == 0 ? lane_ws_bdf2_LiDryer() : lane_ws_bdf2_drm19();
  else per_cell = id == 0 ? lane_ws_bdf5_LiDryer() : lane_ws_bdf5_drm19();
  return per_cell * 32 * size_t(blocks) * (lane_threads(m, integ, qmax) / 32) * sizeof(double);
}

cudaError_t launch_lane_react(const DevMech& m, const ReactOpts& o, const BoxArgs& a, int integ, int blocks,
                              unsigned long long* counter, unsigned long long* gstats, int* status, double* ws,
                              cudaStream_t stream)
{
  if (!lane_supported(m, integ, o)) return cudaErrorNotSupported;
  const int id = mech_id(m), th = lane_theta();
  if (integ == kIntegRK64)
    return id == 0 ? lane_launch_rk64_LiDryer(o, a, m.eint, blocks, counter, gstats, status, ws, th, stream)
                   : lane_launch_rk64_drm19(o, a, m.eint, blocks, counter, gstats, status, ws, th, stream);
  if (o.qmax <= 2)
    return id == 0 ? lane_launch_bdf2_LiDryer(o, a, blocks, counter, gstats, status, ws, th, stream)
                   : lane_launch_bdf2_drm19(o, a, blocks, counter, gstats, status, ws, th, stream);
  return id == 0 ? lane_launch_bdf5_LiDryer(o, a, blocks, counter, gstats, status, ws, th, stream)
                 : lane_launch_bdf5_drm19(o, a, blocks, counter, gstats, status, ws, th, stream);
}

cudaError_t launch_lane_rhs(const DevMech& m, int ncells, double t, const double* y, const double* rhoe, const double* rhoesrc,
                            const double* rYsrc, double* ydot, cudaStream_t stream)
{
  const int id = mech_id(m);
  if (id < 0) return cudaErrorNotSupported;
  return id == 0 ? lane_launch_rhs_LiDryer(m.eint, ncells, t, y, rhoe, rhoesrc, rYsrc, ydot, stream)
                 : lane_launch_rhs_drm19(m.eint, ncells, t, y, rhoe, rhoesrc, rYsrc, ydot, stream);
}

cudaError_t launch_lane_jac(const DevMech& m, int ncells, const double* y, double* J, cudaStream_t stream)
{
  const int id = mech_id(m);
  if (id < 0) return cudaErrorNotSupported;
  return id == 0 ? lane_launch_jac_LiDryer(ncells, y, J, stream) : lane_launch_jac_drm19(ncells, y, J, stream);
}

}  // namespace plm
