// Host <-> kernel interface shared by kernels.cu and capi.cpp.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "../../include/plm_chem.h"
#include "dev_tables.h"

namespace plm {

constexpr int kReactThreads = 128;   // rhs / jac test kernels: 4 warps = 4 cells in flight per CTA (integration kernel: Lay<>::WARPS)
constexpr int kCellsPerGrab = 4;     // consecutive cells a warp takes from the queue at once
constexpr int kIntegCvode = 0, kIntegRK64 = 1;
constexpr int kMaxN = 64;

struct ReactOpts {
  double rtol, atol;
  int qmax, maxcor, ajac, use_typ, clean_init_massfrac;
  int rk64_guess, rk64_min, rk64_max;
  double typ_vals[kMaxN];
};

struct BoxArgs {
  int lo[3], hi[3];
  int64_t ncells;
  int flat;    // 1: AoS flat API (rY [c][n] incl. T, rYsrc [c][ns], rE/rEsrc/FC [c])
  plm_a4d rY, rYsrc, T, rE, rEsrc, FC;
  plm_a4i mask;
  double dt, time;
  // PeleLM::advance_chemistry fused mode (PeleLM.cpp:5736-5803): state and forcing are MKS; the kernel scales to CGS on
  // load (rhoY, F_rhoY x1e-3; rhoH, F_rhoH x10), back on store (x1e3, x0.1) and writes I_R = -(old - new)/dt - F
  int mks;
  plm_a4d IR;      // N components, p == nullptr: not wanted
  double dt_inv;   // 1/dt as the host computes it (PeleLM.cpp:5792)
  // several boxes in one launch (both kernel families; plm_react_boxes): nboxes > 0 => cell c of the call is cell
  // c - off[b] of tab[b], off[b] <= c < off[b + 1]; lo / hi / arrays / mask come from tab[b], everything else from this struct
  const BoxArgs* tab;
  const int64_t* off;
  int nboxes;
};

struct LaunchCfg {
  int blocks, threads;
  size_t smem_bytes;
};

// m: host copy of the descriptor (dimensions for dispatch); dm: its device copy (what the kernels read)
size_t react_smem_bytes(const DevMech& m, int integ, int qmax, int warps_per_block);
size_t react_jsave_doubles(const DevMech& m);
int react_threads(const DevMech& m, int integ, int qmax);  // CTA size of the integration kernel for this mechanism
int react_occupancy(const DevMech& m, int integ, int qmax, int threads, size_t smem);
cudaError_t launch_react(const DevMech& m, const DevMech* dm, const ReactOpts& o, const BoxArgs& a, const LaunchCfg& cfg,
                         int integ, unsigned long long* counter, unsigned long long* gstats, int* status,
                         double* jglobal, cudaStream_t stream);
cudaError_t launch_rhs(const DevMech& m, const DevMech* dm, int ncells, double t, const double* y, const double* rhoe,
                       const double* rhoesrc, const double* rYsrc, double* ydot, cudaStream_t stream);
cudaError_t launch_jac(const DevMech& m, const DevMech* dm, int ncells, const double* y, double* J, cudaStream_t stream);
// PeleLM_K.H:57-101 reactionRateRhoY over a box (MKS in, MKS out; mask < 0 => zero)
cudaError_t launch_reaction_rate(const DevMech& m, const DevMech* dm, const BoxArgs& a, plm_a4d rhoYdot, cudaStream_t stream);
// ---- lane = cell path (lane_kernels.cu): one thread per cell, straight-line generated mechanism code ----
bool lane_supported(const DevMech& m, int integ, const ReactOpts& o);   // the lane kernels cover the request
bool lane_auto(const DevMech& m, int integ, const ReactOpts& o);        // ... and are what the automatic choice runs
int lane_threads(const DevMech& m, int integ, int qmax);   // CTA size; one CTA per SM
size_t lane_ws_bytes(const DevMech& m, int integ, int qmax, int blocks);   // workspace of a launch with `blocks` CTAs
cudaError_t launch_lane_react(const DevMech& m, const ReactOpts& o, const BoxArgs& a, int integ, int blocks,
                              unsigned long long* counter, unsigned long long* gstats, int* status, double* ws,
                              cudaStream_t stream);
cudaError_t launch_lane_rhs(const DevMech& m, int ncells, double t, const double* y, const double* rhoe, const double* rhoesrc,
                            const double* rYsrc, double* ydot, cudaStream_t stream);
cudaError_t launch_lane_jac(const DevMech& m, int ncells, const double* y, double* J, cudaStream_t stream);
cudaError_t launch_dfma_peak(double* scratch, int blocks, int iters, cudaStream_t stream);
cudaError_t launch_box_costs(const double* FC, const int64_t* off, int nboxes, long long* costs, cudaStream_t stream);

}  // namespace plm
