// Lane = cell: one cell's way in and out of the caller's arrays, and one ROUND of the warp.
//
//   lane_cell_load     box_flatten                      PelePhysics/Reactions/ReactorUtils.H:76-136 (+ clean_init_massfrac)
//                      and the MKS -> CGS scaling of PeleLM::advance_chemistry (Source/PeleLM.cpp:5737-5746)
//   lane_cell_out      box_unflatten, first half        ReactorUtils.H:138-180 (species, rhoE/rhoH; T follows in a
//                                                       temperature-only round)
//   lane_cell_store    box_unflatten, second half       ReactorUtils.H:180-207, CGS -> MKS and I_R (PeleLM.cpp:5765-5799)
//   lane_round         one right-hand side per lane and what follows it (lane_integ.h)
#pragma once
#include "lane_integ.h"

namespace plm {
namespace lane {

// where one cell lives in the caller's arrays (element strides between components)
struct CellView {
  double* rY; long long sY;
  const double* rYsrc; long long sS;
  double* T;
  double* rE;
  const double* rEsrc;
  double* FC;
  double* IR; long long sI;  // nullptr: not wanted
};
struct CellCfg {
  double time, dt, dt_inv;
  int flat, mks, clean;
};

template <class M, class I, class LAY>
PLM_LHD void lane_cell_load(I& it, const Ws& w, const CellView& c, const CellCfg& g, const LaneOpts& o)
{
  constexpr int ns = M::NS;
  if (g.flat) {  // 1dArray API: [rhoY, T] per cell, T taken as given (ReactorCvode.cpp:1463-1763, CPU branch)
    for (int k = 0; k <= ns; ++k) w.d(I::YV + k) = c.rY[k];
    for (int k = 0; k < ns; ++k) w.g(LAY::SRC + k) = c.rYsrc[k];
    it.rhoe_init = *c.rE;
    it.rhoesrc = *c.rEsrc;
    it.begin(w, o, g.time, g.dt, true);
    return;
  }
  const bool mks = g.mks != 0;
  it.rhoe_init = mks ? lane_mul(*c.rE, 10.0) : *c.rE;
  it.rhoesrc = mks ? lane_mul(*c.rEsrc, 10.0) : *c.rEsrc;
  double rho_old = 0.0;
  for (int k = 0; k < ns; ++k) {
    double v = c.rY[k * c.sY], sv = c.rYsrc[k * c.sS];
    if (mks) {
      v = lane_mul(v, 1.0e-3);
      sv = lane_mul(sv, 1.0e-3);
    }
    w.d(I::YV + k) = v;
    w.g(LAY::SRC + k) = sv;
    rho_old += v;
  }
  if (g.clean) {
    for (int k = 0; k < ns; ++k) {
      const double v = w.d(I::YV + k);
      const double lo = 0.0 > v ? 0.0 : v;
      w.d(I::YV + k) = lo < rho_old ? lo : rho_old;
    }
  }
  w.d(I::YV + ns) = *c.T;
  it.begin(w, o, g.time, g.dt, false);
}

// integration over (phase PH_OUT): final species into Y, then the temperature-only round (box) or straight to the store
template <class M, class I, class LAY>
PLM_LHD void lane_cell_out(I& it, const Ws& w, const CellView& c, const CellCfg& g)
{
  constexpr int ns = M::NS;
  if (g.flat) {
    it.phase = PH_DONE;
    return;
  }
  if (g.clean) {  // ReactorUtils.H:160-166: the loaded and the clipped state, recomputed from the values still in the FAB
    double rho_old = 0.0, rho_oldclipped = 0.0;
    for (int k = 0; k < ns; ++k) {
      double r0 = c.rY[k * c.sY];
      if (g.mks) r0 = lane_mul(r0, 1.0e-3);
      rho_old += r0;
    }
    for (int k = 0; k < ns; ++k) {
      double r0 = c.rY[k * c.sY];
      if (g.mks) r0 = lane_mul(r0, 1.0e-3);
      const double lo = 0.0 > r0 ? 0.0 : r0;
      rho_oldclipped += (lo < rho_old ? lo : rho_old);
    }
    for (int k = 0; k < ns; ++k) {
      double r0 = c.rY[k * c.sY];
      if (g.mks) r0 = lane_mul(r0, 1.0e-3);
      const double lo = 0.0 > r0 ? 0.0 : r0;
      const double cl = lo < rho_old ? lo : rho_old;
      w.d(I::YV + k) = r0 + (w.d(I::YV + k) - cl) * rho_old / rho_oldclipped;
    }
  }
  it.hin_load = lane_add(it.rhoe_init, lane_mul(g.dt, it.rhoesrc));  // ReactorUtils.H:180, unfused as the host evaluates it
  it.phase = PH_FINISH;
}

template <class M, class I, class LAY>
PLM_LHD void lane_cell_store(I& it, const Ws& w, const CellView& c, const CellCfg& g)
{
  constexpr int ns = M::NS;
  const double rE_new = lane_add(it.rhoe_init, lane_mul(g.dt, it.rhoesrc));
  if (g.flat) {
    for (int k = 0; k <= ns; ++k) c.rY[k] = w.d(I::YV + k);
    *c.rE = rE_new;
  } else {
    for (int k = 0; k < ns; ++k) {
      double* dst = c.rY + k * c.sY;
      const double v = w.d(I::YV + k);
      if (g.mks) {  // CGS -> MKS (PeleLM.cpp:5765-5772), then I_R (PeleLM.cpp:5795-5799) from the old value still in place
        const double vnew = lane_mul(v, 1.0e3);
        if (c.IR) c.IR[k * c.sI] = lane_sub(lane_mul(-lane_sub(*dst, vnew), g.dt_inv), c.rYsrc[k * c.sS]);
        *dst = vnew;
      } else {
        *dst = v;
      }
    }
    *c.T = w.d(I::YV + ns);
    *c.rE = g.mks ? lane_mul(rE_new, 0.1) : rE_new;
  }
  *c.FC = double(it.fc());
  it.phase = PH_IDLE;
}

// (A) + (B): one right-hand side (or temperature solve) for this lane and its bookkeeping
template <class M, int EINT, class I, class LAY>
PLM_LHD void lane_round_rhs(I& it, const Ws& w)
{
  constexpr int N = M::N;
  if (it.rhs_active()) {
    double f[N];
    const int yo = it.rhs_vec();
    WsSrc src{w, LAY::SRC, it.rhoesrc};
    double T;
    if constexpr (M::kTable) {  // loops over the constant-memory tables; vectors stay in fast memory
      const LV fv{&w.d(I::FV)};
      T = lt_rhs<M, EINT>(LV{&w.d(yo)}, fv, LV{&w.d(I::SCV)}, LV{&w.d(I::GV)}, it.hin_rho(), src, it.only_T());
#pragma unroll
      for (int i = 0; i < N; ++i) f[i] = fv(i);
    } else {                    // straight-line generated code on register arrays
      double y[N];
#pragma unroll
      for (int i = 0; i < N; ++i) y[i] = w.d(yo + i);
      T = lane_rhs<M, EINT>(y, f, it.hin_rho(), src, it.only_T());
    }
    it.post_rhs(w, f, T);
  }
}

}  // namespace lane
}  // namespace plm
