// TEST INFRASTRUCTURE -- one-lane host build of the lane = cell headers (pelelm_b200/csrc/lane_*.h and the generated
// gen/lane_<mech>.h), compiled with g++ by tests/test_lane_model.py. It lets the CPU suite check the generated
// straight-line mechanism code and the state-machine integrators against the oracle without a GPU. Never part of
// libplmchem.so: the product compiles the same headers with nvcc for sm_100a only.
#include <cstring>
#include <type_traits>
#include <vector>

#include "lane_chem.h"
#include "gen/lane_LiDryer.h"
#include "gen/lane_drm19.h"
#include "gen/lanetab_LiDryer.h"
#include "gen/lanetab_drm19.h"
#include "gen/lanetab_dodecane_lu.h"
#ifdef LANE_MODEL_DODECANE
#include "gen/lane_dodecane_lu.h"
#endif
#include "lane_cell.h"

using namespace plm::lane;

namespace {
struct Src {
  const double* s;
  double e;
  double operator()(int k) const { return s[k]; }
  double energy() const { return e; }
};
template <int N>
struct JDense {
  double* p;
  double& operator()(int r, int c) { return p[c * N + r]; }
};

template <class M>
void rhs1(int rt, double t, const double* y, double* ydot, double rhoe, double rhoesrc, const double* rys)
{
  double yy[M::N], f[M::N];
  for (int i = 0; i < M::N; ++i) yy[i] = y[i];
  Src s{rys, rhoesrc};
  if constexpr (M::kTable) {
    double sc[M::NS], g[M::NS];
    if (rt == 1) lt_rhs<M, 1>(LV{yy}, LV{f}, LV{sc}, LV{g}, rhoe + rhoesrc * t, s, false);
    else lt_rhs<M, 0>(LV{yy}, LV{f}, LV{sc}, LV{g}, rhoe + rhoesrc * t, s, false);
  } else {
    if (rt == 1) lane_rhs<M, 1>(yy, f, rhoe + rhoesrc * t, s, false);
    else lane_rhs<M, 0>(yy, f, rhoe + rhoesrc * t, s, false);
  }
  for (int i = 0; i < M::N; ++i) ydot[i] = f[i];
}
template <class M>
void jac1(const double* y, double* J)
{
  double yy[M::N];
  for (int i = 0; i < M::N; ++i) yy[i] = y[i];
  JDense<M::N> ja{J};
  for (int i = 0; i < M::N * M::N; ++i) J[i] = 0.0;
  if constexpr (M::kTable) {
    double sc[M::NS], g[M::NS], h[M::NS], wd[M::NS], jT[M::NS], cR[M::NS];
    lt_cjac<M>(LV{yy}, ja, LV{sc}, LV{g}, LV{h}, LV{wd}, LV{jT}, LV{cR});
  } else {
    lane_cjac<M>(yy, ja);
  }
}
}  // namespace

// one cell at a time through the round loop the kernel runs per lane (the warp votes degenerate to one lane: a setup is
// served in the round it is asked for)
template <class M, int EINT, class I, class LAY>
void react_cells(int is_bdf, const LaneOpts& o, int ncells, double* rY, double* rYsrc, double* T, double* rE, double* rEsrc,
                 double* FC, const int* mask, double dt, double time, int flat, int mks, int clean, double* IR, int* stats8)
{
  std::vector<double> fastbuf(size_t(LAY::FAST_BDF) + 8, 0.0), bigbuf(size_t(LAY::BIG_BDF) + 8, 0.0);
  Ws w{fastbuf.data(), bigbuf.data()};
  CellCfg g{time, dt, 1.0 / dt, flat, mks, clean};
  const int ns = M::NS, n = M::N;
  for (int c = 0; c < ncells; ++c) {
    if (is_bdf && !flat && mask && mask[c] == -1) {
      FC[c] = 0.0;
      continue;
    }
    CellView cv;
    if (flat) cv = CellView{rY + size_t(c) * n, 1, rYsrc + size_t(c) * ns, 1, nullptr, rE + c, rEsrc + c, FC + c, nullptr, 0};
    else cv = CellView{rY + c, ncells, rYsrc + c, ncells, T + c, rE + c, rEsrc + c, FC + c, IR ? IR + c : nullptr, ncells};
    I it;
    lane_cell_load<M, I, LAY>(it, w, cv, g, o);
    for (;;) {
      lane_round_rhs<M, EINT, I, LAY>(it, w);
      if constexpr (I::kBdf) {
        if (it.phase == PH_WAIT_SETUP) it.setup(w);
        if (it.phase == PH_SOLVE) it.solve(w);
        it.tail(w);
      }
      if (it.phase == PH_OUT) lane_cell_out<M, I, LAY>(it, w, cv, g);
      if (it.phase == PH_DONE) break;
    }
    LaneStats st;
    it.stats(st);
    lane_cell_store<M, I, LAY>(it, w, cv, g);
    if (stats8) {
      stats8[0] += st.nst; stats8[1] += st.nfe; stats8[2] += st.nje; stats8[3] += st.nsetups; stats8[4] += st.nni;
      stats8[5] += st.ncfn; stats8[6] += st.netf;
      if (st.flag < 0 && stats8[7] == 0) stats8[7] = st.flag;
    }
  }
}

template <class T>
struct Tag {
  using type = T;
};
template <class M>
int react_mech(int integ, int rt, int qmax, const LaneOpts& o, int ncells, double* rY, double* rYsrc, double* T, double* rE,
               double* rEsrc, double* FC, const int* mask, double dt, double time, int flat, int mks, int clean, double* IR,
               int* stats8)
{
  auto go = [&](auto ei, auto it, auto lay) {
    using I = typename decltype(it)::type;
    using LAY = typename decltype(lay)::type;
    react_cells<M, decltype(ei)::value, I, LAY>(integ == 0, o, ncells, rY, rYsrc, T, rE, rEsrc, FC, mask, dt, time, flat, mks,
                                                clean, IR, stats8);
  };
  using std::integral_constant;
  if (integ == 1) {
    if (rt == 1) go(integral_constant<int, 1>{}, Tag<LaneRK64<M, 1>>{}, Tag<LaneLay<M, 2>>{});
    else go(integral_constant<int, 0>{}, Tag<LaneRK64<M, 0>>{}, Tag<LaneLay<M, 2>>{});
  } else if (qmax <= 2) {
    if (rt == 1) go(integral_constant<int, 1>{}, Tag<LaneBdf<M, 2, 1>>{}, Tag<LaneLay<M, 2>>{});
    else go(integral_constant<int, 0>{}, Tag<LaneBdf<M, 2, 0>>{}, Tag<LaneLay<M, 2>>{});
  } else {
    if (rt == 1) go(integral_constant<int, 1>{}, Tag<LaneBdf<M, 5, 1>>{}, Tag<LaneLay<M, 5>>{});
    else go(integral_constant<int, 0>{}, Tag<LaneBdf<M, 5, 0>>{}, Tag<LaneLay<M, 5>>{});
  }
  return 0;
}

#ifdef LANE_MODEL_DODECANE
#define DISPATCH_DODECANE(mech, CALL) else if (!std::strcmp(mech, "dodecane_lu")) { using M = Mech_dodecane_lu; CALL; }
#else
#define DISPATCH_DODECANE(mech, CALL)
#endif
#define DISPATCH(mech, CALL)                                     \
  if (!std::strcmp(mech, "LiDryer")) { using M = Mech_LiDryer; CALL; }   \
  else if (!std::strcmp(mech, "drm19")) { using M = Mech_drm19; CALL; }  \
  else if (!std::strcmp(mech, "LiDryer_tab")) { using M = Tab_LiDryer; CALL; }   \
  else if (!std::strcmp(mech, "drm19_tab")) { using M = Tab_drm19; CALL; }  \
  else if (!std::strcmp(mech, "dodecane_lu_tab")) { using M = Tab_dodecane_lu; CALL; }  \
  DISPATCH_DODECANE(mech, CALL)                                  \
  else return -1;

extern "C" {
#ifdef LANE_MODEL_COUNT_SWAPS
void lane_model_swap_stats(long* out)
{
  out[0] = plm::lane::lane_model_swaps; out[1] = plm::lane::lane_model_pivots; out[2] = plm::lane::lane_model_lus;
  out[3] = plm::lane::lane_model_lus_swapped;
}
#endif
int lane_model_rhs(const char* mech, int rt, double t, const double* y, double* ydot, double rhoe, double rhoesrc,
                   const double* rys)
{
  DISPATCH(mech, rhs1<M>(rt, t, y, ydot, rhoe, rhoesrc, rys));
  return 0;
}
// arrays as oracle/port/react.c orc_react_soa takes them (component-major [ns][ncells]) or, flat != 0, the 1dArray layout
int lane_model_react(const char* mech, int integ /*0 cvode 1 rk64*/, int reactor_type, double rtol, double atol, int qmax,
                     int maxcor, const double* typ_vals, int rk_guess, int rk_min, int rk_max, int ncells, double* rY,
                     double* rYsrc, double* T, double* rE, double* rEsrc, double* FC, const int* mask, double dt, double time,
                     int flat, int mks, int clean, double* IR, int* stats8)
{
  LaneOpts o{};
  o.rtol = rtol; o.atol = atol; o.qmax = qmax; o.maxcor = maxcor;
  o.use_typ = typ_vals && typ_vals[0] > 0.0;
  o.typ_vals = typ_vals;
  o.mxstep = 10000; o.maxnef = 100; o.msbj = 100;
  o.rk64_guess = rk_guess; o.rk64_min = rk_min; o.rk64_max = rk_max;
  DISPATCH(mech, (react_mech<M>(integ, reactor_type, qmax, o, ncells, rY, rYsrc, T, rE, rEsrc, FC, mask, dt, time, flat, mks,
                                integ == 0 ? clean : 0, IR, stats8)));
  return 0;
}
int lane_model_jac(const char* mech, const double* y, double* J)
{
  DISPATCH(mech, jac1<M>(y, J));
  return 0;
}
}
