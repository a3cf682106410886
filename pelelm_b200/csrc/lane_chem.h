// Lane = cell right-hand side and Jacobian: the hand-written part around the generated mechanism kernels
// (gen/lane_<mech>.h).  One thread = one cell.
//
//   lane_get_T    GET_T_GIVEN_HY / GET_T_GIVEN_EY   Support/Mechanism/Models/drm19/mechanism.H:1967-2018, :1912-1963
//   lane_rhs      utils::fKernelSpec                PelePhysics/Reactions/ReactorUtils.H:274-332
//                 (RHY2T / REY2T, RTY2Hi / RTY2Ei, RTY2Cp / RTY2Cv, RTY2WDOT = CKYTCR + CKWC; Fuego.H)
//   lane_cjac     cvode::cJac                       PelePhysics/Reactions/ReactorCvodeJacobian.cpp:84-153
#pragma once
#include "lane_common.h"

namespace plm {
namespace lane {

// The reference evaluates the enthalpy at tmin and tmax on every call to guard the range; h(T) is monotone, so the two
// sweeps only matter when the guess or the iterate leaves [tmin, tmax] -- that path (verbatim) lives here, out of line.
template <class M, int EINT>
PLM_LNOINL double lane_get_T_guarded(double hin, const double (&Y)[M::NS], double Tguess, double& c_out)
{
  const double tmin = 90, tmax = 4000, tol = 1.e-6;
  double hmin, cmin, hmax, cmax;
  M::template mix_hc<EINT>(tmin, Y, hmin, cmin);
  M::template mix_hc<EINT>(tmax, Y, hmax, cmax);
  double t1, h1, c = cmin;
  if (hin < hmin || hin > hmax) {
    t1 = (hin < hmin) ? tmin - (hmin - hin) / cmin : tmax - (hmax - hin) / cmax;
    M::template mix_hc<EINT>(t1, Y, h1, c_out);
    return t1;
  }
  t1 = Tguess;
  if (!(t1 >= tmin && t1 <= tmax)) t1 = tmin + (tmax - tmin) / (hmax - hmin) * (hin - hmin);
  bool fresh = false;
  for (int i = 0; i < 200; ++i) {
    M::template mix_hc<EINT>(t1, Y, h1, c);
    fresh = true;
    double dt = (hin - h1) / c;
    if (dt > 100.) {
      dt = 100.;
    } else if (dt < -100.) {
      dt = -100.;
    } else if (lane_fabs(dt) < tol) {
      break;
    } else if (t1 + dt == t1) {
      break;
    }
    t1 += dt;
    fresh = false;
    if (t1 != t1) break;  // NaN: the reference iterates 200 times on it and returns NaN; same result, no work
  }
  if (!fresh) M::template mix_hc<EINT>(t1, Y, h1, c);
  c_out = c;
  return t1;
}

// Newton iteration on T (tol 1e-6 K, <= 200 iterations, +-100 K clamp; the step that passes the |dt| < tol test is NOT
// applied).  c_out = cp (or cv) of the mixture at the returned T.
template <class M, int EINT>
PLM_LHD double lane_get_T(double hin, const double (&Y)[M::NS], double Tguess, double& c_out)
{
  const double tmin = 90, tmax = 4000, tol = 1.e-6;
  double t1 = Tguess, h1, c = 0.0;
  bool ok = (t1 >= tmin && t1 <= tmax);  // false for a NaN guess
  if (ok) {
    bool fresh = false;
    for (int i = 0; i < 200; ++i) {
      M::template mix_hc<EINT>(t1, Y, h1, c);
      fresh = true;
      double dt = lane_div(hin - h1, c);
      if (dt > 100.) {
        dt = 100.;
      } else if (dt < -100.) {
        dt = -100.;
      } else if (lane_fabs(dt) < tol) {
        break;
      } else if (t1 + dt == t1) {
        break;
      }
      t1 += dt;
      fresh = false;
      if (t1 != t1) break;  // NaN: the reference iterates 200 times on it and returns NaN; same result, no work
    }
    ok = (t1 >= tmin && t1 <= tmax);
    if (ok && !fresh) M::template mix_hc<EINT>(t1, Y, h1, c);
  }
  if (!ok) return lane_get_T_guarded<M, EINT>(hin, Y, Tguess, c_out);
  c_out = c;
  return t1;
}

// fKernelSpec for one cell: y = [rhoY_k, T guess] -> ydot; rys(k) = external species source; returns the temperature
// found (the integrators' load / store phases use it as box_flatten's / box_unflatten's RHY2T).
// only_T: stop after the temperature solve (ydot untouched).
template <class M, int EINT, class SRC>
PLM_LHD double lane_rhs(const double (&y)[M::N], double (&ydot)[M::N], double hin_rho, const SRC& rys, bool only_T)
{
  constexpr int ns = M::NS;
  double massfrac[ns];
  double rho = 0.0;
#pragma unroll
  for (int k = 0; k < ns; ++k) {
    massfrac[k] = y[k];
    rho += massfrac[k];
  }
  const double rho_inv = lane_rcp(rho);
#pragma unroll
  for (int k = 0; k < ns; ++k) massfrac[k] *= rho_inv;
  const double nrg = hin_rho * rho_inv;  // (rhoe_init + rhoesrc_ext * dt_save) / rho
  double c;
  const double T = lane_get_T<M, EINT>(nrg, massfrac, y[ns], c);
  if (only_T) return T;
  if (T != T) {  // a NaN state (the explicit integrator blew up): every component of the reference's result is NaN as well;
                 // skip the rates instead of dragging the other 31 cells of the warp through the library's NaN paths
#pragma unroll
    for (int k = 0; k <= ns; ++k) ydot[k] = T;
    return T;
  }
  double sc[ns], wdot[ns];
#pragma unroll
  for (int k = 0; k < ns; ++k) sc[k] = (rho * massfrac[k] * M::imw_of(k)) * 1.0e6;  // CKYTCR :2403, CKWC :4911
  M::production_rate(sc, T, wdot);
  double ei[ns];
  M::template species_h<EINT>(T, ei);
  double rhoesrc = rys.energy();
#pragma unroll
  for (int k = 0; k < ns; ++k) {
    const double cdot_rYs = (wdot[k] * 1.0e-6) * M::mw_of(k) + rys(k);
    ydot[k] = cdot_rYs;
    rhoesrc -= cdot_rYs * ei[k];
  }
  ydot[ns] = rhoesrc * lane_div(rho_inv, c);
  return T;
}

// cJac for one cell: J(row, col) <- d ydot / d y, y = [rhoY_k, T]
template <class M, class JA>
PLM_LHD void lane_cjac(const double (&y)[M::N], JA& J)
{
  constexpr int ns = M::NS;
  double rho = 0.0;
#pragma unroll
  for (int k = 0; k < ns; ++k) rho = rho + y[k];
  double sc[ns];
#pragma unroll
  for (int k = 0; k < ns; ++k) {
    const double massfrac = y[k] / rho;
    sc[k] = 1.e6 * (rho * massfrac * M::imw_of(k));
  }
  M::ajacobian(J, sc, y[ns]);
}

}  // namespace lane
}  // namespace plm
