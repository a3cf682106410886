/* TEST INFRASTRUCTURE -- per-cell restatement of ReactorRK64::react (PelePhysics/Reactions/ReactorRK64.cpp:206-317,
 * coefficients ReactorRK64.H:11-42): 6-stage low-storage RK with embedded error, absolute-error step
 * controller that never rejects a step. Pinned against oracle/_ref in tests/test_oracle_port_vs_ref.py. */
#include "oracle.h"

#include <math.h>

static const double alpha_rk64[6] = {0.218150805229859, 0.256702469801519, 0.527402592007520,
                                     0.0484864267224467, 1.24517071533530, 0.412366034843237};
static const double beta_rk64[6] = {-0.113554138044166, -0.215118587818400, -0.0510152146250577,
                                    -1.07992686223881, -0.248664241213447, 0.0};
static const double err_rk64[6] = {-0.0554699315064507, 0.158481845574980, -0.0905918835751907,
                                   -0.219084567203338, 0.164022338959433, 0.0426421977505659};

int orc_rk64_cell(const plm_mech_tables* m, const orc_opts* o, double* soln, double rhoe_init, double rhoesrc_ext,
                  const double* rYsrc, double time_init, double dt_react)
{
  return orc_rk64_cell_ex(m, o, soln, rhoe_init, rhoesrc_ext, rYsrc, time_init, dt_react, 1);
}

/* clamp_inside = 1: react(box) (ReactorRK64.cpp:206-317) limits every new step to the time left (:307, inside the loop);
 * clamp_inside = 0: the 1dArray react (ReactorRK64.cpp:24-165) applies that limit only AFTER its loop (:141-142), so its last
 * step runs past time_out -- a quirk of the reference the drop-in reproduces. */
int orc_rk64_cell_ex(const plm_mech_tables* m, const orc_opts* o, double* soln, double rhoe_init, double rhoesrc_ext,
                     const double* rYsrc, double time_init, double dt_react, int clamp_inside)
{
  const int neq = m->ns + 1;
  const double exp1 = 0.25, exp2 = 0.2, betaerr = 1.0, tinyval = 1e-50;
  double carry[ORC_MAXN], err[ORC_MAXN], rhs[ORC_MAXN];
  const double time_out = time_init + dt_react;
  double current_time = time_init;
  for (int i = 0; i < neq; i++) carry[i] = soln[i];
  double dt_rk = dt_react / (double)o->rk64_guess;
  const double dt_rk_min = dt_react / (double)o->rk64_max;
  const double dt_rk_max = dt_react / (double)o->rk64_min;
  int nsteps = 0;
  while (current_time < time_out) {
    for (int i = 0; i < neq; i++) err[i] = 0.0;
    for (int stage = 0; stage < 6; stage++) {
      orc_rhs_rt(m, o->reactor_type, current_time - time_init, soln, rhs, rhoe_init, rhoesrc_ext, rYsrc);
      for (int i = 0; i < neq; i++) {
        err[i] += err_rk64[stage] * dt_rk * rhs[i];
        soln[i] = carry[i] + alpha_rk64[stage] * dt_rk * rhs[i];
        carry[i] = soln[i] + beta_rk64[stage] * dt_rk * rhs[i];
      }
    }
    current_time += dt_rk;
    nsteps++;
    double max_err = tinyval;
    for (int i = 0; i < neq; i++) max_err = fabs(err[i]) > max_err ? fabs(err[i]) : max_err;
    if (max_err < o->atol) {
      const double cf = betaerr * pow((o->atol / max_err), exp1);
      dt_rk = (dt_rk_max < dt_rk * cf) ? dt_rk_max : dt_rk * cf;
    } else {
      const double cf = betaerr * pow((o->atol / max_err), exp2);
      dt_rk = (dt_rk_min > dt_rk * cf) ? dt_rk_min : dt_rk * cf;
    }
    if (clamp_inside) dt_rk = (dt_rk < time_out - current_time) ? dt_rk : time_out - current_time;
  }
  return nsteps;
}
