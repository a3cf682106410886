/* TEST INFRASTRUCTURE -- interface of the plain-C CPU restatement (oracle "port").
 * Never included, linked or loaded by the product (pelelm_b200/csrc). */
#ifndef ORACLE_PORT_H
#define ORACLE_PORT_H
#include "gen/mech_tables.h"

#define ORC_MAXS 64  /* >= largest NUM_SPECIES (dodecane_lu: 53) */
#define ORC_MAXN 65
#define ORC_QMAX 5

#ifdef __cplusplus
extern "C" {
#endif

const plm_mech_tables* orc_mech(const char* name);

/* chem.c */
void orc_cp_R(const plm_mech_tables* m, double* s, const double* tc);
void orc_h_RT(const plm_mech_tables* m, double* s, const double* tc);
void orc_g_RT(const plm_mech_tables* m, double* s, const double* tc);
void orc_dcp_R(const plm_mech_tables* m, double* s, const double* tc);
double orc_cpbs(const plm_mech_tables* m, double T, const double* y);
double orc_hbms(const plm_mech_tables* m, double T, const double* y);
void orc_hms(const plm_mech_tables* m, double T, double* hms);
int orc_get_T_given_hY(const plm_mech_tables* m, double hin, const double* y, double* t);
int orc_get_T_given_eY(const plm_mech_tables* m, double ein, const double* y, double* t);
double orc_cvbs(const plm_mech_tables* m, double T, const double* y);
double orc_ubms(const plm_mech_tables* m, double T, const double* y);
void orc_ums(const plm_mech_tables* m, double T, double* ums);
void orc_production_rate(const plm_mech_tables* m, double* wdot, const double* sc, double T);
void orc_ckwc(const plm_mech_tables* m, double T, const double* C, double* wdot);
void orc_ajacobian(const plm_mech_tables* m, double* J, const double* sc, double T, int consP);
void orc_cjac(const plm_mech_tables* m, const double* y, double* Jout);
void orc_cjac_rt(const plm_mech_tables* m, int reactor_type, const double* y, double* Jout);
void orc_rhs_rt(const plm_mech_tables* m, int reactor_type, double dt_save, const double* y, double* ydot, double rhoe_init,
                double rhoesrc_ext, const double* rYs);
void orc_reaction_rate_rhoY(const plm_mech_tables* m, const double* rhoY, double rhoH, double T, int mask, double* rhoYdot);
void orc_rhs(const plm_mech_tables* m, double dt_save, const double* y, double* ydot, double rhoe_init,
             double rhoesrc_ext, const double* rYs);

/* options shared by the integrators (ReactorCvode.cpp:15-22, :722-726, :262-289; ReactorRK64.cpp:14-20) */
typedef struct orc_opts {
  double rtol, atol;
  double typ_vals[ORC_MAXN]; /* typ_vals[0] > 0 => atol_k = typ_k * atol (ReactorBase.cpp:105-123) */
  int max_order;             /* cvode.max_order (default 2) */
  int max_nls_iters;         /* ode.max_nls_iters (default 3) */
  int clean_init_massfrac;
  int rk64_guess, rk64_min, rk64_max;
  int analytical_jacobian;   /* 1: denseAJ_direct (cJac), 0: dense_direct (difference quotient) */
  int reactor_type;          /* ode.reactor_type: 1 = internal energy (UV), 2 = enthalpy (HP; PeleLM) */
} orc_opts;
void orc_default_opts(orc_opts* o);

typedef struct orc_stats {
  long nst, nfe, nje, nsetups, nni, ncfn, netf;
  int flag; /* 0 ok; <0: CVODE failure code of the reference (CV_TOO_MUCH_WORK -1, CV_ERR_FAILURE -3, CV_CONV_FAILURE -4, ...) */
} orc_stats;

/* bdf.c -- one cell, CVODE 6.2.0 BDF/Newton/dense-LU as configured by PelePhysics */
int orc_cvode_cell(const plm_mech_tables* m, const orc_opts* o, double* y /* n, in/out */, double rhoe_init,
                   double rhoesrc_ext, const double* rYsrc, double t0, double tout, orc_stats* st);
/* rk64.c -- one cell (ReactorRK64.cpp:206-317); returns nsteps */
int orc_rk64_cell(const plm_mech_tables* m, const orc_opts* o, double* y, double rhoe_init, double rhoesrc_ext,
                  const double* rYsrc, double t0, double dt);

int orc_rk64_cell_ex(const plm_mech_tables* m, const orc_opts* o, double* y, double rhoe_init, double rhoesrc_ext,
                     const double* rYsrc, double t0, double dt, int clamp_inside);

/* react.c -- the operator API (ReactorBase.H:26-55) on FAB-layout arrays */
typedef struct orc_a4 {
  double* p;
  long jstride, kstride, nstride;
  int lo[3], hi[3]; /* inclusive FAB bounds */
} orc_a4;
int orc_react_box(const plm_mech_tables* m, const orc_opts* o, int use_rk64, const int lo[3], const int hi[3],
                  orc_a4 rY, orc_a4 rYsrc, orc_a4 T, orc_a4 rE, orc_a4 rEsrc, orc_a4 FC, const int* mask,
                  long mjs, long mks, const int mlo[3], double dt, double time, orc_stats* sum);
int orc_react_flat(const plm_mech_tables* m, const orc_opts* o, int use_rk64, double* rY, double* rYsrc,
                   double* rX, double* rXsrc, double dt, double time, int ncells, double* FC, orc_stats* sum);
/* convenience for ctypes: contiguous component-major box of nx*ny*nz cells, lo = 0 */
int orc_react_soa(const char* mech, const orc_opts* o, int use_rk64, int ncells, double* rY, double* rYsrc,
                  double* T, double* rE, double* rEsrc, double* FC, const int* mask, double dt, double time,
                  int nthreads, orc_stats* sum);

/* knapsack.c -- AMReX_DistributionMapping.cpp:576-716, :846-870 */
void orc_knapsack(const long* wgts, int nboxes, int nprocs, double max_eff, int do_full, int nmax, int* proc_of_box,
                  double* efficiency);

#ifdef __cplusplus
}
#endif
#endif
