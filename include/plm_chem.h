/* plm_chem.h -- C ABI of libplmchem.so: the B200 (sm_100a) drop-in for PelePhysics' per-cell
 * chemistry operator.  Plain C types only; every entry point cites the reference interface it replaces.
 *
 * Pointers handed to the react/eval calls may be DEVICE pointers (GPU-built AMReX: all Array4 data is
 * device accessible, PeleLM.cpp:5752-5758) or HOST pointers (CPU-built caller): the library asks the
 * CUDA runtime which it is and stages host data through its own device buffers.  There is no CPU
 * compute path: every call fails with PLM_ERR_NO_DEVICE when no sm_100-class GPU is present.
 */
#ifndef PLM_CHEM_H
#define PLM_CHEM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* amrex::Array4<T> as a POD (amrex/Src/Base/AMReX_Array4.H:15-23):
 * element (i,j,k,n) = p[(i-lo[0]) + (j-lo[1])*jstride + (k-lo[2])*kstride + n*nstride] */
typedef struct { double* p; int64_t jstride, kstride, nstride; int lo[3], hi[3]; int ncomp; } plm_a4d;
typedef struct { int*    p; int64_t jstride, kstride, nstride; int lo[3], hi[3]; int ncomp; } plm_a4i;

typedef struct plm_reactor plm_reactor; /* opaque: one (mechanism, integrator) pair; ReactorBase object */

enum { PLM_OK = 0, PLM_ERR_ARG = -1, PLM_ERR_NO_DEVICE = -2, PLM_ERR_CUDA = -3, PLM_ERR_MECH = -4,
       PLM_ERR_INTEGRATION = -5 };

/* per-cell status written by the integrators (replaces cvodeErrHandler -> amrex::Abort,
 * ReactorCvodeUtils.cpp:9-21): 0 ok, else the CVODE flag of the failure (-1 too much work,
 * -2 too much accuracy, -3 error-test failures, -4 convergence failures, -22 bad ewt, -27 tout too close) */

/* counters summed over the cells of every react call since the previous plm_reactor_stats()
 * (ReactorCvode::print_final_stats, ReactorCvode.cpp:1929-1992) plus what the roofline needs */
typedef struct {
  int64_t ncells, nskipped;      /* integrated / masked (mask == -1) */
  int64_t nst, nfe, nje, nsetups, nni, ncfn, netf;
  int64_t nfailed;               /* cells with status != 0 */
  int64_t max_fc, min_fc;
  double  kernel_ms;             /* summed device time of the integration kernels (CUDA events per launch) */
} plm_stats;

/* ReactorBase::create(name) (Factory.H:73-80; PeleLM.cpp:701-709).
 * mechanism: "LiDryer" | "drm19" | "dodecane_lu"; integrator: "ReactorCvode" | "ReactorRK64". */
int  plm_reactor_create(plm_reactor** out, const char* mechanism, const char* integrator);

/* ReactorBase::init(reactor_type, ncells) + the ParmParse keys it reads
 * (ReactorCvode.cpp:9-31 ode.{rtol,atol,max_nls_iters,clean_init_massfrac}, :722-726 cvode.{solve_type,max_order};
 * ReactorRK64.cpp:14-20 ode.rk64_nsubsteps_{guess,min,max}).  reactor_type: 1 = internal energy (UV; REY2T / Ei / Cv, consP = 0 in cJac; the ReactEval default),
 * 2 = enthalpy (HP; PeleLM).
 * solve_type: "denseAJ_direct" (analytic Jacobian) | "dense_direct" (difference-quotient Jacobian). */
int  plm_reactor_init(plm_reactor*, int reactor_type, double rtol, double atol, int max_order, int max_nls_iters,
                      int clean_init_massfrac, const char* solve_type, int rk64_guess, int rk64_min, int rk64_max);

/* ReactorBase::set_typ_vals_ode (ReactorBase.cpp:29-56): N+1 values; typ[0] > 0 => atol_k = typ_k * atol */
int  plm_reactor_set_typ_vals(plm_reactor*, const double* typ, int n);

/* ReactorBase::react(box, rY, rYsrc, T, rEner, rEner_src, FC, mask, dt_react, time [,stream])
 * (ReactorBase.H:26-41; CPU semantics of ReactorCvode.cpp:1372-1453 / ReactorRK64.cpp:167-327).
 * dt_react/time are not modified (PeleLM is not built with MOD_REACTOR).  Returns <0 on failure,
 * else 0; per-cell failures are reported through plm_reactor_stats().nfailed and the status array.
 * Device pointers: the call only ENQUEUES work on cuda_stream and returns (the caller synchronises the stream, as
 * PeleLM does at PeleLM.cpp:5774); up to 16 calls on different streams may be in flight per reactor.
 * Host pointers: the library stages the arrays through device buffers and returns when the results are back. */
int  plm_react_box(plm_reactor*, const int lo[3], const int hi[3], plm_a4d rY, plm_a4d rYsrc, plm_a4d T,
                   plm_a4d rEner, plm_a4d rEner_src, plm_a4d FC, plm_a4i mask,
                   double* dt_react, double* time, void* cuda_stream);

/* ReactorBase::react(rY_in, rYsrc_in, rX_in, rX_src_in, dt_react, time, ncells [,stream])
 * (ReactorBase.H:43-55): AoS rY [ncells][N+1] incl. T, rYsrc [ncells][N], rX / rX_src [ncells]. */
int  plm_react_flat(plm_reactor*, double* rY, double* rYsrc, double* rX, double* rX_src,
                    double* dt_react, double* time, int ncells, void* cuda_stream);

/* The body of PeleLM::advance_chemistry's MFIter loop fused into the integration kernel (PeleLM.cpp:5736-5775) plus the
 * I_R update that follows it (PeleLM.cpp:5783-5800).  All arrays are MKS, as PeleLM holds them:
 *   load :  rhoY, frc_rhoY x 1e-3 ; rhoH, frc_rhoH x 10      (MKS -> CGS, :5737-5746; the forcing arrays are NOT modified:
 *                                                              the reference scales a private copy, FTemp)
 *   react:  exactly plm_react_box (time = 0, dt_react = dt)
 *   store:  rhoY x 1e3 ; rhoH x 0.1                           (CGS -> MKS, :5765-5772); T, FC as react() leaves them
 *   I_R  :  -(rhoY_old - rhoY_new) * (1/dt) - frc_rhoY        (:5795-5799), N components; I_R.p == NULL: skipped
 * Cells with mask == -1 get the same x1e-3 x1e3 / x10 x0.1 round trip the reference applies to them.
 * Saves the three extra passes over (2N+4) components of the reference sequence. Same pointer / stream rules as
 * plm_react_box. */
int  plm_advance_chemistry_box(plm_reactor*, const int lo[3], const int hi[3], plm_a4d rhoY, plm_a4d rhoH, plm_a4d T,
                               plm_a4d frc_rhoY, plm_a4d frc_rhoH, plm_a4d I_R, plm_a4d FC, plm_a4i mask, double dt,
                               void* cuda_stream);

/* Several boxes in ONE launch (device arrays only): react (mks = 0) or the fused advance_chemistry body (mks != 0, time = 0)
 * over every box of `boxes` -- the MFIter loop of PeleLM::advance_chemistry (PeleLM.cpp:5725-5780) for all local boxes of
 * a level at once.  The work queue then runs over the cells of all boxes, so its tail (the last, most expensive cells) is
 * paid once per level instead of once per box (both kernel families; the environment variable PLM_BOXES_PER_LAUNCH=1
 * keeps the warp-per-cell kernels on one launch per box, a development A/B).
 * The status array of plm_reactor_cell_status covers the boxes in the order given. */
typedef struct plm_box {
  int lo[3], hi[3];
  plm_a4d rY, rYsrc, T, rEner, rEner_src, FC, I_R;   /* I_R.p == NULL: not wanted (and ignored when mks == 0) */
  plm_a4i mask;                                       /* mask.p == NULL: every cell is integrated */
} plm_box;
int  plm_react_boxes(plm_reactor*, int nboxes, const plm_box* boxes, double dt, double time, int mks, void* cuda_stream);

/* reactionRateRhoY over a box (PeleLM_K.H:57-101, called by PeleLM::compute_instantaneous_reaction_rates,
 * PeleLM.cpp:2010-2070): MKS rhoY (N comps), rhoH, T (guess) -> rhoYdot = W_k * wdot_k in kg/m^3/s (N comps);
 * mask < 0 (EB covered) => 0.  T is not modified. */
int  plm_reaction_rate_rhoY(plm_reactor*, const int lo[3], const int hi[3], plm_a4d rhoY, plm_a4d rhoH, plm_a4d T,
                            plm_a4i mask, plm_a4d rhoYdot, void* cuda_stream);

/* Cells whose integration has failed so far (status != 0), over the life of the reactor: everything plm_reactor_stats has
 * collected plus the calls that have finished on their streams since.  Never blocks.  The reference aborts inside CVODE
 * (cvodeErrHandler, ReactorCvodeUtils.cpp:9-21); a caller that only enqueues work polls this after it has synchronised its
 * stream (PeleLM.cpp:5774) -- the adaptor does, before every react call, at close() and in print_final_stats. */
int  plm_reactor_failed_cells(plm_reactor*, long long* nfailed);
/* Which kernel family the react / eval calls of this reactor run (no reference counterpart; the parity tests use it to hold
 * both families to the same fixtures): 0 = automatic (default: lane = cell kernels -- one thread per cell, straight-line
 * generated mechanism code -- where they cover the request: LiDryer / drm19 with ReactorRK64 or ReactorCvode +
 * denseAJ_direct on the enthalpy reactor; warp-per-cell kernels otherwise), 1 = warp-per-cell kernels, 2 = lane = cell
 * kernels (react calls fail with PLM_ERR_ARG when they do not cover the request). */
int  plm_reactor_set_path(plm_reactor*, int path);
/* 1 when the react calls of this reactor, as initialised, run the lane = cell kernels, else 0 */
int  plm_reactor_uses_lane(const plm_reactor*);
/* waits for every outstanding call of this reactor, then returns and clears the accumulated counters */
int  plm_reactor_stats(plm_reactor*, plm_stats* out);
/* status word per cell of the last call, in box order (device or host destination) */
int  plm_reactor_cell_status(plm_reactor*, int* dst, int64_t ncells);
/* ReactorBase::close() (ReactorCvode.cpp:1908-1927; PeleLM.cpp:930) */
void plm_reactor_close(plm_reactor*);

int  plm_num_species(const plm_reactor*);
/* species k in mechanism order (CKSYMS_STR, mechanism.cpp): lets a binding assert it was compiled for the same mechanism */
const char* plm_species_name(const plm_reactor*, int k);
const char* plm_last_error(void);

/* ---- kernel-level entry points used by the parity tests (reference functions they mirror) ---- */
/* utils::fKernelSpec (ReactorUtils.H:274-332): y [ncells][N+1], ydot [ncells][N+1] */
int  plm_eval_rhs(plm_reactor*, int ncells, double t, const double* y, const double* rhoe_init,
                  const double* rhoesrc_ext, const double* rYsrc, double* ydot, void* cuda_stream);
/* cvode::cJac (ReactorCvodeJacobian.cpp:84-153): J [ncells][(N+1)^2] column-major */
int  plm_eval_jac(plm_reactor*, int ncells, const double* y, double* J, void* cuda_stream);

/* ---- cost model / load balance (PeleLM::getFuncCountDM PeleLM.cpp:5529-5623;
 *      DistributionMapping::KnapSackProcessorMap AMReX_DistributionMapping.cpp:576-870) ---- */
/* per-box cost = sum over the box of FC (device reduction); boxes given as consecutive cell ranges */
int  plm_box_costs(const double* FC, const int64_t* box_offsets, int nboxes, int64_t* costs, void* cuda_stream);
int  plm_knapsack(const int64_t* wgts, int nboxes, int nranks, double max_efficiency, int do_full_knapsack,
                  int nmax, int* rank_of_box, double* efficiency);
/* The same with the reference's default sort = true (AMReX_DistributionMapping.H:139-143): the i-th heaviest bucket (or, with
 * at most one box per rank, box i) goes to rank_order[i]; rank_order = plm_least_used_ranks of the kilobytes each rank holds
 * in FABs (DistributionMapping::LeastUsedCPUs :191-235, one all-gather of a long per rank; KnapSackDoIt :718-844,
 * RoundRobinDoIt :365-447). plm_knapsack is the rank_order = identity case (all ranks equally loaded). */
int  plm_knapsack_ordered(const int64_t* wgts, int nboxes, int nranks, double max_efficiency, int do_full_knapsack,
                          int nmax, const int* rank_order, int* rank_of_box, double* efficiency);
int  plm_least_used_ranks(const int64_t* kbytes, int nranks, int* rank_order);

/* ---- measurement: register-resident DFMA throughput of the current device in TFLOP/s (the FP64 SIMT roofline
 *      denominator; no reference counterpart) ---- */
int  plm_measure_fp64_peak(double* tflops, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
