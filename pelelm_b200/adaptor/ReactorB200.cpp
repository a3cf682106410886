// See ReactorB200.H. Everything here is argument marshalling: amrex::Array4 -> plm_a4d, ParmParse -> plm_reactor_init.
#include "ReactorB200.H"

#include "plm_chem.h"

#include <cstring>

namespace pele {
namespace physics {

// The two classes are header-only, so no translation unit would otherwise emit Register<T>::registered (the static
// whose initialiser inserts the creator into the factory table, Factory.H:96-130): instantiate them here.
template struct Factory<reactions::ReactorBase>::Register<reactions::ReactorCvodeB200>;
template struct Factory<reactions::ReactorBase>::Register<reactions::ReactorRK64B200>;

namespace reactions {

namespace {

const char* mechanism_name()
{
  // the reference fixes the mechanism at compile time (NUM_SPECIES / NUM_REACTIONS, mechanism.H)
  if (NUM_SPECIES == 9 && NUM_REACTIONS == 21) return "LiDryer";
  if (NUM_SPECIES == 21 && NUM_REACTIONS == 84) return "drm19";
  if (NUM_SPECIES == 53 && NUM_REACTIONS == 268) return "dodecane_lu";
  return nullptr;
}

template <class T, class A>
A to_a4(amrex::Array4<T> const& a)
{
  A r;
  r.p = a.p;
  r.jstride = a.jstride;
  r.kstride = a.kstride;
  r.nstride = a.nstride;
  r.lo[0] = a.begin.x; r.lo[1] = a.begin.y; r.lo[2] = a.begin.z;
  r.hi[0] = a.end.x - 1; r.hi[1] = a.end.y - 1; r.hi[2] = a.end.z - 1;  // Array4::end is exclusive
  r.ncomp = a.ncomp;
  return r;
}

void check(int rc, const char* where)
{
  if (rc != PLM_OK) amrex::Abort(std::string("ReactorB200: ") + where + ": " + plm_last_error());
}

} // namespace

int ReactorB200Common::b200_init(const char* integrator, int reactor_type, int verbose)
{
  m_b200_verbose = verbose;
  m_b200_reactor_type = reactor_type;
  m_is_cvode = std::strcmp(integrator, "ReactorCvode") == 0;
  const char* mech = mechanism_name();
  if (mech == nullptr) amrex::Abort("ReactorB200: libplmchem has no tables for this mechanism");
  check(plm_reactor_create(&m_handle, mech, integrator), "plm_reactor_create");
  // species order must be the compiled mechanism's
  amrex::Vector<std::string> names;
  CKSYMS_STR(names);
  for (int k = 0; k < NUM_SPECIES; ++k)
    if (names[k] != plm_species_name(m_handle, k)) amrex::Abort("ReactorB200: species order mismatch at " + names[k]);

  amrex::Real rtol = 1e-6, atol = 1e-10; // ReactorCvode.H:152-153 / ReactorRK64.H defaults
  int max_nls_iters = 3, clean = 0, max_order = 2, g = 10, mn = 5, mx = 500;
  std::string solve_type = "denseAJ_direct";
  amrex::ParmParse pp("ode");
  pp.query("rtol", rtol);
  pp.query("atol", atol);
  pp.query("max_nls_iters", max_nls_iters);
  pp.query("clean_init_massfrac", clean);
  m_b200_clean = clean;
  pp.query("rk64_nsubsteps_guess", g);
  pp.query("rk64_nsubsteps_min", mn);
  pp.query("rk64_nsubsteps_max", mx);
  amrex::ParmParse ppcv("cvode");
  ppcv.query("solve_type", solve_type);
  ppcv.query("max_order", max_order);
  if (m_is_cvode && solve_type != "denseAJ_direct" && solve_type != "dense_direct")
    amrex::Abort("ReactorCvodeB200: cvode.solve_type must be denseAJ_direct or dense_direct");
  if (verbose > 0)
    amrex::Print() << "Initializing " << integrator << "B200 (libplmchem, sm_100a): mechanism " << mech
                   << ", solve_type " << solve_type << ", max_order " << max_order << "\n";
  const int rc = plm_reactor_init(
    m_handle, reactor_type, rtol, atol, max_order, max_nls_iters, clean, solve_type.c_str(), g, mn, mx);
  check(rc, "plm_reactor_init");
  return 0;
}

// The reference aborts the run when CVODE fails in a cell (cvodeErrHandler, ReactorCvodeUtils.cpp:9-21). The batched kernel
// records the failure per cell instead; the count of failed cells of every call that has finished is polled here -- before
// the next call, at close() and in print_final_stats -- so that a GPU build, whose react() only enqueues, does not carry on
// with the partial state of a failed cell.
void ReactorB200Common::b200_abort_on_failed_cells()
{
  long long nfailed = 0;
  if (m_handle != nullptr && plm_reactor_failed_cells(m_handle, &nfailed) == PLM_OK && nfailed > 0)
    amrex::Abort("ReactorB200: integration failed in " + std::to_string(nfailed) + " cells (status per cell: plm_reactor_cell_status)");
}

void ReactorB200Common::b200_close()
{
  b200_abort_on_failed_cells();
  plm_reactor_close(m_handle);
  m_handle = nullptr;
}

int ReactorB200Common::b200_react_box(
  const amrex::Box& box,
  amrex::Array4<amrex::Real> const& rY_in,
  amrex::Array4<amrex::Real> const& rYsrc_in,
  amrex::Array4<amrex::Real> const& T_in,
  amrex::Array4<amrex::Real> const& rEner_in,
  amrex::Array4<amrex::Real> const& rEner_src_in,
  amrex::Array4<amrex::Real> const& FC_in,
  amrex::Array4<int> const& mask,
  amrex::Real& dt_react,
  amrex::Real& time,
  const amrex::Array<amrex::Real, NUM_SPECIES + 1>& typ_vals,
  void* stream)
{
  BL_PROFILE("Pele::ReactorB200::react()");
  b200_abort_on_failed_cells();  // of the calls before this one (a GPU build only enqueues; PeleLM syncs the stream per box)
  check(plm_reactor_set_typ_vals(m_handle, typ_vals.data(), NUM_SPECIES + 1), "plm_reactor_set_typ_vals");
  // Dim3 bounds: a 2-D build (Exec/RegTests/FlameSheet) has two-component IntVects and Array4s with k = 0 only
  const amrex::Dim3 blo = amrex::lbound(box), bhi = amrex::ubound(box);
  const int lo[3] = {blo.x, blo.y, blo.z};
  const int hi[3] = {bhi.x, bhi.y, bhi.z};
  const int rc = plm_react_box(
    m_handle, lo, hi, to_a4<amrex::Real, plm_a4d>(rY_in), to_a4<amrex::Real, plm_a4d>(rYsrc_in),
    to_a4<amrex::Real, plm_a4d>(T_in), to_a4<amrex::Real, plm_a4d>(rEner_in),
    to_a4<amrex::Real, plm_a4d>(rEner_src_in), to_a4<amrex::Real, plm_a4d>(FC_in), to_a4<int, plm_a4i>(mask),
    &dt_react, &time, stream);
  check(rc, "plm_react_box");
#ifdef MOD_REACTOR
  // only ReactEval is built this way
  if (m_is_cvode) {  // ReactorCvode.cpp:1444-1449
    dt_react = time - (time + dt_react);
    time += dt_react;
  } else {           // ReactorRK64.cpp:319-321: time = time_out, dt_react untouched
    time += dt_react;
  }
#endif
#ifndef AMREX_USE_GPU
  // CPU-built caller: results are already back on the host; report like the reference (nfe of the last cell ~ mean)
  plm_stats st;
  check(plm_reactor_stats(m_handle, &st), "plm_reactor_stats");
  if (st.nfailed > 0) amrex::Abort("ReactorB200: integration failed in " + std::to_string(st.nfailed) + " cells");
  return st.ncells > 0 ? static_cast<int>((m_is_cvode ? st.nfe : st.nst) / st.ncells) : 0;
#else
  return 0;
#endif
}

int ReactorB200Common::b200_react_flat(
  amrex::Real* rY_in, amrex::Real* rYsrc_in, amrex::Real* rX_in, amrex::Real* rX_src_in, amrex::Real& dt_react,
  amrex::Real& time, int ncells, const amrex::Array<amrex::Real, NUM_SPECIES + 1>& typ_vals, void* stream)
{
  BL_PROFILE("Pele::ReactorB200::react()");
  check(plm_reactor_set_typ_vals(m_handle, typ_vals.data(), NUM_SPECIES + 1), "plm_reactor_set_typ_vals");
  check(plm_react_flat(m_handle, rY_in, rYsrc_in, rX_in, rX_src_in, &dt_react, &time, ncells, stream), "plm_react_flat");
  plm_stats st;
  check(plm_reactor_stats(m_handle, &st), "plm_reactor_stats");
  if (st.nfailed > 0) amrex::Abort("ReactorB200: integration failed in " + std::to_string(st.nfailed) + " cells");
#ifdef MOD_REACTOR
  if (m_is_cvode) {  // ReactorCvode.cpp:1734-1738
    dt_react = time - (time + dt_react);
    time += dt_react;
  } else {           // ReactorRK64.cpp:147-149
    time += dt_react;
  }
#endif
  return st.ncells > 0 ? static_cast<int>((m_is_cvode ? st.nfe : st.nst) / st.ncells) : 0;
}

void ReactorB200Common::b200_print_stats()
{
  b200_abort_on_failed_cells();
  plm_stats st;
  if (plm_reactor_stats(m_handle, &st) != PLM_OK) return;
  amrex::Print() << "-- Final Statistics (libplmchem) --\n"
                 << "cells = " << st.ncells << " skipped = " << st.nskipped << " failed = " << st.nfailed << "\n"
                 << "nSteps = " << st.nst << " nRHSeval = " << st.nfe << " nLinSetups = " << st.nsetups
                 << " nJacEval = " << st.nje << "\n"
                 << "nNonLinIt = " << st.nni << " nConvfail = " << st.ncfn << " nErrtests = " << st.netf
                 << " kernel_ms = " << st.kernel_ms << "\n";
}

} // namespace reactions
} // namespace physics
} // namespace pele
