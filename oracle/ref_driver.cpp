// TEST INFRASTRUCTURE -- not part of the product path.
//
// C entry points over the UNMODIFIED reference implementation (PelePhysics reactors + SUNDIALS
// CVODE 6.2.0 + AMReX 22.09), compiled from the sources where they lie under /root/reference by
// oracle/Makefile into oracle/_ref/libref_<mechanism>.so.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load that library.
//
// What is driven (reference file:line):
//   ReactorBase::create / init / react(box)   PelePhysics/Reactions/ReactorBase.H:15-98,
//                                             ReactorCvode.cpp:1372-1453, ReactorRK64.cpp:167-327
//   utils::fKernelSpec                        PelePhysics/Reactions/ReactorUtils.H:274-332
//   cvode::cJac                               PelePhysics/Reactions/ReactorCvodeJacobian.cpp:84-153
//   CKWC / GET_T_GIVEN_HY / CKHMS / CKCPBS    Support/Mechanism/Models/<mech>/mechanism.H
//   reactionRateRhoY                          Source/PeleLM_K.H:57-101 (the reference's own inline function)
//   VisMF::Write / VisMF::Read                amrex/Src/Base/AMReX_VisMF.cpp (the chkfile format ReactEval reads,
//                                             PelePhysics/Testing/Exec/ReactEval/main.cpp:358-371)
#include <AMReX.H>
#include <AMReX_ParmParse.H>
#include <AMReX_Box.H>
#include <AMReX_Array4.H>
#include <AMReX_Sundials.H>
#include <AMReX_MultiFab.H>
#include <AMReX_VisMF.H>

#include "mechanism.H"
#include "PelePhysics.H"
#include "ReactorBase.H"
#include "ReactorUtils.H"
#include "ReactorCvodeUtils.H"
#include "ReactorCvodeJacobian.H"
#include "PeleLM_K.H"  // reference Source/PeleLM_K.H (header only: reactionRateRhoY :57-101)

#include <sunmatrix/sunmatrix_dense.h>
#include <nvector/nvector_serial.h>

#include <cstring>
#include <memory>
#include <string>
#include <vector>

using namespace pele::physics;

namespace {
bool g_amrex_up = false;
std::unique_ptr<reactions::ReactorBase> g_reactor;
int g_reactor_type = 2;
} // namespace

extern "C" {

int ref_num_species() { return NUM_SPECIES; }
int ref_num_reactions() { return NUM_REACTIONS; }

// Bring AMReX up once (ParmParse, The_Sundials_Context, arenas).
int ref_setup()
{
  if (g_amrex_up) return 0;
  static std::vector<std::string> args = {"oracle_ref", "amrex.verbose=0", "amrex.v=0",
                                          "amrex.signal_handling=0", "amrex.throw_exception=1"};
  static std::vector<char*> argv;
  for (auto& a : args) argv.push_back(const_cast<char*>(a.c_str()));
  argv.push_back(nullptr);
  int argc = static_cast<int>(args.size());
  char** av = argv.data();
  amrex::Initialize(argc, av, true);
  g_amrex_up = true;
  return 0;
}

// (Re)create the reactor through the reference's own factory + ParmParse keys
// (ReactorCvode.cpp:15-22, :722-726; ReactorRK64.cpp:14-20; PeleLM.cpp:701-709).
// cells per flat (1dArray) react call the next ref_reactor_init sizes the reference's global N_Vector / user data for
// (ReactorCvode::init(reactor_type, ncells), ReactorCvode.cpp:33-38); 1 = the box API's per-cell problems
static int g_init_ncells = 1;
void ref_set_init_ncells(int n) { g_init_ncells = n > 0 ? n : 1; }

int ref_reactor_init(const char* chem_integrator, const char* solve_type, int reactor_type,
                     double rtol, double atol, int max_order, int max_nls_iters,
                     int clean_init_massfrac, int rk64_guess, int rk64_min, int rk64_max)
{
  ref_setup();
  if (g_reactor) {
    g_reactor->close();
    g_reactor.reset();
  }
  {
    amrex::ParmParse pp("ode");
    pp.add("verbose", 0);
    pp.add("rtol", rtol);
    pp.add("atol", atol);
    pp.add("max_nls_iters", max_nls_iters);
    pp.add("clean_init_massfrac", clean_init_massfrac);
    pp.add("rk64_nsubsteps_guess", rk64_guess);
    pp.add("rk64_nsubsteps_min", rk64_min);
    pp.add("rk64_nsubsteps_max", rk64_max);
    amrex::ParmParse ppcv("cvode");
    ppcv.add("solve_type", std::string(solve_type));
    ppcv.add("max_order", max_order);
  }
  g_reactor_type = reactor_type;
  g_reactor = reactions::ReactorBase::create(chem_integrator);
  return g_reactor->init(reactor_type, g_init_ncells);
}

int ref_set_typ_vals(const double* typ)
{
  if (!g_reactor) return 1;
  std::vector<amrex::Real> v(typ, typ + NUM_SPECIES + 1);
  g_reactor->set_typ_vals_ode(v);
  return 0;
}

// Box API. Arrays are AMReX FAB layout: component-major SoA over an nx*ny*nz box with lo = 0.
int ref_react_box(int nx, int ny, int nz, double* rY, double* rYsrc, double* T, double* rE,
                  double* rEsrc, double* FC, int* mask, double dt, double time)
{
  if (!g_reactor) return -1;
  amrex::Box box(amrex::IntVect(0, 0, 0), amrex::IntVect(nx - 1, ny - 1, nz - 1));
  const amrex::Dim3 lo = amrex::lbound(box);
  amrex::Dim3 hi = amrex::ubound(box);
  hi.x += 1; hi.y += 1; hi.z += 1;
  amrex::Array4<amrex::Real> a_rY(rY, lo, hi, NUM_SPECIES);
  amrex::Array4<amrex::Real> a_rYsrc(rYsrc, lo, hi, NUM_SPECIES);
  amrex::Array4<amrex::Real> a_T(T, lo, hi, 1);
  amrex::Array4<amrex::Real> a_rE(rE, lo, hi, 1);
  amrex::Array4<amrex::Real> a_rEsrc(rEsrc, lo, hi, 1);
  amrex::Array4<amrex::Real> a_FC(FC, lo, hi, 1);
  amrex::Array4<int> a_mask(mask, lo, hi, 1);
  amrex::Real dt_react = dt;
  amrex::Real t = time;
  return g_reactor->react(box, a_rY, a_rYsrc, a_T, a_rE, a_rEsrc, a_FC, a_mask, dt_react, t);
}

// reactionRateRhoY (PeleLM_K.H:57-101) over an nx*ny*nz box, FAB layout, MKS in and out.
void ref_reaction_rate_rhoY(int nx, int ny, int nz, const double* rhoY, const double* rhoH, const double* T,
                            const int* mask, double* rhoYdot)
{
  amrex::Box box(amrex::IntVect(0, 0, 0), amrex::IntVect(nx - 1, ny - 1, nz - 1));
  const amrex::Dim3 lo = amrex::lbound(box);
  amrex::Dim3 hi = amrex::ubound(box);
  hi.x += 1; hi.y += 1; hi.z += 1;
  amrex::Array4<const amrex::Real> a_rY(rhoY, lo, hi, NUM_SPECIES);
  amrex::Array4<const amrex::Real> a_rH(rhoH, lo, hi, 1);
  amrex::Array4<const amrex::Real> a_T(T, lo, hi, 1);
  amrex::Array4<const int> a_mask(mask, lo, hi, 1);
  amrex::Array4<amrex::Real> a_out(rhoYdot, lo, hi, NUM_SPECIES);
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) reactionRateRhoY(i, j, k, a_rY, a_rH, a_T, a_mask, a_out);
}

// Flat (1dArray) API: rY [ncells][N+1] incl. T, rYsrc [ncells][N], rX/rX_src [ncells].
int ref_react_flat(int ncells, double* rY, double* rYsrc, double* rX, double* rX_src, double dt,
                   double time)
{
  if (!g_reactor) return -1;
  amrex::Real dt_react = dt;
  amrex::Real t = time;
  return g_reactor->react(rY, rYsrc, rX, rX_src, dt_react, t, ncells);
}

// ODE right-hand side, one call per cell exactly as cF_RHS does on CPU (ReactorCvode.cpp:1765-1794).
void ref_rhs(int ncells, int reactor_type, double t, const double* y, const double* rhoe_init,
             const double* rhoesrc_ext, const double* rYsrc, double* ydot)
{
  for (int c = 0; c < ncells; ++c) {
    reactions::utils::fKernelSpec<reactions::utils::YCOrder>(
      0, 1, t, reactor_type, y + c * (NUM_SPECIES + 1), ydot + c * (NUM_SPECIES + 1),
      rhoe_init + c, rhoesrc_ext + c, rYsrc + c * NUM_SPECIES);
  }
}

// Dense analytic Jacobian as handed to CVODE (column-major (N+1)^2 per cell).
void ref_jac(int ncells, int reactor_type, const double* y, double* Jout)
{
  ref_setup();
  const int n = NUM_SPECIES + 1;
  auto& ctx = *amrex::sundials::The_Sundials_Context();
  N_Vector u = N_VNew_Serial(n, ctx);
  SUNMatrix J = SUNDenseMatrix(n, n, ctx);
  reactions::CVODEUserData ud{};
  ud.ncells = 1;
  ud.reactor_type = reactor_type;
  for (int c = 0; c < ncells; ++c) {
    std::memcpy(N_VGetArrayPointer(u), y + c * n, n * sizeof(double));
    SUNMatZero(J);
    reactions::cvode::cJac(0.0, u, nullptr, J, &ud, nullptr, nullptr, nullptr);
    std::memcpy(Jout + static_cast<size_t>(c) * n * n, SM_DATA_D(J), n * n * sizeof(double));
  }
  SUNMatDestroy(J);
  N_VDestroy(u);
}

// Molar production rates: C in mol/cm^3 -> wdot in mol/cm^3/s (mechanism.H CKWC).
void ref_ckwc(int ncells, const double* T, const double* C, double* wdot)
{
  for (int c = 0; c < ncells; ++c) {
    double Cl[NUM_SPECIES];
    for (int k = 0; k < NUM_SPECIES; ++k) Cl[k] = C[c * NUM_SPECIES + k];
    double Tl = T[c];
    CKWC(&Tl, Cl, wdot + c * NUM_SPECIES);
  }
}

// T from (h, Y) with the caller's guess, plus h_k, cp at the result (Fuego.H:128-134, :293-299, :36-42).
void ref_thermo(int ncells, const double* h, const double* Y, double* T_inout, double* hi,
                double* cp)
{
  auto eos = PhysicsType::eos();
  for (int c = 0; c < ncells; ++c) {
    double Yl[NUM_SPECIES];
    for (int k = 0; k < NUM_SPECIES; ++k) Yl[k] = Y[c * NUM_SPECIES + k];
    double Tl = T_inout[c];
    eos.RHY2T(1.0, h[c], Yl, Tl);
    T_inout[c] = Tl;
    eos.RTY2Hi(1.0, Tl, Yl, hi + c * NUM_SPECIES);
    eos.RTY2Cp(1.0, Tl, Yl, cp[c]);
  }
}

// rho, rho*h from (p, T, Y): used by the synthetic-state generators (Fuego.H PYT2R, TY2H).
void ref_pty2rh(int ncells, double p, const double* T, const double* Y, double* rho, double* h)
{
  auto eos = PhysicsType::eos();
  for (int c = 0; c < ncells; ++c) {
    double Yl[NUM_SPECIES];
    for (int k = 0; k < NUM_SPECIES; ++k) Yl[k] = Y[c * NUM_SPECIES + k];
    eos.PYT2R(p, Yl, T[c], rho[c]);
    eos.TY2H(T[c], Yl, h[c]);
  }
}

// VisMF::Write of a MultiFab made of nboxes boxes of side^3 cells laid along x, ncomp components, filled from
// data [box][comp][k][j][i] (FAB order).
int ref_vismf_write(const char* prefix, int nboxes, int side, int ncomp, const double* data)
{
  ref_setup();
  amrex::BoxList bl;
  for (int b = 0; b < nboxes; ++b)
    bl.push_back(amrex::Box(amrex::IntVect(b * side, 0, 0), amrex::IntVect((b + 1) * side - 1, side - 1, side - 1)));
  amrex::BoxArray ba(bl);
  amrex::DistributionMapping dm(ba);
  amrex::MultiFab mf(ba, dm, ncomp, 0);
  const size_t per_box = size_t(side) * side * side * ncomp;
  for (amrex::MFIter mfi(mf); mfi.isValid(); ++mfi)
    std::memcpy(mf[mfi].dataPtr(), data + size_t(mfi.index()) * per_box, per_box * sizeof(double));
  amrex::VisMF::Write(mf, prefix);
  return 0;
}

// VisMF::Read: meta = {nboxes, ncomp, then lo[3] hi[3] per box}; out = the FABs back to back. Returns doubles written.
long ref_vismf_read(const char* prefix, long max_doubles, double* out, int* meta, int max_meta)
{
  ref_setup();
  amrex::MultiFab mf;
  amrex::VisMF::Read(mf, prefix);
  const int nb = static_cast<int>(mf.boxArray().size());
  if (2 + 6 * nb > max_meta) return -1;
  meta[0] = nb;
  meta[1] = mf.nComp();
  long off = 0;
  for (int b = 0; b < nb; ++b) {
    const amrex::Box& bx = mf.boxArray()[b];
    for (int d = 0; d < 3; ++d) {
      meta[2 + 6 * b + d] = bx.smallEnd(d);
      meta[2 + 6 * b + 3 + d] = bx.bigEnd(d);
    }
    const long cnt = bx.numPts() * mf.nComp();
    if (off + cnt > max_doubles) return -1;
    std::memcpy(out + off, mf[b].dataPtr(), cnt * sizeof(double));
    off += cnt;
  }
  return off;
}

void ref_species_names(char* buf, int buflen)
{
  amrex::Vector<std::string> names;
  CKSYMS_STR(names);
  std::string s;
  for (auto& n : names) { s += n; s += ' '; }
  std::strncpy(buf, s.c_str(), buflen - 1);
  buf[buflen - 1] = 0;
}

void ref_finalize()
{
  if (g_reactor) {
    g_reactor->close();
    g_reactor.reset();
  }
  if (g_amrex_up) {
    amrex::Finalize();
    g_amrex_up = false;
  }
}

} // extern "C"
