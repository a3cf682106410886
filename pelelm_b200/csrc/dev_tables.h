// Device-side mechanism tables: the flat host tables (gen/mech_tables.h) re-laid out for
// warp-cooperative evaluation -- arrays indexed by lane so that a warp's loads are contiguous.
//
//   lane = species   : thermo coefficient planes [range][coef][nsp_pad], mw/imw/tmid [nsp_pad]
//   lane = reaction  : packed 2 x 16-byte records per reaction (A,b | Ea/R, slot bits), falloff rows for the
//                      first 32 reactions (all falloff / three-body reactions sit there), padded
//                      third-body efficiency rows [emax][32]
//   production rates : load-balanced gather -- every lane-slot sums ONE segment of ONE species' reaction list
//                      (reaction order preserved inside a segment), then the species lane adds its <= 4 segments
//   Jacobian scatter : reaction -> net stoichiometry CSR (+ the unpacked per-reaction arrays it reads)
#pragma once
#include <cstddef>
#include <cstdint>

#include <cuda_runtime.h>

#include "gen/mech_tables.h"

namespace plm {

constexpr int kMaxSlots = 3;   // reactant slots; reversible reactions also have <= 3 product slots
constexpr int kMaxSeg = 4;     // a species' reaction list is split into at most this many gather segments

struct DevMech {
  int ns;       // species
  int n;        // ns + 1 (ODE size)
  int nr;       // reactions
  int nsp_pad;  // species (+1 unit slot) padded to a multiple of 32
  int nr_pad;   // reactions padded to a multiple of 32
  int nfall;    // first nfall reactions are falloff (Troe then Lindemann)
  int n3b_end;  // reactions [nfall, n3b_end) are three-body; n3b_end <= 32
  int emax;     // longest third-body efficiency list
  int glen;     // gather segment length (entries per lane-slot)
  int eint;     // 1: internal-energy reactor (ode.reactor_type = 1): cp/h below are the cv/R and e/RT planes
  // species
  const double *mw, *imw, *tmid;
  const double *cp;   // [2][5][nsp_pad]
  const double *h;    // [2][6][nsp_pad]
  const double *g;    // [2][7][nsp_pad]
  // reactions, packed for the rate evaluation
  const double2 *rp0;   // [nr_pad] {A, b}
  const double2 *rp1;   // [nr_pad] {Ea/R, bits}: bits 0-23 reactant slots, 24-47 product slots (species index, ns = unit
                        //          entry), 48-49 type (0 simple, 1 three-body, 2 Lindemann, 3 Troe), 50 reversible,
                        //          52-54 dn + 2
  const double2 *fo0, *fo1, *fo2, *fo3;  // [32] {A0,b0} {Ea0/R, (1-a) or 0} {1/T3, a or 0} {1/T1, T2 (0 if absent)}
  const double  *fo4;                    // [32] 1 if the T2 term is present else 0
  const double2 *eff;   // [emax][32] {efficiency - 1, species index as double bits}; padding = {0, unit slot}
  const uint2   *gat;   // [glen][nsp_pad] {byte offset of the reaction in q[], high word of the net nu}; padding -> zero slot
  const int     *seg;   // [kMaxSeg][nsp_pad] byte offset into part[] of the species' segments; padding -> zero slot
  // unpacked arrays (analytic Jacobian)
  const double *A, *b, *EaR, *A0, *b0, *Ea0R;
  const double *troe_a, *troe_iT3, *troe_iT1, *troe_T3, *troe_T1, *troe_T2;
  const int* flags;    // bits 0-1 type, bit 2 reversible, bit 3 Troe has T2, bit 4 Troe term-1 present, bit 5 term-2 present
  const int* dn;
  const uint32_t *rslots, *pslots;
  const int *eoff, *eidx; const double* em1;
  const int *noff, *nidx; const double* nnu;   // reaction -> (species, net nu)
};

// Host-side owner: builds the arrays from the flat tables and uploads them. Returns nonzero on failure.
struct DevMechOwner {
  const double *cp_h = nullptr, *h_h = nullptr;  // cp/R, h/RT planes (enthalpy reactor)
  const double *cp_e = nullptr, *h_e = nullptr;  // cv/R, e/RT planes (internal-energy reactor): constant term a0 - 1
  DevMech h{};        // host copy of the struct holding DEVICE pointers
  const DevMech* d = nullptr;  // the same struct in device memory (first bytes of the blob)
  void* blob = nullptr;
  size_t blob_bytes = 0;
};
int build_dev_mech(const plm_mech_tables* t, DevMechOwner* out);   // 0 ok
void free_dev_mech(DevMechOwner* o);
// ode.reactor_type: 1 = internal energy (REY2T / Ei / Cv; consP = 0 in the Jacobian), 2 = enthalpy. Re-points the thermo
// planes of the descriptor and refreshes its device copy. Returns nonzero on failure.
int select_reactor_type(DevMechOwner* o, int reactor_type);

}  // namespace plm
