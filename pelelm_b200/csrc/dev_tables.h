// Device-side mechanism tables: the flat host tables (gen/mech_tables.h) re-laid out for
// warp-cooperative evaluation -- SoA arrays indexed by lane so that a warp's loads are contiguous.
//
//   lane = species   : thermo coefficient planes [range][coef][nsp_pad], mw/imw/tmid [nsp_pad],
//                      species -> reactions CSR (wdot gather, reaction order preserved)
//   lane = reaction  : Arrhenius / falloff parameters [nr_pad], reactant/product slot indices
//                      (nu = 2 becomes a repeated slot; unused slots point at a constant-1 entry),
//                      third-body efficiency CSR
//   Jacobian scatter : reaction -> net stoichiometry CSR
#pragma once
#include <cstddef>
#include <cstdint>

#include "gen/mech_tables.h"

namespace plm {

constexpr int kMaxSlots = 3;   // reactant slots; reversible reactions also have <= 3 product slots

struct DevMech {
  int ns;       // species
  int n;        // ns + 1 (ODE size)
  int nr;       // reactions
  int nsp_pad;  // species padded to a multiple of 32
  int nr_pad;   // reactions padded to a multiple of 32
  int nfall;    // first nfall reactions are falloff (Troe then Lindemann)
  int n3b_end;  // reactions [nfall, n3b_end) are three-body
  int max_sp_reac;  // longest species->reaction list
  // species
  const double *mw, *imw, *tmid;
  const double *cp;   // [2][5][nsp_pad]
  const double *h;    // [2][6][nsp_pad]
  const double *g;    // [2][7][nsp_pad]
  const double *dcp;  // [2][4][nsp_pad]
  // reactions
  const double *A, *b, *EaR;           // [nr_pad]
  const double *A0, *b0, *Ea0R;        // [nfall padded to nr_pad]
  const double *troe_a, *troe_iT3, *troe_iT1, *troe_T3, *troe_T1, *troe_T2;  // [nr_pad]
  const int* flags;    // [nr_pad] bits 0-1 type (0 simple,1 3-body,2 Lindemann,3 Troe), bit 2 reversible,
                       //          bit 3 Troe has T2, bit 4 Troe term-1 present, bit 5 Troe term-2 present
  const int* dn;       // [nr_pad] sum nu'' - sum nu'
  const uint32_t* rslots;  // [nr_pad] 3 x 8-bit reactant species (ns = unit entry)
  const uint32_t* pslots;  // [nr_pad] 3 x 8-bit product species (only meaningful when reversible)
  const int *eoff, *eidx; const double* em1;   // third-body lists, value = efficiency - 1
  const int *soff, *sreac; const double* snu;  // species -> (reaction, net nu), reaction-ordered
  const int *noff, *nidx; const double* nnu;   // reaction -> (species, net nu)
};

// Host-side owner: builds the arrays from the flat tables and uploads them. Returns nullptr on CUDA failure.
struct DevMechOwner {
  DevMech h{};        // host copy of the struct holding DEVICE pointers
  void* blob = nullptr;
  size_t blob_bytes = 0;
};
int build_dev_mech(const plm_mech_tables* t, DevMechOwner* out);   // 0 ok
void free_dev_mech(DevMechOwner* o);

}  // namespace plm
