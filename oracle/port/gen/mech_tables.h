/* Flat mechanism tables (data only). Layout written by pelelm_b200/mechgen/emit_tables.py.
 * Reactions are stored in the evaluation order of the reference's generated kernels
 * (Troe falloff, Lindemann falloff, three-body, elementary; ceptr/reaction_info.py:30-95).
 * Units: rate constants in mol/m^3, s; Ea/R in K; NASA-7 tables pre-divided as the generated
 * reference prints them (cp: a0..a4; h: a0,a1/2..a4/5,a5; g: a5,a0-a6,-a0,-a1/2,-a2/6,-a3/12,-a4/20;
 * dcp: a1,2a2,3a3,4a4); index [species][0=low,1=high][coef]. */
#ifndef PLM_MECH_TABLES_H
#define PLM_MECH_TABLES_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct plm_mech_tables {
  const char* name;
  int ns, nr;
  int n_troe, n_lind, n_3b, n_simple;
  const char* const* species;
  const double *mw, *tmid, *cp, *h, *g, *dcp;
  const double* a0m1; /* [species][0=low,1=high]: constant term of cv/R and e/RT (= a0 - 1 in decimal) */
  const double *A, *b, *EaR, *A0, *b0, *Ea0R, *troe;
  const int *troe_np, *type, *rev;
  const double* dn;
  const int* orig;
  const int *roff, *ridx; const double* rval; /* reactants, CSR, sorted by species */
  const int *poff, *pidx; const double* pval; /* products */
  const int *noff, *nidx; const double* nval; /* net stoichiometry (products - reactants) */
  const int *eoff, *eidx; const double* eval; /* third-body efficiencies */
} plm_mech_tables;
extern const plm_mech_tables plm_tables_LiDryer;
extern const plm_mech_tables plm_tables_drm19;
extern const plm_mech_tables plm_tables_dodecane_lu;
#ifdef __cplusplus
}
#endif
#endif
