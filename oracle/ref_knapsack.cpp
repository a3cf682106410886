// TEST INFRASTRUCTURE -- exposes the reference's file-static knapsack()
// (amrex/Src/Base/AMReX_DistributionMapping.cpp:576-716) by compiling that translation unit, unmodified
// and where it lies, through an #include; this TU replaces AMReX_DistributionMapping.o in oracle/_ref.
#include <AMReX_DistributionMapping.cpp>

extern "C" void ref_knapsack(const long* wgts, int nboxes, int nprocs, int do_full, int nmax, int* bucket_of_box,
                             double* efficiency)
{
  std::vector<amrex::Long> w(wgts, wgts + nboxes);
  std::vector<std::vector<int>> result;
  amrex::Real eff = 0;
  amrex::knapsack(w, nprocs, result, eff, do_full != 0, nmax);
  for (int b = 0; b < static_cast<int>(result.size()); ++b)
    for (int id : result[b]) bucket_of_box[id] = b;
  if (efficiency) *efficiency = eff;
}

// BoxList::maxSize (amrex/Src/Base/AMReX_BoxList.cpp:774-827) on one cell-centred box: the pieces, in the reference's
// order, as lo/hi triples. Returns the number of pieces (at most cap are written).
#include <AMReX_BoxList.H>
extern "C" int ref_max_size(const int* lo, const int* hi, const int* chunk, int* out_lohi, int cap)
{
  amrex::BoxList bl;
  bl.push_back(amrex::Box(amrex::IntVect(AMREX_D_DECL(lo[0], lo[1], lo[2])), amrex::IntVect(AMREX_D_DECL(hi[0], hi[1], hi[2]))));
  bl.maxSize(amrex::IntVect(AMREX_D_DECL(chunk[0], chunk[1], chunk[2])));
  int n = 0;
  for (const auto& b : bl) {
    if (n < cap)
      for (int d = 0; d < 3; ++d) {
        out_lohi[6 * n + d] = d < AMREX_SPACEDIM ? b.smallEnd(d) : 0;
        out_lohi[6 * n + 3 + d] = d < AMREX_SPACEDIM ? b.bigEnd(d) : 0;
      }
    ++n;
  }
  return n;
}
