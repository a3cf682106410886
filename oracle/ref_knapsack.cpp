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
