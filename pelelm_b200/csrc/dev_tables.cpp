// Builds the device layout of a mechanism (see dev_tables.h) from the flat generated tables.
#include "dev_tables.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

namespace plm {
namespace {

struct Blob {
  std::vector<unsigned char> bytes;
  template <class T>
  size_t add(const std::vector<T>& v)
  {
    size_t off = (bytes.size() + 127) & ~size_t(127);  // 128-byte aligned sections
    bytes.resize(off + v.size() * sizeof(T));
    if (!v.empty()) std::memcpy(bytes.data() + off, v.data(), v.size() * sizeof(T));
    return off;
  }
};

}  // namespace

int build_dev_mech(const plm_mech_tables* t, DevMechOwner* out)
{
  const int ns = t->ns, nr = t->nr;
  const int nsp = (ns + 1 + 31) / 32 * 32;  // +1: room for the unit/zero pad entry at index ns
  const int nrp = (nr + 31) / 32 * 32;
  DevMech d{};
  d.ns = ns;
  d.n = ns + 1;
  d.nr = nr;
  d.nsp_pad = nsp;
  d.nr_pad = nrp;
  d.nfall = t->n_troe + t->n_lind;
  d.n3b_end = d.nfall + t->n_3b;

  std::vector<double> mw(nsp, 1.0), imw(nsp, 0.0), tmid(nsp, 1000.0);
  for (int k = 0; k < ns; ++k) {
    mw[k] = t->mw[k];
    imw[k] = 1.0 / t->mw[k];  // get_imw: "1.0 / 2.016000" (mechanism.H:88-113)
    tmid[k] = t->tmid[k];
  }
  auto plane = [&](const double* src, int ncoef) {
    std::vector<double> v(size_t(2) * ncoef * nsp, 0.0);
    for (int k = 0; k < ns; ++k)
      for (int r = 0; r < 2; ++r)
        for (int c = 0; c < ncoef; ++c) v[(size_t(r) * ncoef + c) * nsp + k] = src[(k * 2 + r) * ncoef + c];
    return v;
  };
  std::vector<double> cp = plane(t->cp, 5), h = plane(t->h, 6), g = plane(t->g, 7), dcp = plane(t->dcp, 4);

  std::vector<double> A(nrp, 0.0), b(nrp, 0.0), EaR(nrp, 0.0), A0(nrp, 0.0), b0(nrp, 0.0), Ea0R(nrp, 0.0);
  std::vector<double> ta(nrp, 0.0), tiT3(nrp, 0.0), tiT1(nrp, 0.0), tT3(nrp, 1.0), tT1(nrp, 1.0), tT2(nrp, 0.0);
  std::vector<int> flags(nrp, 0), dn(nrp, 0);
  std::vector<uint32_t> rs(nrp, 0), ps(nrp, 0);
  const uint32_t pad = uint32_t(ns);
  for (int r = 0; r < nrp; ++r) rs[r] = ps[r] = pad | (pad << 8) | (pad << 16);
  for (int r = 0; r < nr; ++r) {
    A[r] = t->A[r];
    b[r] = t->b[r];
    EaR[r] = t->EaR[r];
    A0[r] = t->A0[r];
    b0[r] = t->b0[r];
    Ea0R[r] = t->Ea0R[r];
    int f = t->type[r] & 3;
    if (t->rev[r]) f |= 4;
    if (t->type[r] == 3) {
      const double* tr = t->troe + 4 * r;
      ta[r] = tr[0];
      tT3[r] = tr[1];
      tT1[r] = tr[2];
      tT2[r] = tr[3];
      tiT3[r] = 1 / tr[1];
      tiT1[r] = 1 / tr[2];
      if (t->troe_np[r] == 4) f |= 8;
      // production.py:627-650 term selection
      if (std::abs(tr[1]) > 1.0e-100 && (1.0 - tr[0]) != 0) f |= 16;
      if (std::abs(tr[2]) > 1.0e-100 && tr[0] != 0) f |= 32;
    }
    flags[r] = f;
    dn[r] = int(t->dn[r]);
    auto pack = [&](const int* off, const int* idx, const double* val, bool required) -> uint32_t {
      int slot = 0;
      uint32_t w = pad | (pad << 8) | (pad << 16);
      for (int s = off[r]; s < off[r + 1]; ++s) {
        const int mult = int(val[s]);
        if (double(mult) != val[s]) {
          std::fprintf(stderr, "plm: non-integer stoichiometric coefficient in reaction %d\n", r);
          return 0xffffffffu;
        }
        for (int c = 0; c < mult; ++c) {
          if (slot >= kMaxSlots) {
            if (required) return 0xffffffffu;
            return w;  // irreversible products beyond 3 slots never enter a rate expression
          }
          w &= ~(0xffu << (8 * slot));
          w |= uint32_t(idx[s]) << (8 * slot);
          ++slot;
        }
      }
      return w;
    };
    rs[r] = pack(t->roff, t->ridx, t->rval, true);
    ps[r] = pack(t->poff, t->pidx, t->pval, t->rev[r] != 0);
    if (rs[r] == 0xffffffffu || ps[r] == 0xffffffffu) return -1;
  }
  // third-body lists
  std::vector<int> eoff(nrp + 1, 0), eidx;
  std::vector<double> em1;
  for (int r = 0; r < nr; ++r) {
    for (int s = t->eoff[r]; s < t->eoff[r + 1]; ++s) {
      eidx.push_back(t->eidx[s]);
      em1.push_back(t->eval[s] - 1.0);
    }
    eoff[r + 1] = int(eidx.size());
  }
  for (int r = nr; r < nrp; ++r) eoff[r + 1] = eoff[nr];
  // species -> reactions (ascending reaction index = the accumulation order of productionRate)
  std::vector<int> soff(nsp + 1, 0), sreac;
  std::vector<double> snu;
  int maxlen = 0;
  for (int k = 0; k < ns; ++k) {
    for (int r = 0; r < nr; ++r)
      for (int s = t->noff[r]; s < t->noff[r + 1]; ++s)
        if (t->nidx[s] == k) {
          sreac.push_back(r);
          snu.push_back(t->nval[s]);
        }
    soff[k + 1] = int(sreac.size());
    maxlen = std::max(maxlen, soff[k + 1] - soff[k]);
  }
  for (int k = ns; k < nsp; ++k) soff[k + 1] = soff[ns];
  d.max_sp_reac = maxlen;
  std::vector<int> noff(t->noff, t->noff + nr + 1), nidx(t->nidx, t->nidx + t->noff[nr]);
  std::vector<double> nnu(t->nval, t->nval + t->noff[nr]);
  noff.resize(nrp + 1, t->noff[nr]);

  Blob bl;
  const size_t o_mw = bl.add(mw), o_imw = bl.add(imw), o_tmid = bl.add(tmid), o_cp = bl.add(cp), o_h = bl.add(h),
               o_g = bl.add(g), o_dcp = bl.add(dcp), o_A = bl.add(A), o_b = bl.add(b), o_E = bl.add(EaR),
               o_A0 = bl.add(A0), o_b0 = bl.add(b0), o_E0 = bl.add(Ea0R), o_ta = bl.add(ta), o_i3 = bl.add(tiT3),
               o_i1 = bl.add(tiT1), o_t3 = bl.add(tT3), o_t1 = bl.add(tT1), o_t2 = bl.add(tT2), o_fl = bl.add(flags),
               o_dn = bl.add(dn), o_rs = bl.add(rs), o_ps = bl.add(ps), o_eo = bl.add(eoff), o_ei = bl.add(eidx),
               o_em = bl.add(em1), o_so = bl.add(soff), o_sr = bl.add(sreac), o_sn = bl.add(snu),
               o_no = bl.add(noff), o_ni = bl.add(nidx), o_nn = bl.add(nnu);
  void* dev = nullptr;
  if (cudaMalloc(&dev, bl.bytes.size()) != cudaSuccess) return -2;
  if (cudaMemcpy(dev, bl.bytes.data(), bl.bytes.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(dev);
    return -2;
  }
  auto* base = static_cast<unsigned char*>(dev);
  auto D = [&](size_t o) { return reinterpret_cast<const double*>(base + o); };
  auto I = [&](size_t o) { return reinterpret_cast<const int*>(base + o); };
  d.mw = D(o_mw); d.imw = D(o_imw); d.tmid = D(o_tmid);
  d.cp = D(o_cp); d.h = D(o_h); d.g = D(o_g); d.dcp = D(o_dcp);
  d.A = D(o_A); d.b = D(o_b); d.EaR = D(o_E); d.A0 = D(o_A0); d.b0 = D(o_b0); d.Ea0R = D(o_E0);
  d.troe_a = D(o_ta); d.troe_iT3 = D(o_i3); d.troe_iT1 = D(o_i1); d.troe_T3 = D(o_t3); d.troe_T1 = D(o_t1);
  d.troe_T2 = D(o_t2);
  d.flags = I(o_fl); d.dn = I(o_dn);
  d.rslots = reinterpret_cast<const uint32_t*>(base + o_rs);
  d.pslots = reinterpret_cast<const uint32_t*>(base + o_ps);
  d.eoff = I(o_eo); d.eidx = I(o_ei); d.em1 = D(o_em);
  d.soff = I(o_so); d.sreac = I(o_sr); d.snu = D(o_sn);
  d.noff = I(o_no); d.nidx = I(o_ni); d.nnu = D(o_nn);
  out->h = d;
  out->blob = dev;
  out->blob_bytes = bl.bytes.size();
  return 0;
}

void free_dev_mech(DevMechOwner* o)
{
  if (o->blob) cudaFree(o->blob);
  o->blob = nullptr;
}

}  // namespace plm
