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
  // cv_R / speciesInternalEnergy: the same polynomials with the constant term a0 - 1 (tables: a0m1)
  std::vector<double> cpe = cp, he = h;
  for (int k = 0; k < ns; ++k)
    for (int r = 0; r < 2; ++r) {
      cpe[(size_t(r) * 5 + 0) * nsp + k] = t->a0m1[k * 2 + r];
      he[(size_t(r) * 6 + 0) * nsp + k] = t->a0m1[k * 2 + r];
    }

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
    if (rs[r] == 0xffffffffu || ps[r] == 0xffffffffu) return (std::fprintf(stderr, "plm: build_dev_mech check 1 failed\n"), -1);
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
  (void)maxlen;
  std::vector<int> noff(t->noff, t->noff + nr + 1), nidx(t->nidx, t->nidx + t->noff[nr]);
  std::vector<double> nnu(t->nval, t->nval + t->noff[nr]);
  noff.resize(nrp + 1, t->noff[nr]);

  // ---- packed records for the rate evaluation ----
  if (d.n3b_end > 32) {
    std::fprintf(stderr, "plm: more than 32 falloff/three-body reactions are not supported by the warp layout\n");
    return (std::fprintf(stderr, "plm: build_dev_mech check 2 failed\n"), -1);
  }
  std::vector<double2> rp0(nrp, make_double2(0.0, 0.0)), rp1(nrp, make_double2(0.0, 0.0));
  std::vector<double2> fo0(32, make_double2(0.0, 0.0)), fo1(32, make_double2(0.0, 0.0)), fo2(32, make_double2(0.0, 0.0)),
      fo3(32, make_double2(0.0, 0.0));
  std::vector<double> fo4(32, 0.0);
  for (int r = 0; r < nrp; ++r) {
    uint64_t bits = uint64_t(rs[r]) | (uint64_t(ps[r]) << 24);
    if (r < nr) {
      bits |= uint64_t(flags[r] & 3) << 48;
      bits |= uint64_t((flags[r] >> 2) & 1) << 50;
      int dnr = dn[r];
      if (!(flags[r] & 4)) dnr = 0;  // only the reverse rate uses refC^(-dn)
      if (dnr < -2 || dnr > 2) return (std::fprintf(stderr, "plm: build_dev_mech check 3 failed\n"), -1);
      bits |= uint64_t(dnr + 2) << 52;
      rp0[r] = make_double2(A[r], b[r]);
    } else {
      bits |= uint64_t(2) << 52;
    }
    double bd;
    std::memcpy(&bd, &bits, 8);
    rp1[r] = make_double2(EaR[r], bd);
  }
  for (int r = 0; r < std::min(nr, 32); ++r) {
    if ((flags[r] & 3) < 2) continue;
    fo0[r] = make_double2(A0[r], b0[r]);
    const bool troe = (flags[r] & 3) == 3;
    fo1[r] = make_double2(Ea0R[r], (troe && (flags[r] & 16)) ? 1.0 - ta[r] : 0.0);
    fo2[r] = make_double2(troe ? tiT3[r] : 0.0, (troe && (flags[r] & 32)) ? ta[r] : 0.0);
    fo3[r] = make_double2(troe ? tiT1[r] : 0.0, (troe && (flags[r] & 8)) ? tT2[r] : 0.0);
    fo4[r] = (troe && (flags[r] & 8)) ? 1.0 : 0.0;
  }
  int emax = 0;
  for (int r = 0; r < nr; ++r) emax = std::max(emax, eoff[r + 1] - eoff[r]);
  d.emax = emax;
  std::vector<double2> eff(size_t(std::max(emax, 1)) * 32);
  {
    const uint64_t unit = uint64_t(ns);
    double ud;
    std::memcpy(&ud, &unit, 8);
    for (auto& e : eff) e = make_double2(0.0, ud);
    for (int r = 0; r < std::min(nr, 32); ++r)
      for (int s = eoff[r], j = 0; s < eoff[r + 1]; ++s, ++j) {
        const uint64_t idx = uint64_t(eidx[s]);
        double id;
        std::memcpy(&id, &idx, 8);
        eff[size_t(j) * 32 + r] = make_double2(em1[s], id);
      }
  }
  // ---- load-balanced production-rate gather ----
  // lane-slots available: nsp; every species gets >= 1 segment, the rest go to the longest lists
  std::vector<int> nseg(ns, 1);
  {
    int free_slots = nsp - ns;
    auto seglen = [&](int k) { const int len = soff[k + 1] - soff[k]; return (len + nseg[k] - 1) / nseg[k]; };
    while (free_slots > 0) {
      int best = -1, bl = 0;
      for (int k = 0; k < ns; ++k)
        if (nseg[k] < kMaxSeg && seglen(k) > bl) {
          bl = seglen(k);
          best = k;
        }
      if (best < 0 || bl <= 4) break;
      nseg[best]++;
      free_slots--;
    }
    int glen = 1;
    for (int k = 0; k < ns; ++k) glen = std::max(glen, seglen(k));
    d.glen = glen;
  }
  const uint32_t zero_q = uint32_t(nrp) * 8u;       // q[nr_pad] is kept at 0
  const int zero_part = nsp * 8;                    // part[nsp_pad] is kept at 0
  std::vector<uint2> gat(size_t(d.glen) * nsp, make_uint2(zero_q, 0u));
  std::vector<int> seg(size_t(kMaxSeg) * nsp, zero_part);
  {
    int slot = 0;
    for (int k = 0; k < ns; ++k) {
      const int len = soff[k + 1] - soff[k];
      const int per = (len + nseg[k] - 1) / nseg[k];
      for (int sgi = 0; sgi < nseg[k]; ++sgi, ++slot) {
        seg[size_t(sgi) * nsp + k] = slot * 8;
        for (int j = 0; j < per; ++j) {
          const int e = soff[k] + sgi * per + j;
          if (e >= soff[k + 1]) break;
          uint64_t nb;
          std::memcpy(&nb, &snu[e], 8);
          if (uint32_t(nb) != 0u) return (std::fprintf(stderr, "plm: build_dev_mech check 4 failed\n"), -1);  // net nu must be a small integer (low word zero)
          gat[size_t(j) * nsp + slot] = make_uint2(uint32_t(sreac[e]) * 8u, uint32_t(nb >> 32));
        }
      }
    }
    if (slot > nsp) return (std::fprintf(stderr, "plm: build_dev_mech check 5 failed\n"), -1);
  }

  Blob bl;
  bl.bytes.resize(512, 0);  // room for the DevMech struct itself at offset 0
  const size_t o_mw = bl.add(mw), o_imw = bl.add(imw), o_tmid = bl.add(tmid), o_cp = bl.add(cp), o_h = bl.add(h),
               o_g = bl.add(g), o_dcp = bl.add(dcp), o_A = bl.add(A), o_b = bl.add(b), o_E = bl.add(EaR),
               o_A0 = bl.add(A0), o_b0 = bl.add(b0), o_E0 = bl.add(Ea0R), o_ta = bl.add(ta), o_i3 = bl.add(tiT3),
               o_i1 = bl.add(tiT1), o_t3 = bl.add(tT3), o_t1 = bl.add(tT1), o_t2 = bl.add(tT2), o_fl = bl.add(flags),
               o_dn = bl.add(dn), o_rs = bl.add(rs), o_ps = bl.add(ps), o_eo = bl.add(eoff), o_ei = bl.add(eidx),
               o_em = bl.add(em1), o_so = bl.add(soff), o_sr = bl.add(sreac), o_sn = bl.add(snu),
               o_no = bl.add(noff), o_ni = bl.add(nidx), o_nn = bl.add(nnu), o_rp0 = bl.add(rp0), o_rp1 = bl.add(rp1),
               o_fo0 = bl.add(fo0), o_fo1 = bl.add(fo1), o_fo2 = bl.add(fo2), o_fo3 = bl.add(fo3), o_fo4 = bl.add(fo4),
               o_eff = bl.add(eff), o_gat = bl.add(gat), o_seg = bl.add(seg), o_cpe = bl.add(cpe), o_he = bl.add(he);
  (void)o_dcp; (void)o_so; (void)o_sr; (void)o_sn;
  static_assert(sizeof(DevMech) <= 512, "DevMech header");
  void* dev = nullptr;
  if (cudaMalloc(&dev, bl.bytes.size()) != cudaSuccess) return -2;
  auto* base = static_cast<unsigned char*>(dev);
  auto D = [&](size_t o) { return reinterpret_cast<const double*>(base + o); };
  auto I = [&](size_t o) { return reinterpret_cast<const int*>(base + o); };
  d.mw = D(o_mw); d.imw = D(o_imw); d.tmid = D(o_tmid);
  d.cp = D(o_cp); d.h = D(o_h); d.g = D(o_g);
  d.eint = 0;
  out->cp_h = D(o_cp); out->h_h = D(o_h); out->cp_e = D(o_cpe); out->h_e = D(o_he);
  auto D2 = [&](size_t o) { return reinterpret_cast<const double2*>(base + o); };
  d.rp0 = D2(o_rp0); d.rp1 = D2(o_rp1); d.fo0 = D2(o_fo0); d.fo1 = D2(o_fo1); d.fo2 = D2(o_fo2); d.fo3 = D2(o_fo3);
  d.fo4 = D(o_fo4); d.eff = D2(o_eff);
  d.gat = reinterpret_cast<const uint2*>(base + o_gat);
  d.seg = I(o_seg);
  d.A = D(o_A); d.b = D(o_b); d.EaR = D(o_E); d.A0 = D(o_A0); d.b0 = D(o_b0); d.Ea0R = D(o_E0);
  d.troe_a = D(o_ta); d.troe_iT3 = D(o_i3); d.troe_iT1 = D(o_i1); d.troe_T3 = D(o_t3); d.troe_T1 = D(o_t1);
  d.troe_T2 = D(o_t2);
  d.flags = I(o_fl); d.dn = I(o_dn);
  d.rslots = reinterpret_cast<const uint32_t*>(base + o_rs);
  d.pslots = reinterpret_cast<const uint32_t*>(base + o_ps);
  d.eoff = I(o_eo); d.eidx = I(o_ei); d.em1 = D(o_em);
  d.noff = I(o_no); d.nidx = I(o_ni); d.nnu = D(o_nn);
  std::memcpy(bl.bytes.data(), &d, sizeof(DevMech));
  if (cudaMemcpy(dev, bl.bytes.data(), bl.bytes.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(dev);
    return -2;
  }
  out->h = d;
  out->d = static_cast<const DevMech*>(dev);
  out->blob = dev;
  out->blob_bytes = bl.bytes.size();
  return 0;
}

int select_reactor_type(DevMechOwner* o, int reactor_type)
{
  if (!o || !o->blob) return -1;
  const bool e = reactor_type == 1;
  o->h.eint = e ? 1 : 0;
  o->h.cp = e ? o->cp_e : o->cp_h;
  o->h.h = e ? o->h_e : o->h_h;
  return cudaMemcpy(o->blob, &o->h, sizeof(DevMech), cudaMemcpyHostToDevice) == cudaSuccess ? 0 : -2;
}

void free_dev_mech(DevMechOwner* o)
{
  if (o->blob) cudaFree(o->blob);
  o->blob = nullptr;
}

}  // namespace plm
