// C ABI of libplmchem.so (include/plm_chem.h): argument checking, host/device pointer staging,
// launch configuration, statistics. No chemistry is computed on the host.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/plm_chem.h"
#include "kernels.h"

namespace {

// NVTX ranges named like the reference's BL_PROFILE regions (ReactorCvode.cpp:1156, :1407, :1477; PeleLM.cpp:5632), so that a
// timeline of a run through the adaptor reads like the reference's TinyProfiler output
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

thread_local std::string g_err;
int fail(int code, const std::string& msg)
{
  g_err = msg;
  return code;
}
#define PLM_CUDA(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess) return fail(PLM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)

const plm_mech_tables* find_mech(const char* name)
{
  if (!name) return nullptr;
  if (!std::strcmp(name, "LiDryer")) return &plm_tables_LiDryer;
  if (!std::strcmp(name, "drm19")) return &plm_tables_drm19;
  if (!std::strcmp(name, "dodecane_lu")) return &plm_tables_dodecane_lu;
  return nullptr;
}

bool is_device_ptr(const void* p)
{
  if (!p) return false;
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes)
  {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return -1;
    cap = bytes;
    return 0;
  }
  void release()
  {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

constexpr int kSlots = 16;   // react calls that may be in flight at once (one per stream) on one reactor
constexpr int kStatWords = 16;

// per-call resources: a call owns one slot from launch until its kernel has finished
struct CallSlot {
  DevBuf status;   // int per cell
  DevBuf jglobal;  // saved Jacobians when they do not fit in shared memory
  DevBuf lanews;   // lane = cell path: per-cell vectors and matrices of every resident lane
  DevBuf fcflat;   // 1dArray API: FC of the call (the API returns it through the statistics only)
  DevBuf boxtab;   // plm_react_boxes: the boxes of the call (BoxArgs[nboxes], then the cell offsets int64[nboxes + 1])
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool used = false;
  int64_t ncells = 0;
};

struct plm_reactor {
  const plm_mech_tables* tables = nullptr;
  plm::DevMechOwner mech;
  int integ = plm::kIntegCvode;
  bool inited = false;
  int reactor_type = 2;
  plm::ReactOpts opts{};
  int num_sms = 0;
  DevBuf ctrl;     // [kSlots] work-queue counters, then kStatWords cumulative statistics (unsigned long long)
  CallSlot slot[kSlots];
  int next_slot = 0, last_slot = -1;
  unsigned long long* host_stats = nullptr;  // pinned; [kStatWords] = the failed-cell counter as copied after each launch
  long long failed_total = 0;                // failed cells already collected by plm_reactor_stats
  DevBuf stage[8]; // host-pointer staging: rY rYsrc T rE rEsrc FC mask I_R
  std::mutex host_mu;  // the staging buffers are one set per reactor: host-pointer calls (CPU-built AMReX, possibly from
                       // inside an OpenMP tile loop, PeleLM.cpp:5725) are serialised
  DevBuf scratch[6];
  plm_stats last{};
  double pending_ms = 0.0;  // kernel time of completed-but-unreported calls
  // cached launch configuration
  plm::LaunchCfg cfg{};
  int cfg_qmax = -1;
  int path = 0;  // 0: automatic, 1: warp-per-cell kernels, 2: lane = cell kernels (plm_reactor_set_path)
};

namespace {
bool use_lane(const plm_reactor* r)
{
  if (r->path == 1) return false;
  if (r->path == 2) return plm::lane_supported(r->mech.h, r->integ, r->opts);
  return plm::lane_auto(r->mech.h, r->integ, r->opts);
}
}  // namespace

extern "C" {

const char* plm_last_error(void) { return g_err.c_str(); }

int plm_reactor_create(plm_reactor** out, const char* mechanism, const char* integrator)
{
  if (!out) return fail(PLM_ERR_ARG, "null output handle");
  *out = nullptr;
  const plm_mech_tables* t = find_mech(mechanism);
  if (!t) return fail(PLM_ERR_MECH, std::string("unknown mechanism '") + (mechanism ? mechanism : "(null)") +
                                      "'; valid: LiDryer, drm19, dodecane_lu");
  int integ;
  if (integrator && !std::strcmp(integrator, "ReactorCvode")) integ = plm::kIntegCvode;
  else if (integrator && !std::strcmp(integrator, "ReactorRK64")) integ = plm::kIntegRK64;
  else  // Factory.H:134-150 aborts listing the valid keys
    return fail(PLM_ERR_ARG, std::string("unknown chem_integrator '") + (integrator ? integrator : "(null)") +
                               "'; valid: ReactorCvode, ReactorRK64");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(PLM_ERR_NO_DEVICE, "no CUDA device: libplmchem has no CPU path");
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp prop{};
  PLM_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major < 10) return fail(PLM_ERR_NO_DEVICE, "device is not sm_100-class; this library is built for sm_100a only");
  auto* r = new plm_reactor();
  r->tables = t;
  r->integ = integ;
  r->num_sms = prop.multiProcessorCount;
  if (plm::build_dev_mech(t, &r->mech) != 0) {
    delete r;
    return fail(PLM_ERR_CUDA, "failed to build/upload mechanism tables");
  }
  if (r->ctrl.ensure((kSlots + kStatWords) * sizeof(unsigned long long)) != 0 ||
      cudaMemset(r->ctrl.p, 0, (kSlots + kStatWords) * sizeof(unsigned long long)) != cudaSuccess ||
      cudaMallocHost(reinterpret_cast<void**>(&r->host_stats), (kStatWords + 1) * sizeof(unsigned long long)) != cudaSuccess) {
    plm::free_dev_mech(&r->mech);
    r->ctrl.release();
    delete r;
    return fail(PLM_ERR_CUDA, "cudaMalloc failed");
  }
  for (auto& sl : r->slot) {
    cudaEventCreate(&sl.ev0);
    cudaEventCreate(&sl.ev1);
  }
  r->host_stats[kStatWords] = 0;
  *out = r;
  return PLM_OK;
}

int plm_num_species(const plm_reactor* r) { return r ? r->tables->ns : 0; }

const char* plm_species_name(const plm_reactor* r, int k)
{
  if (!r || k < 0 || k >= r->tables->ns) return "";
  return r->tables->species[k];
}

int plm_reactor_init(plm_reactor* r, int reactor_type, double rtol, double atol, int max_order, int max_nls_iters,
                     int clean_init_massfrac, const char* solve_type, int rk64_guess, int rk64_min, int rk64_max)
{
  if (!r) return fail(PLM_ERR_ARG, "null reactor");
  if (reactor_type != 1 && reactor_type != 2)  // ReactorTypes.H:15-21: 1 = internal energy (UV), 2 = enthalpy (HP; PeleLM)
    return fail(PLM_ERR_ARG, "reactor_type must be 1 (internal energy) or 2 (enthalpy)");
  if (!(rtol > 0) || !(atol > 0)) return fail(PLM_ERR_ARG, "rtol/atol must be positive");
  if (max_order < 1 || max_order > 5) return fail(PLM_ERR_ARG, "cvode.max_order must be in 1..5");
  if (max_nls_iters < 1) return fail(PLM_ERR_ARG, "ode.max_nls_iters must be >= 1");
  int ajac = 1;
  if (solve_type && !std::strcmp(solve_type, "denseAJ_direct")) ajac = 1;
  else if (solve_type && !std::strcmp(solve_type, "dense_direct")) ajac = 0;
  else if (r->integ == plm::kIntegCvode)
    return fail(PLM_ERR_ARG, "cvode.solve_type must be denseAJ_direct or dense_direct (others are out of scope)");
  if (rk64_guess < 1 || rk64_min < 1 || rk64_max < 1) return fail(PLM_ERR_ARG, "rk64 substep counts must be >= 1");
  plm::ReactOpts& o = r->opts;
  const bool had_typ = o.use_typ;
  o.rtol = rtol;
  o.atol = atol;
  o.qmax = max_order;
  o.maxcor = max_nls_iters;
  o.ajac = ajac;
  o.clean_init_massfrac = clean_init_massfrac;
  o.rk64_guess = rk64_guess;
  o.rk64_min = rk64_min;
  o.rk64_max = rk64_max;
  o.use_typ = had_typ;
  if (plm::select_reactor_type(&r->mech, reactor_type) != 0) return fail(PLM_ERR_CUDA, "could not update the mechanism descriptor");
  r->reactor_type = reactor_type;
  r->inited = true;
  r->cfg_qmax = -1;
  return PLM_OK;
}

int plm_reactor_set_typ_vals(plm_reactor* r, const double* typ, int n)
{
  if (!r || !typ) return fail(PLM_ERR_ARG, "null argument");
  if (n != r->tables->ns + 1) return fail(PLM_ERR_ARG, "typical values must have NUM_SPECIES+1 entries");
  for (int i = 0; i < n; ++i) r->opts.typ_vals[i] = typ[i];
  r->opts.use_typ = typ[0] > 0.0;  // ReactorBase.cpp:105
  return PLM_OK;
}

}  // extern "C"

namespace {

int configure(plm_reactor* r)
{
  if (r->cfg_qmax == r->opts.qmax) return PLM_OK;
  const plm::DevMech& m = r->mech.h;
  const int wpb = plm::react_threads(m, r->integ, r->opts.qmax) / 32;
  if (wpb <= 0) return fail(PLM_ERR_CUDA, "no kernel for this mechanism");
  size_t smem = plm::react_smem_bytes(m, r->integ, r->opts.qmax, wpb);
  if (smem == 0 || smem > 227 * 1024) return fail(PLM_ERR_CUDA, "no launch configuration fits in shared memory");
  if (const char* lim = std::getenv("PLM_OCC_LIMIT")) {  // development: cap resident CTAs per SM by padding shared memory
    const int k = std::atoi(lim);
    if (k > 0) smem = std::max(smem, size_t(227 * 1024 / (k + 1) + 1024));
  }
  const int occ = plm::react_occupancy(m, r->integ, r->opts.qmax, wpb * 32, smem);
  if (occ <= 0) return fail(PLM_ERR_CUDA, "kernel cannot be resident (occupancy query failed)");
  r->cfg.blocks = occ * r->num_sms;  // persistent grid: every resident CTA slot of every SM
  r->cfg.threads = wpb * 32;
  r->cfg.smem_bytes = smem;
  r->cfg_qmax = r->opts.qmax;
  return PLM_OK;
}

// Enqueue one react call on `stream`. Never blocks the host: statistics accumulate on the device and are
// fetched by collect_stats(); each call owns a slot (queue counter, status array, saved-Jacobian scratch) so that
// calls issued on different streams may overlap.
int run_react(plm_reactor* r, plm::BoxArgs& a, cudaStream_t stream)
{
  const bool lane = use_lane(r);
  if (r->path == 2 && !lane) return fail(PLM_ERR_ARG, "the lane = cell kernels do not cover this mechanism / integrator / option set");
  int rc = lane ? PLM_OK : configure(r);
  if (rc != PLM_OK) return rc;
  const plm::DevMech& m = r->mech.h;
  const int si = r->next_slot;
  r->next_slot = (r->next_slot + 1) % kSlots;
  CallSlot& sl = r->slot[si];
  if (sl.used) {  // slot still owned by an earlier call: its kernel must have finished before the counter is reused
    PLM_CUDA(cudaEventSynchronize(sl.ev1));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sl.ev0, sl.ev1) == cudaSuccess) r->pending_ms += ms;
    sl.used = false;
  }
  if (sl.status.ensure(std::max<int64_t>(a.ncells, 1) * sizeof(int)) != 0) return fail(PLM_ERR_CUDA, "cudaMalloc failed");
  if (a.flat && !a.FC.p) {  // per-call scratch: calls in flight on other streams own other slots
    if (sl.fcflat.ensure(std::max<int64_t>(a.ncells, 1) * sizeof(double)) != 0) return fail(PLM_ERR_CUDA, "cudaMalloc failed");
    a.FC.p = static_cast<double*>(sl.fcflat.p);
  }
  auto* ctrl = static_cast<unsigned long long*>(r->ctrl.p);
  if (a.nboxes > 0) {  // the boxes of a multi-box call: host vectors (a.tab / a.off) -> this slot's device table
    const size_t tb = size_t(a.nboxes) * sizeof(plm::BoxArgs), ob = size_t(a.nboxes + 1) * sizeof(int64_t);
    if (sl.boxtab.ensure(tb + ob) != 0) return fail(PLM_ERR_CUDA, "cudaMalloc failed");
    PLM_CUDA(cudaMemcpyAsync(sl.boxtab.p, a.tab, tb, cudaMemcpyHostToDevice, stream));
    PLM_CUDA(cudaMemcpyAsync(static_cast<char*>(sl.boxtab.p) + tb, a.off, ob, cudaMemcpyHostToDevice, stream));
    PLM_CUDA(cudaStreamSynchronize(stream));  // the host vectors are the caller's locals (pageable memory: the copies are staged)
    a.tab = static_cast<const plm::BoxArgs*>(sl.boxtab.p);
    a.off = reinterpret_cast<const int64_t*>(static_cast<char*>(sl.boxtab.p) + tb);
  }
  if (lane) {  // one thread per cell, one persistent CTA per SM: every lane pulls its own cells from the queue
    const int threads = plm::lane_threads(m, r->integ, r->opts.qmax);
    const int blocks = int(std::max<int64_t>(1, std::min<int64_t>(r->num_sms, (a.ncells + threads - 1) / threads)));
    if (sl.lanews.ensure(plm::lane_ws_bytes(m, r->integ, r->opts.qmax, blocks)) != 0) return fail(PLM_ERR_CUDA, "cudaMalloc failed");
    PLM_CUDA(cudaMemsetAsync(ctrl + si, 0, sizeof(unsigned long long), stream));
    PLM_CUDA(cudaEventRecord(sl.ev0, stream));
    PLM_CUDA(plm::launch_lane_react(m, r->opts, a, r->integ, blocks, ctrl + si, ctrl + kSlots, static_cast<int*>(sl.status.p),
                                    static_cast<double*>(sl.lanews.p), stream));
    PLM_CUDA(cudaEventRecord(sl.ev1, stream));
    PLM_CUDA(cudaMemcpyAsync(r->host_stats + kStatWords, ctrl + kSlots + 7, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    sl.used = true;
    sl.ncells = a.ncells;
    r->last_slot = si;
    return PLM_OK;
  }
  plm::LaunchCfg cfg = r->cfg;
  const int64_t warps_needed = (a.ncells + plm::kCellsPerGrab - 1) / plm::kCellsPerGrab;
  const int wpb = cfg.threads / 32;
  cfg.blocks = int(std::max<int64_t>(1, std::min<int64_t>(cfg.blocks, (warps_needed + wpb - 1) / wpb)));
  {  // saved Jacobians: one (N+1) x LD matrix per resident warp (L2 resident; read/written only at linear-solver setups)
    const size_t nw = size_t(r->cfg.blocks) * wpb;
    if (sl.jglobal.ensure(nw * plm::react_jsave_doubles(m) * sizeof(double)) != 0) return fail(PLM_ERR_CUDA, "cudaMalloc failed");
  }
  PLM_CUDA(cudaMemsetAsync(ctrl + si, 0, sizeof(unsigned long long), stream));
  PLM_CUDA(cudaEventRecord(sl.ev0, stream));
  PLM_CUDA(plm::launch_react(m, r->mech.d, r->opts, a, cfg, r->integ, ctrl + si, ctrl + kSlots,
                             static_cast<int*>(sl.status.p), static_cast<double*>(sl.jglobal.p), stream));
  PLM_CUDA(cudaEventRecord(sl.ev1, stream));
  PLM_CUDA(cudaMemcpyAsync(r->host_stats + kStatWords, ctrl + kSlots + 7, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
  sl.used = true;
  sl.ncells = a.ncells;
  r->last_slot = si;
  return PLM_OK;
}

// Wait for every outstanding call, then move the cumulative device counters into r->last and zero them.
int collect_stats(plm_reactor* r)
{
  for (auto& sl : r->slot)
    if (sl.used) {
      PLM_CUDA(cudaEventSynchronize(sl.ev1));
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, sl.ev0, sl.ev1) == cudaSuccess) r->pending_ms += ms;
      sl.used = false;
    }
  auto* g_dev = static_cast<unsigned long long*>(r->ctrl.p) + kSlots;
  PLM_CUDA(cudaMemcpy(r->host_stats, g_dev, kStatWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  PLM_CUDA(cudaMemset(g_dev, 0, kStatWords * sizeof(unsigned long long)));
  const unsigned long long* g = r->host_stats;
  plm_stats& s = r->last;
  s.nst = int64_t(g[0]); s.nfe = int64_t(g[1]); s.nje = int64_t(g[2]); s.nsetups = int64_t(g[3]);
  s.nni = int64_t(g[4]); s.ncfn = int64_t(g[5]); s.netf = int64_t(g[6]); s.nfailed = int64_t(g[7]);
  s.nskipped = int64_t(g[8]); s.ncells = int64_t(g[9]);
  s.max_fc = int64_t(g[10]);
  s.min_fc = g[9] ? int64_t(0xffffffffull - g[11]) : 0;  // the kernel keeps max(0xffffffff - fc)
  s.kernel_ms = r->pending_ms;
  r->pending_ms = 0.0;
  r->failed_total += s.nfailed;
  r->host_stats[kStatWords] = 0;  // the device counter was just cleared
  return PLM_OK;
}

// copy a host FAB view to the device; returns device pointer in *dp
template <class T>
int stage_in(DevBuf& buf, const T* host, size_t count, T** dp, cudaStream_t st, bool copy)
{
  if (buf.ensure(std::max<size_t>(count, 1) * sizeof(T)) != 0) return fail(PLM_ERR_CUDA, "cudaMalloc failed");
  *dp = static_cast<T*>(buf.p);
  if (copy && count) PLM_CUDA(cudaMemcpyAsync(*dp, host, count * sizeof(T), cudaMemcpyHostToDevice, st));
  return PLM_OK;
}

// doubles of the first `ncomp` components of a FAB view: what the kernel owns of it. PeleLM passes overlapping views of one
// FAB (rhoY = STemp comps 0.., rhoH = comp N, T = comp N+1; Array4::ncomp counts the REMAINING components of the FAB), so
// staging or copying back more than the owned components would let one view overwrite another's result.
size_t a4_count(const plm_a4d& a, int ncomp) { return size_t(a.nstride) * size_t(std::max(ncomp, 1)); }

// 1 device, 0 host, -1 mixed
int pointer_kind(std::initializer_list<const void*> ps)
{
  int kind = -2;
  for (const void* p : ps) {
    if (!p) continue;
    const int k = is_device_ptr(p) ? 1 : 0;
    if (kind == -2) kind = k;
    else if (kind != k) return -1;
  }
  return kind == -2 ? 0 : kind;
}

}  // namespace

extern "C" {

}  // extern "C"

namespace {
// react over a box; mks != 0: the fused PeleLM::advance_chemistry body (MKS arrays, optional I_R output)
int react_box_impl(plm_reactor* r, const int lo[3], const int hi[3], plm_a4d rY, plm_a4d rYsrc, plm_a4d T,
                   plm_a4d rEner, plm_a4d rEner_src, plm_a4d FC, plm_a4i mask, const double* dt_react,
                   const double* time, void* cuda_stream, int mks, plm_a4d IR)
{
  NvtxRange nvtx(mks ? "PLM::advance_chemistry()" : "Pele::ReactorCvode::react()");
  if (!r || !lo || !hi || !dt_react || !time) return fail(PLM_ERR_ARG, "null argument");
  if (!r->inited) return fail(PLM_ERR_ARG, "plm_reactor_init has not been called");
  if (!rY.p || !rYsrc.p || !T.p || !rEner.p || !rEner_src.p || !FC.p) return fail(PLM_ERR_ARG, "null array");
  const int ns = r->tables->ns;
  if (rY.ncomp < ns || rYsrc.ncomp < ns) return fail(PLM_ERR_ARG, "rY/rYsrc need NUM_SPECIES components");
  for (int d = 0; d < 3; ++d) {
    if (hi[d] < lo[d]) return PLM_OK;  // empty box: nothing to do (amrex::ParallelFor over an empty box)
    const plm_a4d* arrs[6] = {&rY, &rYsrc, &T, &rEner, &rEner_src, &FC};
    for (auto* a : arrs)
      if (lo[d] < a->lo[d] || hi[d] > a->hi[d]) return fail(PLM_ERR_ARG, "box is not contained in an array's bounds");
    if (mask.p && (lo[d] < mask.lo[d] || hi[d] > mask.hi[d])) return fail(PLM_ERR_ARG, "box is not contained in mask bounds");
  }
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  plm::BoxArgs a{};
  for (int d = 0; d < 3; ++d) {
    a.lo[d] = lo[d];
    a.hi[d] = hi[d];
  }
  a.ncells = int64_t(hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1);
  a.flat = 0;
  a.dt = *dt_react;
  a.time = *time;
  const int kind = pointer_kind({rY.p, rYsrc.p, T.p, rEner.p, rEner_src.p, FC.p, mask.p, IR.p});
  if (kind < 0) return fail(PLM_ERR_ARG, "the arrays of one call must be all host or all device memory");
  const bool dev = kind == 1;
  a.rY = rY; a.rYsrc = rYsrc; a.T = T; a.rE = rEner; a.rEsrc = rEner_src; a.FC = FC; a.mask = mask;
  a.mks = mks;
  a.IR = IR;
  a.dt_inv = 1.0 / a.dt;
  if (IR.p) {
    if (IR.ncomp < ns) return fail(PLM_ERR_ARG, "I_R needs NUM_SPECIES components");
    for (int d = 0; d < 3; ++d)
      if (lo[d] < IR.lo[d] || hi[d] > IR.hi[d]) return fail(PLM_ERR_ARG, "box is not contained in I_R bounds");
  }
  std::unique_lock<std::mutex> host_lock;
  if (!dev) {
    host_lock = std::unique_lock<std::mutex>(r->host_mu);
    int rc;
    if (IR.p && (rc = stage_in(r->stage[7], IR.p, a4_count(IR, ns), &a.IR.p, st, true))) return rc;
    if ((rc = stage_in(r->stage[0], rY.p, a4_count(rY, ns), &a.rY.p, st, true))) return rc;
    if ((rc = stage_in(r->stage[1], rYsrc.p, a4_count(rYsrc, ns), &a.rYsrc.p, st, true))) return rc;
    if ((rc = stage_in(r->stage[2], T.p, a4_count(T, 1), &a.T.p, st, true))) return rc;
    if ((rc = stage_in(r->stage[3], rEner.p, a4_count(rEner, 1), &a.rE.p, st, true))) return rc;
    if ((rc = stage_in(r->stage[4], rEner_src.p, a4_count(rEner_src, 1), &a.rEsrc.p, st, true))) return rc;
    if ((rc = stage_in(r->stage[5], FC.p, a4_count(FC, 1), &a.FC.p, st, true))) return rc;
    if (mask.p) {
      if ((rc = stage_in(r->stage[6], mask.p, size_t(mask.nstride), &a.mask.p, st, true))) return rc;
    }
  }
  int rc = run_react(r, a, st);
  if (rc != PLM_OK) return rc;
  if (!dev) {
    PLM_CUDA(cudaMemcpyAsync(rY.p, a.rY.p, a4_count(rY, ns) * sizeof(double), cudaMemcpyDeviceToHost, st));
    PLM_CUDA(cudaMemcpyAsync(T.p, a.T.p, a4_count(T, 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
    PLM_CUDA(cudaMemcpyAsync(rEner.p, a.rE.p, a4_count(rEner, 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
    PLM_CUDA(cudaMemcpyAsync(FC.p, a.FC.p, a4_count(FC, 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (IR.p) PLM_CUDA(cudaMemcpyAsync(IR.p, a.IR.p, a4_count(IR, ns) * sizeof(double), cudaMemcpyDeviceToHost, st));
    PLM_CUDA(cudaStreamSynchronize(st));
  }
  return PLM_OK;
}
}  // namespace

extern "C" {

int plm_react_box(plm_reactor* r, const int lo[3], const int hi[3], plm_a4d rY, plm_a4d rYsrc, plm_a4d T,
                  plm_a4d rEner, plm_a4d rEner_src, plm_a4d FC, plm_a4i mask, double* dt_react, double* time,
                  void* cuda_stream)
{
  return react_box_impl(r, lo, hi, rY, rYsrc, T, rEner, rEner_src, FC, mask, dt_react, time, cuda_stream, 0, plm_a4d{});
}

int plm_react_boxes(plm_reactor* r, int nboxes, const plm_box* boxes, double dt, double time, int mks, void* cuda_stream)
{
  NvtxRange nvtx(mks ? "PLM::advance_chemistry()" : "Pele::ReactorCvode::react()");
  if (!r || (nboxes > 0 && !boxes)) return fail(PLM_ERR_ARG, "null argument");
  if (!r->inited) return fail(PLM_ERR_ARG, "plm_reactor_init has not been called");
  if (nboxes <= 0) return PLM_OK;
  const double t0 = mks ? 0.0 : time;  // PeleLM.cpp:5751
  for (int b = 0; b < nboxes; ++b)
    if (!is_device_ptr(boxes[b].rY.p)) return fail(PLM_ERR_ARG, "plm_react_boxes takes device arrays");
  // One launch over the cells of every box (both kernel families: the work queue runs over all of them, so its tail is paid
  // once per call).  PLM_BOXES_PER_LAUNCH=1 (development A/B) and the two-cells-per-warp ReactorRK64 experiment keep the
  // warp-per-cell kernels on one launch per box.
  static const bool per_box_env = [] { const char* e = std::getenv("PLM_BOXES_PER_LAUNCH"); return e && e[0] == '1'; }();
  static const bool rk_dual_env = [] { const char* e = std::getenv("PLM_RK64_DUAL"); return e && e[0] == '1'; }();
  if (!use_lane(r) && (per_box_env || (rk_dual_env && r->integ == plm::kIntegRK64))) {
    for (int b = 0; b < nboxes; ++b) {
      const plm_box& x = boxes[b];
      const int rc = react_box_impl(r, x.lo, x.hi, x.rY, x.rYsrc, x.T, x.rEner, x.rEner_src, x.FC, x.mask, &dt, &t0, cuda_stream,
                                    mks, mks ? x.I_R : plm_a4d{});
      if (rc != PLM_OK) return rc;
    }
    return PLM_OK;
  }
  const int ns = r->tables->ns;
  std::vector<plm::BoxArgs> tab;
  std::vector<int64_t> off(1, 0);
  tab.reserve(nboxes);
  for (int b = 0; b < nboxes; ++b) {
    const plm_box& x = boxes[b];
    if (x.hi[0] < x.lo[0] || x.hi[1] < x.lo[1] || x.hi[2] < x.lo[2]) continue;  // empty box
    if (!x.rY.p || !x.rYsrc.p || !x.T.p || !x.rEner.p || !x.rEner_src.p || !x.FC.p) return fail(PLM_ERR_ARG, "null array");
    if (x.rY.ncomp < ns || x.rYsrc.ncomp < ns) return fail(PLM_ERR_ARG, "rY/rYsrc need NUM_SPECIES components");
    const plm_a4d* arrs[6] = {&x.rY, &x.rYsrc, &x.T, &x.rEner, &x.rEner_src, &x.FC};
    for (int d = 0; d < 3; ++d) {
      for (auto* ar : arrs)
        if (x.lo[d] < ar->lo[d] || x.hi[d] > ar->hi[d]) return fail(PLM_ERR_ARG, "box is not contained in an array's bounds");
      if (x.mask.p && (x.lo[d] < x.mask.lo[d] || x.hi[d] > x.mask.hi[d])) return fail(PLM_ERR_ARG, "box is not contained in mask bounds");
      if (mks && x.I_R.p && (x.lo[d] < x.I_R.lo[d] || x.hi[d] > x.I_R.hi[d])) return fail(PLM_ERR_ARG, "box is not contained in I_R bounds");
    }
    plm::BoxArgs e{};
    for (int d = 0; d < 3; ++d) {
      e.lo[d] = x.lo[d];
      e.hi[d] = x.hi[d];
    }
    e.ncells = int64_t(x.hi[0] - x.lo[0] + 1) * (x.hi[1] - x.lo[1] + 1) * (x.hi[2] - x.lo[2] + 1);
    e.rY = x.rY; e.rYsrc = x.rYsrc; e.T = x.T; e.rE = x.rEner; e.rEsrc = x.rEner_src; e.FC = x.FC; e.mask = x.mask;
    e.IR = mks ? x.I_R : plm_a4d{};
    tab.push_back(e);
    off.push_back(off.back() + e.ncells);
  }
  if (tab.empty()) return PLM_OK;
  plm::BoxArgs a = tab[0];
  a.ncells = off.back();
  a.flat = 0;
  a.dt = dt;
  a.time = t0;
  a.mks = mks;
  a.dt_inv = 1.0 / dt;
  a.tab = tab.data();
  a.off = off.data();
  a.nboxes = int(tab.size());
  return run_react(r, a, static_cast<cudaStream_t>(cuda_stream));
}

int plm_advance_chemistry_box(plm_reactor* r, const int lo[3], const int hi[3], plm_a4d rhoY, plm_a4d rhoH, plm_a4d T,
                              plm_a4d frc_rhoY, plm_a4d frc_rhoH, plm_a4d I_R, plm_a4d FC, plm_a4i mask, double dt,
                              void* cuda_stream)
{
  const double time = 0.0;  // PeleLM.cpp:5751
  return react_box_impl(r, lo, hi, rhoY, frc_rhoY, T, rhoH, frc_rhoH, FC, mask, &dt, &time, cuda_stream, 1, I_R);
}

int plm_reaction_rate_rhoY(plm_reactor* r, const int lo[3], const int hi[3], plm_a4d rhoY, plm_a4d rhoH, plm_a4d T,
                           plm_a4i mask, plm_a4d rhoYdot, void* cuda_stream)
{
  NvtxRange nvtx("PLM::reactionRateRhoY()");
  if (!r || !lo || !hi) return fail(PLM_ERR_ARG, "null argument");
  if (!rhoY.p || !rhoH.p || !T.p || !rhoYdot.p) return fail(PLM_ERR_ARG, "null array");
  if (r->reactor_type != 2) return fail(PLM_ERR_ARG, "reactionRateRhoY works on enthalpy (PeleLM_K.H:84-86): reactor_type must be 2");
  const int ns = r->tables->ns;
  if (rhoY.ncomp < ns || rhoYdot.ncomp < ns) return fail(PLM_ERR_ARG, "rhoY/rhoYdot need NUM_SPECIES components");
  for (int d = 0; d < 3; ++d) {
    if (hi[d] < lo[d]) return PLM_OK;
    const plm_a4d* arrs[4] = {&rhoY, &rhoH, &T, &rhoYdot};
    for (auto* a : arrs)
      if (lo[d] < a->lo[d] || hi[d] > a->hi[d]) return fail(PLM_ERR_ARG, "box is not contained in an array's bounds");
    if (mask.p && (lo[d] < mask.lo[d] || hi[d] > mask.hi[d])) return fail(PLM_ERR_ARG, "box is not contained in mask bounds");
  }
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  plm::BoxArgs a{};
  for (int d = 0; d < 3; ++d) {
    a.lo[d] = lo[d];
    a.hi[d] = hi[d];
  }
  a.ncells = int64_t(hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1);
  a.rY = rhoY; a.rE = rhoH; a.T = T; a.mask = mask;
  plm_a4d out = rhoYdot;
  const int kind = pointer_kind({rhoY.p, rhoH.p, T.p, mask.p, rhoYdot.p});
  if (kind < 0) return fail(PLM_ERR_ARG, "the arrays of one call must be all host or all device memory");
  const bool dev = kind == 1;
  std::unique_lock<std::mutex> host_lock;
  if (!dev) {
    host_lock = std::unique_lock<std::mutex>(r->host_mu);
    int rc;
    if ((rc = stage_in(r->stage[0], rhoY.p, a4_count(rhoY, ns), &a.rY.p, st, true))) return rc;
    if ((rc = stage_in(r->stage[2], T.p, a4_count(T, 1), &a.T.p, st, true))) return rc;
    if ((rc = stage_in(r->stage[3], rhoH.p, a4_count(rhoH, 1), &a.rE.p, st, true))) return rc;
    if ((rc = stage_in(r->stage[1], rhoYdot.p, a4_count(rhoYdot, ns), &out.p, st, true))) return rc;
    if (mask.p) {
      if ((rc = stage_in(r->stage[6], mask.p, size_t(mask.nstride), &a.mask.p, st, true))) return rc;
    }
  }
  PLM_CUDA(plm::launch_reaction_rate(r->mech.h, r->mech.d, a, out, st));
  if (!dev) {
    PLM_CUDA(cudaMemcpyAsync(rhoYdot.p, out.p, a4_count(rhoYdot, ns) * sizeof(double), cudaMemcpyDeviceToHost, st));
    PLM_CUDA(cudaStreamSynchronize(st));
  }
  return PLM_OK;
}

int plm_react_flat(plm_reactor* r, double* rY, double* rYsrc, double* rX, double* rX_src, double* dt_react,
                   double* time, int ncells, void* cuda_stream)
{
  NvtxRange nvtx("Pele::ReactorCvode::react():flat");
  if (!r || !rY || !rYsrc || !rX || !rX_src || !dt_react || !time) return fail(PLM_ERR_ARG, "null argument");
  if (!r->inited) return fail(PLM_ERR_ARG, "plm_reactor_init has not been called");
  if (ncells < 0) return fail(PLM_ERR_ARG, "negative ncells");
  if (ncells == 0) return PLM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int ns = r->tables->ns, n = ns + 1;
  plm::BoxArgs a{};
  a.ncells = ncells;
  a.flat = 1;
  a.hi[0] = ncells - 1;
  a.dt = *dt_react;
  a.time = *time;
  const int kind = pointer_kind({rY, rYsrc, rX, rX_src});
  if (kind < 0) return fail(PLM_ERR_ARG, "the arrays of one call must be all host or all device memory");
  const bool dev = kind == 1;
  double *dY = rY, *dS = rYsrc, *dX = rX, *dXs = rX_src, *dFC = nullptr;
  int rc;
  std::unique_lock<std::mutex> host_lock;
  if (!dev) {
    host_lock = std::unique_lock<std::mutex>(r->host_mu);
    if ((rc = stage_in(r->stage[0], rY, size_t(ncells) * n, &dY, st, true))) return rc;
    if ((rc = stage_in(r->stage[1], rYsrc, size_t(ncells) * ns, &dS, st, true))) return rc;
    if ((rc = stage_in(r->stage[3], rX, size_t(ncells), &dX, st, true))) return rc;
    if ((rc = stage_in(r->stage[4], rX_src, size_t(ncells), &dXs, st, true))) return rc;
  }
  a.rY.p = dY; a.rYsrc.p = dS; a.T.p = dY; a.rE.p = dX; a.rEsrc.p = dXs; a.FC.p = dFC; a.mask.p = nullptr;
  rc = run_react(r, a, st);
  if (rc != PLM_OK) return rc;
  if (!dev) {
    PLM_CUDA(cudaMemcpyAsync(rY, dY, size_t(ncells) * n * sizeof(double), cudaMemcpyDeviceToHost, st));
    PLM_CUDA(cudaMemcpyAsync(rX, dX, size_t(ncells) * sizeof(double), cudaMemcpyDeviceToHost, st));
    PLM_CUDA(cudaStreamSynchronize(st));
  }
  return PLM_OK;
}

int plm_reactor_failed_cells(plm_reactor* r, long long* nfailed)
{
  if (!r || !nfailed) return fail(PLM_ERR_ARG, "null argument");
  *nfailed = r->failed_total + static_cast<long long>(r->host_stats[kStatWords]);
  return PLM_OK;
}

int plm_reactor_uses_lane(const plm_reactor* r) { return (r && r->inited && use_lane(r)) ? 1 : 0; }

int plm_reactor_set_path(plm_reactor* r, int path)
{
  if (!r) return fail(PLM_ERR_ARG, "null reactor");
  if (path < 0 || path > 2) return fail(PLM_ERR_ARG, "path must be 0 (automatic), 1 (warp per cell) or 2 (lane = cell)");
  r->path = path;
  return PLM_OK;
}

int plm_reactor_stats(plm_reactor* r, plm_stats* out)
{
  if (!r || !out) return fail(PLM_ERR_ARG, "null argument");
  const int rc = collect_stats(r);
  if (rc != PLM_OK) return rc;
  *out = r->last;
  return PLM_OK;
}

int plm_reactor_cell_status(plm_reactor* r, int* dst, int64_t ncells)
{
  if (!r || !dst) return fail(PLM_ERR_ARG, "null argument");
  if (r->last_slot < 0 || ncells > r->slot[r->last_slot].ncells)
    return fail(PLM_ERR_ARG, "more cells requested than the last call processed");
  CallSlot& sl = r->slot[r->last_slot];
  if (sl.used) PLM_CUDA(cudaEventSynchronize(sl.ev1));
  PLM_CUDA(cudaMemcpy(dst, sl.status.p, size_t(ncells) * sizeof(int),
                      is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost));
  return PLM_OK;
}

void plm_reactor_close(plm_reactor* r)
{
  if (!r) return;
  plm::free_dev_mech(&r->mech);
  cudaDeviceSynchronize();
  r->ctrl.release();
  for (auto& sl : r->slot) {
    sl.status.release();
    sl.jglobal.release();
    sl.lanews.release();
    sl.fcflat.release();
    sl.boxtab.release();
    if (sl.ev0) cudaEventDestroy(sl.ev0);
    if (sl.ev1) cudaEventDestroy(sl.ev1);
  }
  if (r->host_stats) cudaFreeHost(r->host_stats);
  for (auto& b : r->stage) b.release();
  for (auto& b : r->scratch) b.release();
  delete r;
}

int plm_eval_rhs(plm_reactor* r, int ncells, double t, const double* y, const double* rhoe_init,
                 const double* rhoesrc_ext, const double* rYsrc, double* ydot, void* cuda_stream)
{
  if (!r || !y || !rhoe_init || !rhoesrc_ext || !rYsrc || !ydot) return fail(PLM_ERR_ARG, "null argument");
  if (ncells <= 0) return PLM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int ns = r->tables->ns, n = ns + 1;
  const bool dev = is_device_ptr(y);
  double *dy = const_cast<double*>(y), *de = const_cast<double*>(rhoe_init), *des = const_cast<double*>(rhoesrc_ext),
         *ds = const_cast<double*>(rYsrc), *dout = ydot;
  int rc;
  if (!dev) {
    if ((rc = stage_in(r->scratch[0], y, size_t(ncells) * n, &dy, st, true))) return rc;
    if ((rc = stage_in(r->scratch[1], rhoe_init, size_t(ncells), &de, st, true))) return rc;
    if ((rc = stage_in(r->scratch[2], rhoesrc_ext, size_t(ncells), &des, st, true))) return rc;
    if ((rc = stage_in(r->scratch[3], rYsrc, size_t(ncells) * ns, &ds, st, true))) return rc;
    if ((rc = stage_in(r->scratch[4], static_cast<double*>(nullptr), size_t(ncells) * n, &dout, st, false))) return rc;
  }
  if (r->path == 2 ? plm::lane_supported(r->mech.h, plm::kIntegRK64, r->opts) : (r->path == 0 && plm::lane_auto(r->mech.h, plm::kIntegRK64, r->opts)))
    PLM_CUDA(plm::launch_lane_rhs(r->mech.h, ncells, t, dy, de, des, ds, dout, st));
  else
    PLM_CUDA(plm::launch_rhs(r->mech.h, r->mech.d, ncells, t, dy, de, des, ds, dout, st));
  if (!dev) PLM_CUDA(cudaMemcpyAsync(ydot, dout, size_t(ncells) * n * sizeof(double), cudaMemcpyDeviceToHost, st));
  PLM_CUDA(cudaStreamSynchronize(st));
  return PLM_OK;
}

int plm_eval_jac(plm_reactor* r, int ncells, const double* y, double* J, void* cuda_stream)
{
  if (!r || !y || !J) return fail(PLM_ERR_ARG, "null argument");
  if (ncells <= 0) return PLM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int n = r->tables->ns + 1;
  const bool dev = is_device_ptr(y);
  double *dy = const_cast<double*>(y), *dJ = J;
  int rc;
  if (!dev) {
    if ((rc = stage_in(r->scratch[0], y, size_t(ncells) * n, &dy, st, true))) return rc;
    if ((rc = stage_in(r->scratch[5], static_cast<double*>(nullptr), size_t(ncells) * n * n, &dJ, st, false))) return rc;
  }
  if (r->mech.h.eint == 0 &&
      (r->path == 2 ? plm::lane_supported(r->mech.h, plm::kIntegRK64, r->opts) : (r->path == 0 && plm::lane_auto(r->mech.h, plm::kIntegRK64, r->opts))))
    PLM_CUDA(plm::launch_lane_jac(r->mech.h, ncells, dy, dJ, st));
  else
    PLM_CUDA(plm::launch_jac(r->mech.h, r->mech.d, ncells, dy, dJ, st));
  if (!dev) PLM_CUDA(cudaMemcpyAsync(J, dJ, size_t(ncells) * n * n * sizeof(double), cudaMemcpyDeviceToHost, st));
  PLM_CUDA(cudaStreamSynchronize(st));
  return PLM_OK;
}

int plm_measure_fp64_peak(double* tflops, void* cuda_stream)
{
  if (!tflops) return fail(PLM_ERR_ARG, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(PLM_ERR_NO_DEVICE, "no CUDA device");
  }
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  PLM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8, iters = 4096;
  double* scratch = nullptr;
  PLM_CUDA(cudaMalloc(&scratch, size_t(blocks) * 256 * sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, st);
    cudaError_t e = plm::launch_dfma_peak(scratch, blocks, iters, st);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    if (e != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16.0 * iters * 256.0 * blocks;
    if (rep > 0 && ms > 0) best = std::max(best, flops / (ms * 1e-3) * 1e-12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(scratch);
  *tflops = best;
  return best > 0 ? PLM_OK : fail(PLM_ERR_CUDA, "DFMA probe failed");
}

int plm_box_costs(const double* FC, const int64_t* box_offsets, int nboxes, int64_t* costs, void* cuda_stream)
{
  if (!FC || !box_offsets || !costs || nboxes < 0) return fail(PLM_ERR_ARG, "bad argument");
  if (nboxes == 0) return PLM_OK;
  if (!is_device_ptr(FC) || !is_device_ptr(box_offsets) || !is_device_ptr(costs))
    return fail(PLM_ERR_ARG, "plm_box_costs takes device pointers");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  PLM_CUDA(plm::launch_box_costs(FC, box_offsets, nboxes, reinterpret_cast<long long*>(costs), st));
  return PLM_OK;
}

}  // extern "C"
