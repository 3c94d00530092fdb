// sclane_api.cu -- C ABI of libsclane_b200.so (include/sclane_b200.h): host orchestration of the
// per-lineage / per-batch kernel sequence, gene sharding over GPUs, and the small drop-in helpers.
// There is no CPU compute path in this file: every entry point that computes launches CUDA kernels
// and fails with SCLANE_ERR_CUDA when no device is usable.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <future>
#include <map>
#include <mutex>
#include <thread>

#include "sclane_comp_device.cuh"
#include "sclane_wg_device.cuh"
#include "sclane_exact_device.cuh"
#include "sclane_device.cuh"
#include "sclane_gee_device.cuh"
#include "sclane_k0_device.cuh"
#include "sclane_gee_host.hpp"
#include "sclane_host.hpp"

namespace sclane {
bool narrow_range_u16(const int32_t* src, size_t n, uint16_t* dst);      // sclane_narrow.cpp (AVX2 when the CPU has it)
}
using namespace sclane;

#ifndef SCLANE_BATCH_CAP_GB
#define SCLANE_BATCH_CAP_GB 24          // upper bound of the per-batch work buffers of one run_shard (the counts are extra)
#endif

namespace {

std::atomic<long long> g_launches{0};
std::mutex g_prof_mu;
sclane_profile g_prof;

void set_err(char* errbuf, size_t errlen, const char* fmt, ...) {
  if (!errbuf || errlen == 0) return;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(errbuf, errlen, fmt, ap);
  va_end(ap);
}

// SCLANE_TRACE=1: host-side timestamps (ms since the start of the call) of the phases of a batch call, on stderr
struct Trace {
  bool on;
  std::chrono::steady_clock::time_point t0;
  Trace() : on([] { const char* e = std::getenv("SCLANE_TRACE"); return e && e[0] == '1'; }()), t0(std::chrono::steady_clock::now()) {}
  void start() { t0 = std::chrono::steady_clock::now(); }
  void at(const char* what) const {
    if (on) fprintf(stderr, "[sclane trace] %8.3f ms  %s\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), what);
  }
};
Trace g_trace;

struct CudaFail {
  std::string msg;
};
struct PlanFail {
  std::string msg;
};

// The library is NOT re-entrant: the per-device buffer pools, their streams and the profile of the last call are process-wide.
// Every exported compute entry point (and sclane_release) therefore runs under this one lock -- concurrent callers are
// serialised, never corrupted -- and inside a try block: nothing is ever thrown across the C ABI.
std::mutex g_call_mu;
template <class F>
int32_t guarded(char* errbuf, size_t errlen, F&& body) {
  std::lock_guard<std::mutex> lk(g_call_mu);
  try {
    return body();
  } catch (const CudaFail& f) {
    set_err(errbuf, errlen, "%s", f.msg.c_str());
    return SCLANE_ERR_CUDA;
  } catch (const PlanFail& f) {
    set_err(errbuf, errlen, "%s", f.msg.c_str());
    return SCLANE_ERR_ARG;
  } catch (const std::bad_alloc&) {
    set_err(errbuf, errlen, "out of host memory");
    return SCLANE_ERR_NOMEM;
  } catch (const std::exception& e) {
    set_err(errbuf, errlen, "internal error: %s", e.what());
    return SCLANE_ERR_INTERNAL;
  } catch (...) {
    set_err(errbuf, errlen, "internal error (unknown exception)");
    return SCLANE_ERR_INTERNAL;
  }
}

// joins its threads on every exit path (an exception between creation and join would otherwise terminate the host)
struct ThreadGroup {
  std::vector<std::thread> th;
  template <class F, class... A> void run(F&& f, A&&... a) { th.emplace_back(std::forward<F>(f), std::forward<A>(a)...); }
  void join() { for (auto& t : th) if (t.joinable()) t.join(); th.clear(); }
  ~ThreadGroup() { join(); }
};
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char b_[512];                                                                                \
      snprintf(b_, sizeof(b_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      throw CudaFail{b_};                                                                          \
    }                                                                                              \
  } while (0)

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  void alloc(size_t count) {
    release();
    if (count == 0) return;
    CK(cudaMalloc(&p, count * sizeof(T)));
    n = count;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr; n = 0;
  }
  ~DevBuf() { release(); }
};

// Per-device cache of the large work buffers (and one stream): cudaMalloc / cudaFree cost milliseconds each,
// which is a visible share of a 2,000-gene call.  Buffers only grow; sclane_release() frees them.
struct DevicePool {
  struct Slot { void* p = nullptr; size_t bytes = 0; bool host = false; };
  std::map<std::string, Slot> slots;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;      // H2D of the next batch overlaps the kernels of the current one
  std::vector<cudaStream_t> extra_streams; // compute streams of the batch slots beyond the first (run_shard)
  void* get(const std::string& name, size_t bytes, bool pinned_host = false) {
    Slot& sl = slots[name];
    if (sl.bytes >= bytes && sl.p) return sl.p;
    if (sl.p) { if (sl.host) cudaFreeHost(sl.p); else cudaFree(sl.p); sl.p = nullptr; sl.bytes = 0; }
    if (bytes == 0) return nullptr;
    const size_t want = bytes + bytes / 8;           // a little slack so that slightly larger calls reuse the buffer
    if (pinned_host) { if (cudaHostAlloc(&sl.p, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); CK(cudaHostAlloc(&sl.p, bytes, cudaHostAllocDefault)); sl.bytes = bytes; sl.host = true; return sl.p; } }
    else { if (cudaMalloc(&sl.p, want) != cudaSuccess) { cudaGetLastError(); sl.p = nullptr; CK(cudaMalloc(&sl.p, bytes)); sl.bytes = bytes; sl.host = false; return sl.p; } }
    sl.bytes = want; sl.host = pinned_host;
    return sl.p;
  }
  size_t held() const { size_t t = 0; for (auto& kv : slots) if (!kv.second.host) t += kv.second.bytes; return t; }
  void release() {
    for (auto& kv : slots) if (kv.second.p) { if (kv.second.host) cudaFreeHost(kv.second.p); else cudaFree(kv.second.p); }
    slots.clear();
    if (stream) { cudaStreamDestroy(stream); stream = nullptr; }
    if (copy_stream) { cudaStreamDestroy(copy_stream); copy_stream = nullptr; }
    for (auto& e : extra_streams) if (e) cudaStreamDestroy(e);
    extra_streams.clear();
  }
};
std::mutex g_pool_mu;
std::map<int, DevicePool> g_pools;
DevicePool& pool_for(int device) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  return g_pools[device];
}

struct EventTimer {
  std::vector<cudaEvent_t> ev;
  std::vector<int> stage;     // stage of interval [i, i+1)
  cudaStream_t s;
  explicit EventTimer(cudaStream_t st) : s(st) {}
  void mark(int stage_of_next) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    CK(cudaEventRecord(e, s));
    ev.push_back(e);
    stage.push_back(stage_of_next);
  }
  // call after the stream is synchronised
  void collect(sclane_profile& p) {
    for (size_t i = 0; i + 1 < ev.size(); ++i) {
      float ms = 0.f;
      if (stage[i] >= 0 && cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess) p.ms[stage[i]] += ms;
    }
    for (auto e : ev) cudaEventDestroy(e);
    ev.clear(); stage.clear();
  }
  ~EventTimer() { for (auto e : ev) cudaEventDestroy(e); }
};

// Wait for everything queued on `s` WITHOUT spinning: cudaStreamSynchronize busy-waits on a core by default, and during a
// host-fed run every core is needed by the narrowing threads of the transfer.
void stream_wait_blocking(cudaStream_t s) {
  cudaEvent_t e;
  CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync));
  cudaError_t r = cudaEventRecord(e, s);
  if (r == cudaSuccess) r = cudaEventSynchronize(e);
  cudaEventDestroy(e);
  CK(r);
}


// ---- K0 on the device (sclane_k0_device.cuh) ---------------------------------------------------------------------------
// What the host decides up front from one scan of the pseudotime column (all of it depends on N and the range only).
struct K0Host {
  bool eligible = false;
  int N = 0;
  double xmin = 0.0, xmax = 0.0;
  long long key0 = 0;
  int passes = 0;                 // 8-bit radix passes that cover the key range
  int minspan = 0, maxspan = 0;
  K0Quant Q;
};

K0Host k0_prescan(const double* pt_col, int64_t n_cells, const sclane_opts& o, bool has_size_factor) {
  K0Host H;
  std::memset(&H.Q, 0, sizeof(H.Q));
  if (o.host_plan || !o.approx_knot || has_size_factor || o.mode != 0 || n_cells > (int64_t)(1 << 30)) return H;
  {
    static const bool env_host = [] { const char* e = std::getenv("SCLANE_HOST_PLAN"); return e && e[0] == '1'; }();
    if (env_host) return H;
  }
  int64_t N = 0;
  double lo = std::numeric_limits<double>::infinity(), hi = -lo;
  for (int64_t i = 0; i < n_cells; ++i) {
    const double v = pt_col[i];
    if (v == v) { ++N; if (v < lo) lo = v; if (v > hi) hi = v; }
  }
  if (N < 30) return H;                                   // the host plan reports the error
  const double Xlo = r_round(lo, 1e4), Xhi = r_round(hi, 1e4);
  if (!(std::isfinite(Xlo) && std::isfinite(Xhi) && std::fabs(Xlo) < 1e8 && std::fabs(Xhi) < 1e8)) return H;
  const long long k0 = std::llround(Xlo * 1e4);
  const long long R = std::llround(Xhi * 1e4) - k0 + 1;
  if (R < 1 || R > (1ll << 24)) return H;
  int bits = 0;
  while ((1ll << bits) < R) ++bits;
  H.passes = bits <= 8 ? 1 : (bits <= 16 ? 2 : 3);
  int minspan = o.minspan;                                // min_span.R:24-26 with q = 1
  if (minspan < 0) minspan = (int)std::nearbyint(-std::log2(-(1.0 / (1.0 * (double)N)) * std::log(1.0 - 0.05)) / 2.5);
  if (minspan < 0 || minspan > (1 << 28)) return H;
  H.minspan = minspan;
  H.maxspan = (int)std::nearbyint(3.0 - std::log2(0.05 / 1.0));      // max_span.R:20 with q = 1
  const double ps[2] = {0.05, 0.95};
  for (int w = 0; w < 2; ++w) {                           // quantile7() of sclane_host.hpp
    const double index = 1.0 + (double)(N - 1) * ps[w];
    const double flo = std::floor(index), fhi = std::ceil(index);
    H.Q.lo[w] = (int)flo;
    H.Q.has_hi[w] = fhi > flo ? 1 : 0;
    H.Q.above[w] = index > flo ? 1 : 0;
    H.Q.h[w] = index - flo;
    H.Q.one_minus_h[w] = 1.0 - H.Q.h[w];
  }
  H.N = (int)N; H.xmin = lo; H.xmax = hi; H.key0 = k0;
  H.eligible = true;
  return H;
}

struct DevicePlan;

struct DevicePlan {
  int* perm = nullptr;
  double* u = nullptr;
  double* off = nullptr;
  Lineage* lin = nullptr;
  int* pstart = nullptr;          // X-compressed layout (sclane_comp.cuh)
  double* pn = nullptr;
  double* pu = nullptr;
  unsigned char* pb = nullptr;
  int* pbstart = nullptr;
  void upload_points(const LineagePlan& P, DevicePool& pool, cudaStream_t s) {
    pstart = (int*)pool.get("pstart", P.pstart.size() * sizeof(int));
    pn = (double*)pool.get("pn", P.pn.size() * sizeof(double));
    pu = (double*)pool.get("pu", P.pu.size() * sizeof(double));
    pb = (unsigned char*)pool.get("pb", P.pb.size());
    CK(cudaMemcpyAsync(pstart, P.pstart.data(), P.pstart.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(pn, P.pn.data(), P.pn.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(pu, P.pu.data(), P.pu.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(pb, P.pb.data(), P.pb.size(), cudaMemcpyHostToDevice, s));
    pbstart = (int*)pool.get("pbstart", sizeof(P.pbstart));
    CK(cudaMemcpyAsync(pbstart, P.pbstart, sizeof(P.pbstart), cudaMemcpyHostToDevice, s));
  }
  // exact-knot mode
  double* px = nullptr;
  int* cand = nullptr;
  void upload_exact(const LineagePlan& P, DevicePool& pool, cudaStream_t s) {
    px = (double*)pool.get("ex_px", P.px.size() * sizeof(double));
    cand = (int*)pool.get("ex_cand", P.cand.size() * sizeof(int));
    CK(cudaMemcpyAsync(px, P.px.data(), P.px.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(cand, P.cand.data(), P.cand.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  }
  // offset quadrature tables
  OffsetNodes* on = nullptr;
  int* oord = nullptr;
  double* oxi_ord = nullptr;
  double* off_ord = nullptr;
  int* ogrp = nullptr;
  std::vector<double> h_oxi_ord, h_off_ord;        // staging (must outlive the async copies: DevicePlan lives until the shard is done)
  void upload_offset_nodes(const LineagePlan& P, DevicePool& pool, cudaStream_t s) {
    const size_t N = (size_t)P.L.N;
    on = (OffsetNodes*)pool.get("off_nodes", sizeof(OffsetNodes));
    oord = (int*)pool.get("off_oord", N * sizeof(int));
    oxi_ord = (double*)pool.get("off_oxi_ord", N * sizeof(double));
    off_ord = (double*)pool.get("off_off_ord", N * sizeof(double));
    h_oxi_ord.resize(N); h_off_ord.resize(N);
    for (size_t i = 0; i < N; ++i) { const size_t p = (size_t)P.oord[i]; h_oxi_ord[i] = P.oxi[p]; h_off_ord[i] = P.off[p]; }
    CK(cudaMemcpyAsync(on, &P.ON, sizeof(OffsetNodes), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(oord, P.oord.data(), N * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(oxi_ord, h_oxi_ord.data(), N * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(off_ord, h_off_ord.data(), N * sizeof(double), cudaMemcpyHostToDevice, s));
    ogrp = (int*)pool.get("off_ogrp", P.ogrp.size() * sizeof(int));
    CK(cudaMemcpyAsync(ogrp, P.ogrp.data(), P.ogrp.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  }
  void upload(const LineagePlan& P, DevicePool& pool, cudaStream_t s) {
    perm = (int*)pool.get("perm", P.perm.size() * sizeof(int));
    u = (double*)pool.get("u", P.u.size() * sizeof(double));
    lin = (Lineage*)pool.get("lineage", sizeof(Lineage));
    CK(cudaMemcpyAsync(perm, P.perm.data(), P.perm.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(u, P.u.data(), P.u.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(lin, &P.L, sizeof(Lineage), cudaMemcpyHostToDevice, s));
    off = nullptr;
    if (!P.off.empty()) {
      off = (double*)pool.get("off", P.off.size() * sizeof(double));
      CK(cudaMemcpyAsync(off, P.off.data(), P.off.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    }
  }
};

// Runs K0 on `s`; on return (stream synchronised) D points at the device tables and hdr carries what the host needs
// (Lineage, D, pbstart, minspan, pseudotime range).  Throws PlanFail when the lineage has no candidate knot.
void device_k0(DevicePool& pool, cudaStream_t s, const double* pt_col, int64_t n_cells, const K0Host& H, const sclane_opts& o,
               LineagePlan& hdr, DevicePlan& D, sclane_profile* prof) {
  const int N = H.N, Npad = (N + 127) / 128 * 128;
  const int nb_cells = (int)((n_cells + kK0Threads - 1) / kK0Threads), nb = (N + kK0Threads - 1) / kK0Threads, nb_pad = (Npad + kK0Threads - 1) / kK0Threads;
  K0Args A;
  std::memset(&A, 0, sizeof(A));
  double* d_pt = (double*)pool.get("k0_pt", (size_t)n_cells * 8);
  A.pt = d_pt; A.n_cells = (int)n_cells; A.N = N; A.Npad = Npad; A.key0 = H.key0; A.step = H.minspan + 1; A.maxspan = H.maxspan;
  A.n_knot_max = o.n_knot_max; A.Q = H.Q;
  A.blk = (int*)pool.get("k0_blk", ((size_t)nb_cells + 2) * 4);
  A.cells = (int*)pool.get("k0_cells", (size_t)N * 4);
  A.xc = (double*)pool.get("k0_xc", (size_t)N * 8);
  A.Xc = (double*)pool.get("k0_Xc", (size_t)N * 8);
  A.key[0] = (unsigned*)pool.get("k0_key0", (size_t)N * 4);
  A.key[1] = (unsigned*)pool.get("k0_key1", (size_t)N * 4);
  A.idx[0] = (int*)pool.get("k0_idx0", (size_t)N * 4);
  A.idx[1] = (int*)pool.get("k0_idx1", (size_t)N * 4);
  A.hist = (int*)pool.get("k0_hist", ((size_t)256 * (size_t)nb + 1) * 4);
  A.Xs = (double*)pool.get("k0_Xs", (size_t)N * 8);
  A.xs = (double*)pool.get("k0_xs", (size_t)N * 8);
  A.qstat = (double*)pool.get("k0_qstat", 4 * 8);
  A.scal = (int*)pool.get("k0_scal", 4 * 4);
  A.perm = D.perm = (int*)pool.get("perm", (size_t)N * 4);
  A.u = D.u = (double*)pool.get("u", (size_t)Npad * 8);
  A.lin = D.lin = (Lineage*)pool.get("lineage", sizeof(Lineage));
  A.pstart = D.pstart = (int*)pool.get("pstart", ((size_t)N + 1) * 4);
  A.pn = D.pn = (double*)pool.get("pn", (size_t)N * 8);
  A.pu = D.pu = (double*)pool.get("pu", (size_t)N * 8);
  A.pb = D.pb = (unsigned char*)pool.get("pb", (size_t)N);
  A.pbstart = D.pbstart = (int*)pool.get("pbstart", sizeof(int) * (NBMAX + 1));
  D.off = nullptr;
  A.res = (K0Result*)pool.get("k0_res", sizeof(K0Result));
  K0Result* h_res = (K0Result*)pool.get("k0_res_h", sizeof(K0Result), true);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  CK(cudaEventCreate(&e0));
  struct EvGuard { cudaEvent_t& a; cudaEvent_t& b; ~EvGuard() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); } } evg{e0, e1};
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, s));
  CK(cudaMemcpyAsync(d_pt, pt_col, (size_t)n_cells * 8, cudaMemcpyHostToDevice, s));
  int launches = 0;
  k0_valid_count<<<nb_cells, kK0Threads, 0, s>>>(A);
  k0_scan<<<1, kK0Threads, 0, s>>>(A.blk, nb_cells);
  k0_compact<<<nb_cells, kK0Threads, 0, s>>>(A);
  launches += 3;
  int cur = 0;
  for (int p = 0; p < H.passes; ++p) {
    k0_radix_hist<<<nb, kK0Threads, 0, s>>>(A.key[cur], N, 8 * p, nb, A.hist);
    k0_scan<<<1, kK0Threads, 0, s>>>(A.hist, 256 * nb);
    k0_radix_scatter<<<nb, kK0Threads, 0, s>>>(A.key[cur], A.idx[cur], A.key[cur ^ 1], A.idx[cur ^ 1], N, 8 * p, nb, A.hist);
    cur ^= 1;
    launches += 3;
  }
  A.key_s = A.key[cur]; A.idx_s = A.idx[cur];
  k0_sorted<<<nb, kK0Threads, 0, s>>>(A);
  k0_scan<<<1, kK0Threads, 0, s>>>(A.blk, nb);
  k0_points<<<nb, kK0Threads, 0, s>>>(A, nb);
  k0_quantile_stats<<<4, kK0Threads, 0, s>>>(A);
  k0_knots<<<1, kK0Threads, 0, s>>>(A);
  k0_bucket_moments<<<NBMAX, 256, 0, s>>>(A);
  k0_fill<<<nb_pad, kK0Threads, 0, s>>>(A);
  launches += 7;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h_res, A.res, sizeof(K0Result), cudaMemcpyDeviceToHost, s));
  CK(cudaEventRecord(e1, s));
  CK(cudaStreamSynchronize(s));
  g_launches.fetch_add(launches);
  if (prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) prof->ms[SCLANE_STAGE_PLAN] += ms;
    prof->launches[SCLANE_STAGE_PLAN] += launches;
    // (h2d_bytes / d2h_bytes count the counts and the records, as with the host plan, whose tables were not counted either)
  }
  hdr.L = h_res->L;
  hdr.D = h_res->D;
  for (int b = 0; b <= NBMAX; ++b) hdr.pbstart[b] = h_res->pbstart[b];
  hdr.minspan = H.minspan;
  hdr.pt_min = H.xmin; hdr.pt_max = H.xmax;
  hdr.exact = false;
  hdr.ON.on = 0;
  hdr.error.clear();
  if (hdr.L.T < 1) { hdr.error = "no candidate knots (pseudotime has too few distinct values)"; throw PlanFail{hdr.error}; }
}

template <int STAGE, bool HYB = false>
void launch_stage(const StageArgs& A, cudaStream_t s) {
  k_stage<STAGE, HYB><<<A.n_genes, kThreads, 0, s>>>(A);
  CK(cudaGetLastError());
  g_launches.fetch_add(1);
}

// theta.ml on Chebyshev nodes in k_comp_stage<FINAL> (sclane_comp.cuh); SCLANE_THETA_NODES=0 walks the points instead (A/B).
// bucket nodes of the X-compressed route (sclane_comp.cuh); SCLANE_BUCKET_NODES=0 walks the points instead (A/B, tests)
bool bucket_nodes_enabled() {
  const char* e = std::getenv("SCLANE_BUCKET_NODES");
  return !(e && e[0] == '0');
}
bool theta_nodes_enabled() {
  const char* e = std::getenv("SCLANE_THETA_NODES");      // read per call: the tests switch it inside one process
  return !(e && e[0] == '0');
}

// Hybrid final stage on thread-block clusters: CL CTAs (on CL SMs of one GPC) per gene, partial moments exchanged through
// distributed shared memory (sclane_device.cuh).  The cost of glm.nb varies 30x between genes and a single CTA cannot use more
// than its SM's FP64 pipe, so the genes that hit the alternation limit used to run alone for tens of ms at the end of a launch.
template <int CL>
void launch_final_cluster(const StageArgs& A, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)A.n_genes * CL, 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, k_stage<STAGE_FINAL, true, CL>, A));
  g_launches.fetch_add(1);
}

// Cluster size of the hybrid final stage.  Measured on C4 (10,000 genes x 2 lineages of 30,000 cells, 1 x B200): final stage
// 249 ms with CL = 1, 241 ms (2), 294 ms (4), 406 ms (8).  With 16,000 gene-lineages per run the launch is throughput bound
// (sum of CTA time / 444 resident CTAs = 230 ms of the 264 ms): a cluster shortens the slowest gene 8x but every CTA of it now
// idles at the cluster barrier and during the (replicated) serial algebra while holding one of the SM's three CTA slots.  So
// the default is CL = 1; the cluster kernels stay built for small shards (SCLANE_FINAL_CLUSTER = 2 | 4 | 8), where the launch
// is as long as its slowest gene.  The choice may only depend on the environment / lineage, never on the batch: the
// summation order -- hence the last bits of a record -- differs between cluster sizes.
int final_cluster_size(int n_cells_lineage) {
  static const int forced = [] { const char* e = std::getenv("SCLANE_FINAL_CLUSTER"); return e ? std::atoi(e) : 0; }();
  if ((forced == 2 || forced == 4 || forced == 8) && n_cells_lineage >= 4096) return forced;
  return 1;
}

// K2 = k_sweep_moments (streaming, HBM bound) + k_sweep_select (warp per gene).  `first` also gathers the y moments.
void launch_sweep(const StageArgs& A, bool first, cudaStream_t s, EventTimer* tm, sclane_profile* prof);

int forward_iterations(int M) { return (M - 1 + 1) / 2; }   // m = 1, 3, 5, ... until m >= M (marge2.R:752-757)

// Host side of the staged transfer: int32 -> uint16 with a pool of threads.  Returns false (and the caller sends the batch
// as int32) when a value does not fit.
// Worker threads that live as long as a transfer: a chunk of 128 MB is narrowed in ~1.5 ms, which is what spawning and joining
// 16 threads per chunk would cost too.  run() hands every worker a slice of the chunk and returns when all are done; false when
// a value does not fit (the caller then sends the rows as int32).
class NarrowPool {
 public:
  explicit NarrowPool(int n_threads) : nt_(n_threads < 1 ? 1 : n_threads), bad_((size_t)nt_, 0) {
    for (int t = 0; t < nt_; ++t) th_.emplace_back([this, t] { loop(t); });
  }
  ~NarrowPool() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; ++gen_; }
    cv_.notify_all();
    for (auto& t : th_) if (t.joinable()) t.join();
  }
  bool run(const int32_t* src, size_t n, uint16_t* dst) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      src_ = src; dst_ = dst; n_ = n; pending_ = nt_; ++gen_;
    }
    cv_.notify_all();
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [&] { return pending_ == 0; });
    for (int b : bad_) if (b) return false;
    return true;
  }

 private:
  void loop(int t) {
    long long seen = 0;
    for (;;) {
      const int32_t* src; uint16_t* dst; size_t n;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        if (stop_) return;
        seen = gen_; src = src_; dst = dst_; n = n_;
      }
      const size_t a = (n * (size_t)t / (size_t)nt_) & ~(size_t)15, b = t + 1 == nt_ ? n : ((n * (size_t)(t + 1) / (size_t)nt_) & ~(size_t)15);
      bad_[(size_t)t] = narrow_range_u16(src + a, b - a, dst + a) ? 0 : 1;
      bool last;
      { std::lock_guard<std::mutex> lk(mu_); last = --pending_ == 0; }
      if (last) done_.notify_one();
    }
  }
  int nt_;
  std::vector<int> bad_;
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const int32_t* src_ = nullptr;
  uint16_t* dst_ = nullptr;
  size_t n_ = 0;
  int pending_ = 0;
  long long gen_ = 0;
  bool stop_ = false;
};

// ---------------------------------------------------------------------------------------------
// Host -> device transfer of the caller's counts: its own host thread + copy stream, running ahead of the kernels.
//
// The PCIe link (55 GB/s measured) bounds the end-to-end time of a large matrix: 16 GB of int32 take 291 ms, the kernels 126 ms.
// Counts almost always fit 16 bits, so the bytes can be halved on the host on their way out -- but narrowing costs host memory
// bandwidth too (16 threads: ~60 GB/s of int32 consumed on the benchmark box).  The feeder therefore SPLITS every batch: the
// genes [0, nA) are narrowed into a pinned staging buffer and sent as uint16, the genes [nA, nb) are copied straight from the
// caller's buffer as int32 (no CPU involved).  Per batch it first queues the direct part -- the DMA engine is busy at once --
// and narrows the staged part in chunks while that copy is on the wire; nA follows the measured rates (r_n = narrowing,
// r_w = wire): both legs end together when f = nA / nb = 2 r_n / (2 r_w + r_n).  Pageable caller memory (an R matrix) cannot be
// DMA-ed directly, so it is always staged (f = 1).  Three things overlap: narrow(k), wire(k - 1 .. k), kernels(k - 1).
// Layout of a batch in its device buffer: the direct rows sit where they would sit in an all-int32 batch; the uint16 rows are
// packed from offset 0 (inside the space the int32 rows of the same genes would take), so a batch whose staged part does not
// fit 16 bits is simply re-sent as int32.
// ---------------------------------------------------------------------------------------------
class H2DFeeder {
 public:
  struct Layout { int64_t nA = 0; int a16 = 0; };      // genes [0, nA): uint16 rows (a16 = 1) or int32 rows; genes [nA, nb): int32 rows
  enum Mode { DIRECT = 0, STAGED = 1, MIXED = 2 };

  H2DFeeder(int device, cudaStream_t cs, const int32_t* h_counts, int64_t n_cells, int64_t g0, int64_t g1, int64_t B, Mode mode,
            int* d_raw0, int* d_raw1, uint16_t* h_stage0, uint16_t* h_stage1)
      : device_(device), cs_(cs), h_(h_counts), n_cells_(n_cells), g0_(g0), g1_(g1), B_(B), mode_(mode) {
    d_raw_[0] = d_raw0; d_raw_[1] = d_raw1; h_stage_[0] = h_stage0; h_stage_[1] = h_stage1;
    n_batches_ = (g1 - g0 + B - 1) / B;
    layouts_.resize((size_t)n_batches_);
    for (int i = 0; i < 2; ++i) {
      CK(cudaEventCreateWithFlags(&ev_h2d_[i], cudaEventDisableTiming | cudaEventBlockingSync));
      CK(cudaEventCreateWithFlags(&ev_free_[i], cudaEventDisableTiming));
      CK(cudaEventCreate(&ev_t0_[i]));
      CK(cudaEventCreate(&ev_t1_[i]));
    }
    th_ = std::thread([this] { run(); });
  }
  ~H2DFeeder() {
    { std::lock_guard<std::mutex> lk(mu_); abort_ = true; }
    cv_.notify_all();
    if (th_.joinable()) th_.join();
    cudaStreamSynchronize(cs_);                       // nothing of this transfer may be in flight when the buffers are reused
    for (int i = 0; i < 2; ++i) {
      if (ev_h2d_[i]) cudaEventDestroy(ev_h2d_[i]);
      if (ev_free_[i]) cudaEventDestroy(ev_free_[i]);
      if (ev_t0_[i]) cudaEventDestroy(ev_t0_[i]);
      if (ev_t1_[i]) cudaEventDestroy(ev_t1_[i]);
    }
  }
  H2DFeeder(const H2DFeeder&) = delete;
  H2DFeeder& operator=(const H2DFeeder&) = delete;

  // compute thread: blocks until every copy of batch k and its completion event are queued on the copy stream
  Layout wait(int64_t k) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return enqueued_ > k || failed_; });
    if (enqueued_ <= k) throw CudaFail{error_};
    return layouts_[(size_t)k];
  }
  cudaEvent_t ready(int64_t k) const { return ev_h2d_[k & 1]; }
  const int* raw(int64_t k) const { return d_raw_[k & 1]; }
  // compute thread: the gather of batch k is queued on `s`; its device buffer may be overwritten once that has run
  void consumed(int64_t k, cudaStream_t s) {
    CK(cudaEventRecord(ev_free_[k & 1], s));
    { std::lock_guard<std::mutex> lk(mu_); gathered_ = k + 1; }
    cv_.notify_all();
  }
  double h2d_bytes() const { return bytes_; }          // valid after the last wait() returned
  long long copies() const { return copies_; }

 private:
  void run() {
    try {
      CK(cudaSetDevice(device_));
      double r_n = 75e9, r_w = 52e9;                   // bytes of int32 per second: narrowing (all host threads), wire
      std::unique_ptr<NarrowPool> pool;
      if (mode_ != DIRECT) {
        const unsigned hw = host_threads();
        pool.reset(new NarrowPool(hw == 0 ? 4 : (int)std::min<unsigned>(hw, 16)));
        if (mode_ == MIXED && n_batches_ > 0) {
          // calibrate the narrowing rate on a first slice (<= 32 MB) before the split of batch 0 is chosen: the default describes a
          // host that is otherwise idle, and eight ranks sharing one host see a small fraction of it
          const int64_t nb0 = std::min<int64_t>(B_, g1_ - g0_);
          const size_t n_el = std::min<size_t>((size_t)nb0 * (size_t)n_cells_, (size_t)1 << 23);
          const auto t0 = std::chrono::steady_clock::now();
          pool->run(h_ + (size_t)g0_ * (size_t)n_cells_, n_el, h_stage_[0]);
          const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
          if (dt > 1e-5) r_n = std::min(r_n, (double)n_el * 4.0 / dt);
        }
      }
      bool had_direct[2] = {false, false};
      size_t direct_bytes[2] = {0, 0};
      for (int64_t k = 0; k < n_batches_; ++k) {
        const int buf = (int)(k & 1);
        const int64_t kb0 = g0_ + k * B_, kb1 = std::min<int64_t>(kb0 + B_, g1_);
        const int64_t nbk = kb1 - kb0;
        const int32_t* src = h_ + (size_t)kb0 * (size_t)n_cells_;
        if (k >= 2) {
          {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return gathered_ >= k - 1 || abort_; });
            if (abort_) return;
          }
          CK(cudaStreamWaitEvent(cs_, ev_free_[buf], 0));              // the gather of batch k - 2 has consumed this device buffer
          CK(cudaEventSynchronize(ev_h2d_[buf]));                      // ... and its copies have left the staging buffer
          if (had_direct[buf]) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ev_t0_[buf], ev_t1_[buf]) == cudaSuccess && ms > 0.05f) r_w = (double)direct_bytes[buf] / (ms * 1e-3);
            else cudaGetLastError();
          }
        } else {
          std::lock_guard<std::mutex> lk(mu_);
          if (abort_) return;
        }
        double f = mode_ == DIRECT ? 0.0 : (mode_ == STAGED ? 1.0 : 2.0 * r_n / (2.0 * r_w + r_n));
        if (f > 0.93) f = 1.0;
        if (f < 0.07) f = 0.0;
        int64_t nA = (int64_t)std::llround(f * (double)nbk);
        if (nA > nbk) nA = nbk;
        Layout lay;
        lay.nA = nA; lay.a16 = nA > 0 ? 1 : 0;
        char* dbase = reinterpret_cast<char*>(d_raw_[buf]);
        had_direct[buf] = false;
        if (nA < nbk) {                                                // direct part first: keeps the link busy while the CPU narrows
          const size_t off = (size_t)nA * (size_t)n_cells_ * 4, bytes = (size_t)(nbk - nA) * (size_t)n_cells_ * 4;
          CK(cudaEventRecord(ev_t0_[buf], cs_));
          CK(cudaMemcpyAsync(dbase + off, src + (size_t)nA * (size_t)n_cells_, bytes, cudaMemcpyHostToDevice, cs_));
          CK(cudaEventRecord(ev_t1_[buf], cs_));
          had_direct[buf] = true; direct_bytes[buf] = bytes;
          bytes_ += (double)bytes; copies_++;
        }
        if (nA > 0) {
          const size_t n_el = (size_t)nA * (size_t)n_cells_;
          const size_t chunk = (size_t)1 << 25;                        // 128 MB of int32 per chunk: ~2 ms of narrowing
          bool fits = true;
          double t_n = 0.0;
          for (size_t c0 = 0; c0 < n_el && fits; c0 += chunk) {
            const size_t c1 = std::min(n_el, c0 + chunk);
            const auto t0 = std::chrono::steady_clock::now();
            fits = pool->run(src + c0, c1 - c0, h_stage_[buf] + c0);
            t_n += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (fits) {
              CK(cudaMemcpyAsync(dbase + c0 * 2, h_stage_[buf] + c0, (c1 - c0) * 2, cudaMemcpyHostToDevice, cs_));
              bytes_ += (double)(c1 - c0) * 2.0; copies_++;
            }
            { std::lock_guard<std::mutex> lk(mu_); if (abort_) return; }
          }
          if (!fits) {                                                 // a count > 65534 (or < 0): this part travels as int32 after all
            lay.a16 = 0;
            CK(cudaMemcpyAsync(dbase, src, n_el * 4, cudaMemcpyHostToDevice, cs_));
            bytes_ += (double)n_el * 4.0; copies_++;
          } else if (t_n > 1e-4) {
            r_n = (double)n_el * 4.0 / t_n;
          }
        }
        CK(cudaEventRecord(ev_h2d_[buf], cs_));
        {
          std::lock_guard<std::mutex> lk(mu_);
          layouts_[(size_t)k] = lay;
          enqueued_ = k + 1;
        }
        cv_.notify_all();
      }
    } catch (const CudaFail& e) {
      { std::lock_guard<std::mutex> lk(mu_); failed_ = true; error_ = e.msg; }
      cv_.notify_all();
    } catch (const std::exception& e) {
      { std::lock_guard<std::mutex> lk(mu_); failed_ = true; error_ = std::string("transfer thread: ") + e.what(); }
      cv_.notify_all();
    }
  }

  int device_;
  cudaStream_t cs_;
  const int32_t* h_;
  int64_t n_cells_, g0_, g1_, B_, n_batches_ = 0;
  Mode mode_;
  int* d_raw_[2];
  uint16_t* h_stage_[2];
  cudaEvent_t ev_h2d_[2] = {nullptr, nullptr}, ev_free_[2] = {nullptr, nullptr}, ev_t0_[2] = {nullptr, nullptr}, ev_t1_[2] = {nullptr, nullptr};
  std::thread th_;
  std::mutex mu_;
  std::condition_variable cv_;
  int64_t enqueued_ = 0, gathered_ = 0;
  bool abort_ = false, failed_ = false;
  std::string error_;
  std::vector<Layout> layouts_;
  double bytes_ = 0.0;
  long long copies_ = 0;
};

// The lineage plan (host: knots, sort, buckets, points) is built on a background thread while the first batch of counts is
// already on its way to the device; get() joins it.
struct PlanSource {
  const LineagePlan* plan = nullptr;
  std::shared_future<bool> ready;          // false = the plan could not be built (plan->error says why)
  const LineagePlan& get() {
    if (ready.valid() && !ready.get()) throw PlanFail{plan->error};
    return *plan;
  }
  // K0 on the device (sclane_k0_device.cuh): every shard builds the tables on its own GPU; `header` receives the host-side
  // summary (Lineage, D, ...) of the shard with publish set
  const K0Host* k0 = nullptr;
  const double* pt_col = nullptr;
  LineagePlan* header = nullptr;
  bool publish = false;
};

// Run genes [g0, g1) of one lineage on one device.  Exactly one of h_counts / d_counts is non-null;
// both are gene-major with leading dimension n_cells.  Records go to out[(g - out_g0) * n_lineages + lin_idx].
void run_shard(int device, const int32_t* h_counts, const int32_t* d_counts, int64_t n_cells, int64_t g0, int64_t g1,
               PlanSource& PS, const sclane_opts& o, bool use_offset, sclane_record* out, int64_t out_g0, int n_lineages,
               int lin_idx, sclane_profile& prof) {
  if (g1 <= g0) { if (!PS.k0) PS.get(); return; }
  CK(cudaSetDevice(device));
  DevicePool& pool = pool_for(device);
  if (!pool.stream) CK(cudaStreamCreateWithFlags(&pool.stream, cudaStreamNonBlocking));
  cudaStream_t s = pool.stream;
  const int64_t G = g1 - g0;
  // batch size from free memory (+ what the pool already holds): raw (host path) + sorted counts + mu + moments + records
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  const size_t pool_held = pool.held();
  // batch slots in flight (see Slot below): sclane_opts.batch_streams, default 2
  const int n_slots = o.batch_streams > 0 ? std::min(o.batch_streams, 4) : 2;
  auto batch_size = [&](int Npad, int Dpad, bool use_comp) -> int64_t {
    const size_t per_gene_slot = (size_t)Npad * 12 + sizeof(GeneOutDev) + sizeof(GeneStats) +
                                 (size_t)NBMAX * NACC * 8 + (use_comp ? (size_t)Dpad * 16 + (size_t)kHistCap * 4 + (size_t)kBigCap * 8 + sizeof(CompGeneDev) : 0) +
                                 (o.approx_knot ? 0 : (size_t)Dpad * 16) + (size_t)(NBMAX * kOffJ) * 16 + (use_offset ? 3 * (size_t)kHtCap * 8 + (size_t)(kOffKMax * kOffJ) * 16 : 0);
    const size_t per_gene = (h_counts ? (size_t)n_cells * 8 : 0) + (size_t)n_slots * per_gene_slot;
    size_t budget = (size_t)((double)(free_b + pool_held) * 0.6);
    const size_t cap = (size_t)SCLANE_BATCH_CAP_GB << 30;
    if (budget > cap) budget = cap;
    int64_t Bv = (int64_t)(budget / per_gene);
    // host counts: aim for ~8 batches (but >= 2048 genes each, several waves of CTAs) so that the H2D copy of batch k + 1
    // runs on the copy stream while the kernels of batch k execute
    if (o.genes_per_batch > 0) { if (o.genes_per_batch < Bv) Bv = o.genes_per_batch; }
    else if (h_counts) {
      // per-cell final fits (offsets) are compute bound and every launch ends in a tail of slow genes (~160 ms at 30,000
      // cells): two batches are enough to hide the transfer; otherwise ~8 batches keep the PCIe link busy
      const bool heavy_final = use_offset || !use_comp;
      const int64_t want = std::max<int64_t>(2048, heavy_final ? (G + 1) / 2 : (G + 7) / 8);
      if (want < Bv) Bv = want;
    } else if (n_slots > 1 && !(use_offset || !use_comp)) {
      // device-resident counts: two batches per slot (but not below a couple of waves of CTAs each), so that every launch tail
      // and every host-side pause between batches has another batch's kernels to hide behind.  Not for the per-cell final fits
      // (offsets): there a launch is as long as its slowest genes (25 glm.nb alternations), which must all start at once
      // (measured, tools/ab_slots.sh: a 2,500-gene shard gains nothing from being split -- 16.4 ms in one batch, 16.8 in two
      // or three -- while 20,000 genes go from 117.7 ms to 110.0)
      static const int64_t min_batch = [] { const char* e = std::getenv("SCLANE_MIN_BATCH"); const long v = e ? std::atol(e) : 0; return (int64_t)(v > 0 ? v : 2048); }();
      if (G >= 2 * min_batch) {
        const int64_t want = std::max<int64_t>(min_batch, (G + 2 * n_slots - 1) / (2 * n_slots));
        if (want < Bv) Bv = want;
      }
    }
    if (Bv > G) Bv = G;
    if (Bv < 1) throw CudaFail{"not enough free device memory for a single gene"};
    return (G + ((G + Bv - 1) / Bv) - 1) / ((G + Bv - 1) / Bv);      // equal batches: a short last batch is mostly launch tail
  };
  // Before the plan exists only its size bounds are known: guess the common case (pseudotime in [0, 1]: X-compressed route,
  // <= 10,001 abscissae), start the first transfer, and redo it in the rare case the real plan asks for a smaller batch.
  const int Npad_guess = (int)((n_cells + 127) / 128 * 128);
  int64_t B = batch_size(Npad_guess, 10016, o.force_per_cell == 0);
  int64_t n_batches = (G + B - 1) / B;
  // transfer of the caller's counts (H2DFeeder above; sclane_opts.h2d_mode: 0 automatic, 1 direct int32, 2 staged uint16).
  // Automatic: small inputs are copied directly; large pageable ones (an R integer matrix: cudaMemcpyAsync from it is staged by
  // the driver at ~12 GB/s and blocks) go through the library's own pinned staging; large pinned ones are split between the two.
  std::unique_ptr<H2DFeeder> feeder;
  cudaStream_t cs = nullptr;
  auto start_feeder = [&]() {                       // (re)sizes the transfer buffers for the current B and starts the transfer
    feeder.reset();
    H2DFeeder::Mode mode = H2DFeeder::DIRECT;
    if (o.h2d_mode == 2) mode = H2DFeeder::STAGED;
    else if (o.h2d_mode == 0 && (double)G * (double)n_cells * 4.0 >= 256e6) {
      bool pageable = false;
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, h_counts) == cudaSuccess) pageable = at.type == cudaMemoryTypeUnregistered;
      else cudaGetLastError();
      mode = pageable ? H2DFeeder::STAGED : H2DFeeder::MIXED;
    }
    int* r0 = (int*)pool.get("raw", (size_t)B * (size_t)n_cells * 4);
    int* r1 = n_batches > 1 ? (int*)pool.get("raw_b", (size_t)B * (size_t)n_cells * 4) : r0;
    uint16_t *h0 = nullptr, *h1 = nullptr;
    if (mode != H2DFeeder::DIRECT) {
      h0 = (uint16_t*)pool.get("stage_a", (size_t)B * (size_t)n_cells * 2, true);
      h1 = n_batches > 1 ? (uint16_t*)pool.get("stage_b", (size_t)B * (size_t)n_cells * 2, true) : h0;
    }
    feeder.reset(new H2DFeeder(device, cs, h_counts, n_cells, g0, g1, B, mode, r0, r1, h0, h1));
  };
  if (h_counts) {
    if (!pool.copy_stream) CK(cudaStreamCreateWithFlags(&pool.copy_stream, cudaStreamNonBlocking));
    cs = pool.copy_stream;
    start_feeder();                                 // the first batch travels while the plan is being built
  }
  const LineagePlan* Pp = nullptr;
  LineagePlan k0_hdr;                               // device-built plan: only the header lives on the host
  DevicePlan D;
  if (PS.k0) {
    try { device_k0(pool, s, PS.pt_col, n_cells, *PS.k0, o, k0_hdr, D, &prof); }
    catch (const PlanFail&) { if (PS.publish) { PS.header->L = k0_hdr.L; PS.header->error = k0_hdr.error; } throw; }
    if (PS.publish) {
      LineagePlan& Hd = *PS.header;
      Hd.L = k0_hdr.L; Hd.D = k0_hdr.D; Hd.minspan = k0_hdr.minspan; Hd.pt_min = k0_hdr.pt_min; Hd.pt_max = k0_hdr.pt_max;
      Hd.exact = false; Hd.error.clear();
    }
    Pp = &k0_hdr;
  } else Pp = &PS.get();                            // (on failure the feeder's destructor drains the copy stream)
  const LineagePlan& P = *Pp;
  const Lineage& L = P.L;
  g_trace.at("plan ready");
  // X-compressed path (sclane_comp.cuh): worth it when the 4-decimal rounding leaves clearly fewer abscissae than cells
  // approx.knot = FALSE always runs on the abscissae (its sweep is a scan over them; sclane_exact_device.cuh)
  const bool exact = P.exact;
  const bool use_comp = exact || (!o.force_per_cell && P.D > 0 && (double)P.D <= 0.8 * (double)L.N);
  const int Dpad = (P.D + 31) / 32 * 32;
  {
    const int64_t B_real = batch_size(L.Npad, Dpad, use_comp);
    if (B_real != B) {                              // the guess was off (other route / more abscissae): restart the transfer
      feeder.reset();
      B = B_real;
      n_batches = (G + B - 1) / B;
      if (h_counts) start_feeder();
    }
  }
  if (!PS.k0) {
    D.upload(P, pool, s);
    if (use_comp) D.upload_points(P, pool, s);
  }
  if (exact) D.upload_exact(P, pool, s);
  if (exact && use_offset && !(P.ON.on && !o.offset_per_cell)) throw CudaFail{"approx.knot = FALSE with a size-factor offset needs the offset quadrature (offsets spanning more than e^32, or offset_per_cell set)"};
  // fits with an offset: intercept-only models through the offset quadrature (needs the histogram of k_aggregate)
  const bool use_offq = use_comp && use_offset && P.ON.on && !o.offset_per_cell && !o.basis_only;
  if (use_offq) D.upload_offset_nodes(P, pool, s);
  CK(cudaStreamSynchronize(s));                     // plan tables are in place before the kernels start
  // Work buffers of one batch slot.  The batches of a shard run on n_slots streams, batch k on slot k % n_slots: the kernels of
  // a batch are a chain of ~10 dependent launches, each ending in a tail of half-empty SMs (one CTA per gene, gene cost varies
  // 10x), and between two batches the device used to wait for the host (D2H of the records, their finalisation, the next
  // batch's launches).  With two chains in flight the block scheduler fills the tail of one with the CTAs of the other.
  // bucket nodes of the X-compressed route: the gene independent part, once per lineage
  BucketNodes* d_bn = nullptr;
  if (use_comp && !exact && bucket_nodes_enabled()) {
    d_bn = (BucketNodes*)pool.get("bucket_nodes", sizeof(BucketNodes));
    k_bucket_nodes<<<1, kThreads, 0, s>>>(D.lin, D.pn, D.pu, D.pbstart, d_bn);
    CK(cudaGetLastError());
    g_launches.fetch_add(1);
    prof.launches[SCLANE_STAGE_PLAN]++;
    CK(cudaStreamSynchronize(s));                   // the other batch slots' streams read it
  }
  // hybrid final stage (offsets): theta.ml on Chebyshev nodes of the (offset bin, bucket) groups (sclane_device.cuh)
  const bool hyb_theta_nodes = use_offq && !exact && theta_nodes_enabled() && P.ON.K * (L.T + 1) <= kHtUnits;
  struct Slot {
    cudaStream_t s = nullptr;
    int* d_ys = nullptr; double* d_mu = nullptr; double* d_mom = nullptr; GeneStats* d_stats = nullptr; GeneOutDev* d_out = nullptr; GeneOutDev* h_out = nullptr;
    double* d_s1 = nullptr; double* d_q = nullptr; int* d_tail = nullptr; int* d_big = nullptr; int* d_big_tmp = nullptr; double* d_okeys = nullptr; int* d_order = nullptr;
    CompGeneDev* d_cg = nullptr; double* d_off_s = nullptr; double* d_off_q = nullptr; OffGeneDev* d_offg = nullptr; int* d_counters = nullptr;
    double* d_a0p = nullptr; double* d_b0p = nullptr; double* d_ex_scratch = nullptr;
    double* d_ht = nullptr;            // hybrid final stage: theta.ml node lists, [B][3][kHtCap]
    double* d_bn_s = nullptr; double* d_bn_q = nullptr;      // bucket nodes: per-gene node weights [B][NBMAX][16]
    std::unique_ptr<EventTimer> tm;
    int64_t b0 = 0; int nb = 0; bool busy = false;
  };
  static int sm_count[SCLANE_MAX_GPUS] = {0};
  if (sm_count[device % SCLANE_MAX_GPUS] == 0) { int v = 0; CK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device)); sm_count[device % SCLANE_MAX_GPUS] = v; }
  const int ex_slots = sm_count[device % SCLANE_MAX_GPUS] * kExactCtasPerSm;
  const int n_slots_used = (int)std::min<int64_t>((int64_t)n_slots, n_batches);      // a single batch needs a single set of buffers
  std::vector<Slot> slots((size_t)n_slots_used);
  for (int si = 0; si < n_slots_used; ++si) {
    Slot& Z = slots[(size_t)si];
    if (si == 0) Z.s = s;
    else {
      if (pool.extra_streams.size() < (size_t)si) pool.extra_streams.resize((size_t)si, nullptr);
      if (!pool.extra_streams[(size_t)si - 1]) CK(cudaStreamCreateWithFlags(&pool.extra_streams[(size_t)si - 1], cudaStreamNonBlocking));
      Z.s = pool.extra_streams[(size_t)si - 1];
    }
    const std::string sfx = si == 0 ? std::string() : "#" + std::to_string(si);
    Z.d_ys = (int*)pool.get("ys" + sfx, (size_t)B * (size_t)L.Npad * 4);
    Z.d_mu = (double*)pool.get("mu" + sfx, (size_t)B * (size_t)L.Npad * 8);
    Z.d_mom = (double*)pool.get("mom" + sfx, (size_t)B * (size_t)NBMAX * NACC * 8);
    Z.d_stats = (GeneStats*)pool.get("stats" + sfx, (size_t)B * sizeof(GeneStats));
    Z.d_out = (GeneOutDev*)pool.get("out" + sfx, (size_t)B * sizeof(GeneOutDev));
    Z.h_out = (GeneOutDev*)pool.get("h_out" + sfx, (size_t)B * sizeof(GeneOutDev), true);
    if (use_comp) {
      Z.d_s1 = (double*)pool.get("comp_s1" + sfx, (size_t)B * (size_t)Dpad * 8);
      Z.d_q = (double*)pool.get("comp_q" + sfx, (size_t)B * (size_t)Dpad * 8);
      Z.d_tail = (int*)pool.get("comp_tail" + sfx, (size_t)B * (size_t)kHistCap * 4);
      Z.d_big = (int*)pool.get("comp_big" + sfx, (size_t)B * (size_t)kBigCap * 4);
      Z.d_big_tmp = (int*)pool.get("comp_big_tmp" + sfx, (size_t)B * (size_t)kBigCap * 4);
      Z.d_cg = (CompGeneDev*)pool.get("comp_hdr" + sfx, (size_t)B * sizeof(CompGeneDev));
      if (d_bn) {
        Z.d_bn_s = (double*)pool.get("bn_s" + sfx, (size_t)B * (size_t)(NBMAX * kOffJ) * 8);
        Z.d_bn_q = (double*)pool.get("bn_q" + sfx, (size_t)B * (size_t)(NBMAX * kOffJ) * 8);
      }
    }
    Z.d_okeys = (double*)pool.get("order_keys" + sfx, (size_t)B * 8);
    Z.d_order = (int*)pool.get("order" + sfx, (size_t)B * 4);
    if (use_offq) {
      Z.d_off_s = (double*)pool.get("off_s" + sfx, (size_t)B * (size_t)(kOffKMax * kOffJ) * 8);
      Z.d_off_q = (double*)pool.get("off_q" + sfx, (size_t)B * (size_t)(kOffKMax * kOffJ) * 8);
      Z.d_offg = (OffGeneDev*)pool.get("off_gene" + sfx, (size_t)B * sizeof(OffGeneDev));
      if (hyb_theta_nodes) Z.d_ht = (double*)pool.get("hyb_theta" + sfx, (size_t)B * 3 * (size_t)kHtCap * 8);
    }
    Z.d_counters = (int*)pool.get("wg_counters" + sfx, 8 * sizeof(int));
    if (exact) {
      Z.d_a0p = (double*)pool.get("ex_a0p" + sfx, (size_t)B * (size_t)Dpad * 8);
      Z.d_b0p = (double*)pool.get("ex_b0p" + sfx, (size_t)B * (size_t)Dpad * 8);
      Z.d_ex_scratch = (double*)pool.get("ex_scratch" + sfx, (size_t)ex_slots * (size_t)(2 + kExMom) * (size_t)Dpad * 8);
    }
    Z.tm.reset(new EventTimer(Z.s));
  }
  // persistent warp-per-gene kernels: one CTA slot per resident CTA of the device
  const int wg_slots = sm_count[device % SCLANE_MAX_GPUS] * SCLANE_WG_MINB;
  const bool use_wg = use_comp && o.comp_warp_per_gene && !exact;
  if (exact) {
    static std::once_flag ex_once[SCLANE_MAX_GPUS];
    std::call_once(ex_once[device % SCLANE_MAX_GPUS], [] { CK(cudaFuncSetAttribute(k_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ExactShared))); });
  }
  const int n_iter = forward_iterations(o.M);
  // a finished batch: records finalised on the host (z / p-values, LRT, pchisq per gene, ~1-2 us each -- spread over host
  // threads, chunk sums in chunk order) while the other slot's kernels keep the device busy
  auto finish = [&](Slot& Z) {
    if (!Z.busy) return;
    if (h_counts) stream_wait_blocking(Z.s); else CK(cudaStreamSynchronize(Z.s));
    g_trace.at("batch done on the device");
    Z.tm->collect(prof);
    const int nb = Z.nb;
    const int64_t b0 = Z.b0;
    GeneOutDev* h_out = Z.h_out;
    long long sw_chunk[kPlanChunks] = {0};
    const bool basis_only = o.basis_only != 0;
    parallel_chunks((size_t)nb, [&](int c, size_t i0, size_t i1) {
      long long sw = 0;
      for (size_t i = i0; i < i1; ++i) {
        GeneOut go;
        go.st = h_out[i].st; go.alt = h_out[i].alt; go.null = h_out[i].null;
        sclane_record& r = out[((b0 - out_g0) + (int64_t)i) * n_lineages + lin_idx];
        finalize_record(go, L, r, basis_only);
        sw += go.st.fwd_iters;                       // sweeps actually run for this gene: one per forward iteration it entered
      }
      sw_chunk[c] = sw;
    }, 1024);
    long long sw = 0;
    for (int c = 0; c < kPlanChunks; ++c) sw += sw_chunk[c];
    prof.sweep_gene_iters += sw;
    prof.sweep_bytes += 12.0 * (double)L.N * (double)sw;
    prof.sweep_bytes += 8.0 * (double)L.N * (double)n_iter;
    g_trace.at("records finalised");
    Z.busy = false;
  };
  for (int64_t b0 = g0, kb = 0; b0 < g1; b0 += B, ++kb) {
    const int64_t b1 = std::min<int64_t>(b0 + B, g1);
    const int nb = (int)(b1 - b0);
    Slot& Z = slots[(size_t)(kb % n_slots_used)];
    finish(Z);                                        // the batch that used this slot before (no-op for the first n_slots batches)
    Z.b0 = b0; Z.nb = nb; Z.busy = true;
    cudaStream_t s = Z.s;
    EventTimer& tm = *Z.tm;
    int* const d_ys = Z.d_ys; double* const d_mu = Z.d_mu; double* const d_mom = Z.d_mom; GeneStats* const d_stats = Z.d_stats;
    GeneOutDev* const d_out = Z.d_out; GeneOutDev* const h_out = Z.h_out; double* const d_s1 = Z.d_s1; double* const d_q = Z.d_q;
    int* const d_tail = Z.d_tail; int* const d_big = Z.d_big; int* const d_big_tmp = Z.d_big_tmp; double* const d_okeys = Z.d_okeys; int* const d_order = Z.d_order;
    CompGeneDev* const d_cg = Z.d_cg; double* const d_off_s = Z.d_off_s; double* const d_off_q = Z.d_off_q; OffGeneDev* const d_offg = Z.d_offg;
    int* const d_counters = Z.d_counters; double* const d_a0p = Z.d_a0p; double* const d_b0p = Z.d_b0p; double* const d_ex_scratch = Z.d_ex_scratch;
    tm.mark(SCLANE_STAGE_COPY);             // on the compute stream this interval is the EXPOSED wait for the batch's H2D
    const int* raw = nullptr;
    H2DFeeder::Layout lay;                          // device-resident counts: one int32 part
    lay.nA = 0; lay.a16 = 0;
    if (h_counts) {
      lay = feeder->wait(kb);
      CK(cudaStreamWaitEvent(s, feeder->ready(kb), 0));
      raw = feeder->raw(kb);
    } else {
      raw = d_counts + (size_t)b0 * (size_t)n_cells;
    }
    tm.mark(SCLANE_STAGE_GATHER);
    // genes [gene0, gene0 + cnt) of the batch, whose rows start at `base` as uint16 or int32
    auto gather_part = [&](const void* base, bool u16, int gene0, int cnt) {
      int* ys_p = d_ys + (size_t)gene0 * (size_t)L.Npad;
      if (use_comp) {
        // compressed route: plain permutation copy, K CTAs per gene; k_aggregate derives the per-gene constants
        int K = (int)((L.N + 8191) / 8192);
        if (K < 1) K = 1;
        if (K > 32) K = 32;
        while (K > 1 && (((L.Npad / K + 127) / 128) * 128) * (K - 1) >= L.Npad) --K;      // no empty trailing part
        if (u16) k_gather_copy<unsigned short><<<(unsigned)((size_t)cnt * K), kThreads, 0, s>>>((const unsigned short*)base, (long long)n_cells, D.perm, L.N, L.Npad, ys_p, cnt, K);
        else k_gather_copy<int><<<(unsigned)((size_t)cnt * K), kThreads, 0, s>>>((const int*)base, (long long)n_cells, D.perm, L.N, L.Npad, ys_p, cnt, K);
      } else {
        if (u16) k_gather_stats<unsigned short><<<cnt, kThreads, 0, s>>>((const unsigned short*)base, (long long)n_cells, D.perm, L.N, L.Npad, ys_p, d_stats + gene0, d_out + gene0, cnt);
        else k_gather_stats<int><<<cnt, kThreads, 0, s>>>((const int*)base, (long long)n_cells, D.perm, L.N, L.Npad, ys_p, d_stats + gene0, d_out + gene0, cnt);
      }
      CK(cudaGetLastError());
      g_launches.fetch_add(1);
      prof.launches[SCLANE_STAGE_GATHER]++;
    };
    if (lay.nA > 0) gather_part(raw, lay.a16 != 0, 0, (int)lay.nA);
    if (lay.nA < nb) gather_part(reinterpret_cast<const char*>(raw) + (size_t)lay.nA * (size_t)n_cells * 4, false, (int)lay.nA, nb - (int)lay.nA);
    if (h_counts) feeder->consumed(kb, s);
    StageArgs A;
    A.lineage = D.lin; A.ys = d_ys; A.u = D.u; A.off = D.off; A.mu = d_mu; A.stats = d_stats; A.out = d_out;
    A.scores = nullptr; A.mom = d_mom; A.n_genes = nb; A.M = o.M; A.tols_score = o.tols_score; A.use_offset = use_offset ? 1 : 0;
    A.order = nullptr; A.cg = nullptr; A.tail = nullptr; A.big = nullptr; A.exact = 0; A.pstart = nullptr;
    A.on = nullptr; A.oord = nullptr; A.off_ord = nullptr; A.ogrp = nullptr; A.ht_scratch = nullptr;
    CompArgs CA;
    CA.order = nullptr;
    CA.theta_nodes = 0;
    WgArgs WA;
    WA.on = D.on; WA.off_s = d_off_s; WA.off_q = d_off_q; WA.offg = d_offg; WA.counter = d_counters;
    auto wg_grid = [&](int n) { const int want = (n + kWgWarps - 1) / kWgWarps; return want < wg_slots ? want : wg_slots; };
    if (use_comp) {
      CA.lineage = D.lin; CA.pstart = D.pstart; CA.pn = D.pn; CA.pu = D.pu; CA.pb = D.pb; CA.pbstart = D.pbstart; CA.D = P.D; CA.Dpad = Dpad; CA.ys = d_ys;
      CA.s1 = d_s1; CA.q = d_q; CA.tail = d_tail; CA.big = d_big; CA.big_tmp = d_big_tmp; CA.cg = d_cg; CA.stats = d_stats; CA.out = d_out; CA.n_genes = nb; CA.M = o.M;
      CA.n_iter = n_iter; CA.tols_score = o.tols_score; CA.compute_stats = 1; CA.a0p = d_a0p; CA.b0p = d_b0p;
      CA.bn = d_bn; CA.bn_s = Z.d_bn_s; CA.bn_q = Z.d_bn_q;
      tm.mark(SCLANE_STAGE_AGGREGATE);
      k_aggregate<<<nb, kThreads, 0, s>>>(CA);
      CK(cudaGetLastError());
      g_launches.fetch_add(1);
      prof.launches[SCLANE_STAGE_AGGREGATE]++;
      if (d_bn) {
        const long long units = (long long)nb * (L.T + 1);
        k_gene_bucket_nodes<<<(unsigned)((units + kWarps - 1) / kWarps), kThreads, 0, s>>>(CA);
        CK(cudaGetLastError());
        g_launches.fetch_add(1);
        prof.launches[SCLANE_STAGE_AGGREGATE]++;
      }
      tm.mark(SCLANE_STAGE_COMP_FWD);
      if (exact) {
        // forward + backward pass (+ the final fits when there is no offset) of every gene, per-gene buckets
        CK(cudaMemsetAsync(d_counters, 0, 8 * sizeof(int), s));
        ExactArgs EA;
        EA.C = CA; EA.px = D.px; EA.cand = D.cand; EA.n_cand = (int)P.cand.size(); EA.scratch = d_ex_scratch; EA.counter = d_counters + 4;
        EA.do_final = (!use_offset && !o.basis_only) ? 1 : 0;
        k_exact<<<nb < ex_slots ? nb : ex_slots, kCompThreads, sizeof(ExactShared), s>>>(EA);
      } else if (use_wg) {
        CK(cudaMemsetAsync(d_counters, 0, 8 * sizeof(int), s));
        WA.C = CA; WA.counter = d_counters + 0;
        k_wg_stage<WG_FORWARD><<<wg_grid(nb), kWgThreads, 0, s>>>(WA);
      } else k_comp_stage<CSTAGE_FORWARD><<<nb, kCompThreads, 0, s>>>(CA);
      CK(cudaGetLastError());
      g_launches.fetch_add(1);
      prof.launches[SCLANE_STAGE_COMP_FWD]++;
    }
    // per-cell kernels: every gene when the compressed path is off, otherwise only the genes k_aggregate routed away
    // (counts >= kHistCap); genes that already finished a stage return at once
    for (int it = 0; it < n_iter && !exact; ++it) {
      tm.mark(SCLANE_STAGE_FWD_FIT);
      launch_stage<STAGE_FWD_FIT>(A, s);
      prof.launches[SCLANE_STAGE_FWD_FIT]++;
      launch_sweep(A, it == 0, s, &tm, &prof);
    }
    // backward stage: widest forward models first (same reason as the final stage's order below)
    tm.mark(SCLANE_STAGE_SELECT);
    k_order_final<<<(nb + 255) / 256, 256, 0, s>>>(d_out, nb, d_okeys, d_order, 2);
    k_order_final<<<(nb + 255) / 256, 256, 0, s>>>(d_out, nb, d_okeys, d_order, 1);
    CK(cudaGetLastError());
    g_launches.fetch_add(2);
    prof.launches[SCLANE_STAGE_SELECT] += 2;
    A.order = d_order;
    CA.order = d_order;
    if (use_comp && !exact) {
      tm.mark(SCLANE_STAGE_COMP_BWD);
      if (use_wg) {
        WA.C = CA; WA.counter = d_counters + 1;
        k_wg_stage<WG_BACKWARD><<<wg_grid(nb), kWgThreads, 0, s>>>(WA);
      } else k_comp_stage<CSTAGE_BACKWARD><<<nb, kCompThreads, 0, s>>>(CA);
      CK(cudaGetLastError());
      g_launches.fetch_add(1);
      prof.launches[SCLANE_STAGE_COMP_BWD]++;
    }
    if (!exact) {
      tm.mark(SCLANE_STAGE_BACKWARD);
      launch_stage<STAGE_BACKWARD>(A, s);
      prof.launches[SCLANE_STAGE_BACKWARD]++;
    }
    if (!o.basis_only) {
    if (use_offq) {
      // offsets: node weights of every gene, then the intercept-only fits on the nodes (warp per gene); the per-cell kernel
      // below then only fits the alternatives that kept a hinge
      tm.mark(SCLANE_STAGE_OFF_AGG);
      OffAggArgs OA;
      OA.lineage = D.lin; OA.on = D.on; OA.oord = D.oord; OA.oxi_ord = D.oxi_ord; OA.off_ord = D.off_ord; OA.ys = d_ys; OA.cg = d_cg;
      OA.off_s = d_off_s; OA.off_q = d_off_q; OA.offg = d_offg; OA.n_genes = nb;
      k_off_aggregate<<<nb, kThreads, 0, s>>>(OA);
      CK(cudaGetLastError());
      g_launches.fetch_add(1);
      prof.launches[SCLANE_STAGE_OFF_AGG]++;
      tm.mark(SCLANE_STAGE_OFF_NULL);
      if (!use_wg) CK(cudaMemsetAsync(d_counters, 0, 8 * sizeof(int), s));      // (the warp-per-gene forward launch zeroes them otherwise)
      WA.C = CA; WA.C.order = nullptr; WA.counter = d_counters + 3;
      k_wg_stage<WG_OFFNULL><<<wg_grid(nb), kWgThreads, 0, s>>>(WA);
      CK(cudaGetLastError());
      g_launches.fetch_add(1);
      prof.launches[SCLANE_STAGE_OFF_NULL]++;
    }
    // final stage: expensive (near-Poisson) genes first, so that no slow CTA is left running alone at the end of the launch
    tm.mark(SCLANE_STAGE_SELECT);
    k_order_final<<<(nb + 255) / 256, 256, 0, s>>>(d_out, nb, d_okeys, d_order, use_offset ? 3 : 0);
    k_order_final<<<(nb + 255) / 256, 256, 0, s>>>(d_out, nb, d_okeys, d_order, 1);
    CK(cudaGetLastError());
    g_launches.fetch_add(2);
    prof.launches[SCLANE_STAGE_SELECT] += 2;
    A.order = d_order;
    CA.order = d_order;
    if (use_comp && !use_offset && !exact) {
      tm.mark(SCLANE_STAGE_COMP_FINAL);
      if (use_wg) {
        WA.C = CA; WA.counter = d_counters + 2;
        k_wg_stage<WG_FINAL><<<wg_grid(nb), kWgThreads, 0, s>>>(WA);
      } else {
        static std::once_flag tn_once[SCLANE_MAX_GPUS];
        std::call_once(tn_once[device % SCLANE_MAX_GPUS], [] { CK(cudaFuncSetAttribute(k_comp_stage<CSTAGE_FINAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ThetaNodes))); });
        CA.theta_nodes = theta_nodes_enabled() ? 1 : 0;
        k_comp_stage<CSTAGE_FINAL><<<nb, kCompThreads, CA.theta_nodes ? sizeof(ThetaNodes) : 0, s>>>(CA);
      }
      CK(cudaGetLastError());
      g_launches.fetch_add(1);
      prof.launches[SCLANE_STAGE_COMP_FINAL]++;
    }
    tm.mark(SCLANE_STAGE_FINAL);
    if (use_offq) {
      // alternatives with hinges: hybrid per-cell passes (mu-free terms from the histogram); genes without a histogram
      // (route 0) are left to the plain per-cell launch below, which skips everything already done
      A.cg = d_cg; A.tail = d_tail; A.big = d_big; A.exact = exact ? 1 : 0; A.pstart = D.pstart;
      if (hyb_theta_nodes) { A.on = D.on; A.oord = D.oord; A.off_ord = D.off_ord; A.ogrp = D.ogrp; A.ht_scratch = Z.d_ht; }
      switch (final_cluster_size(P.L.N)) {
        case 8: launch_final_cluster<8>(A, s); break;
        case 4: launch_final_cluster<4>(A, s); break;
        case 2: launch_final_cluster<2>(A, s); break;
        default: launch_stage<STAGE_FINAL, true>(A, s);
      }
      prof.launches[SCLANE_STAGE_FINAL]++;
    }
    launch_stage<STAGE_FINAL>(A, s);
    prof.launches[SCLANE_STAGE_FINAL]++;
    }
    tm.mark(SCLANE_STAGE_COPY);
    CK(cudaMemcpyAsync(h_out, d_out, (size_t)nb * sizeof(GeneOutDev), cudaMemcpyDeviceToHost, s));
    prof.d2h_bytes += (double)nb * sizeof(GeneOutDev);
    prof.launches[SCLANE_STAGE_COPY]++;
    tm.mark(-1);
    g_trace.at("batch queued");
  }
  // drain: oldest batch first
  {
    std::vector<Slot*> left;
    for (auto& Z : slots) if (Z.busy) left.push_back(&Z);
    std::sort(left.begin(), left.end(), [](const Slot* a, const Slot* b) { return a->b0 < b->b0; });
    for (Slot* Z : left) finish(*Z);
  }
  if (feeder) { prof.h2d_bytes += feeder->h2d_bytes(); prof.launches[SCLANE_STAGE_COPY] += feeder->copies(); }
}

// GEE mode: genes [g0, g1) of one lineage on one device.  Same batching as run_shard; one k_gee launch per batch.
void run_shard_gee(int device, const int32_t* h_counts, int64_t n_cells, int64_t g0, int64_t g1, const LineagePlan& P,
                   const GeePlan& GP, const sclane_opts& o, bool use_offset, sclane_record* out, int n_lineages, int lin_idx,
                   sclane_profile& prof) {
  if (g1 <= g0) return;
  CK(cudaSetDevice(device));
  DevicePool& pool = pool_for(device);
  if (!pool.stream) CK(cudaStreamCreateWithFlags(&pool.stream, cudaStreamNonBlocking));
  cudaStream_t s = pool.stream;
  const Lineage& L = P.L;
  int* d_cells = (int*)pool.get("gee_cells", GP.cells.size() * sizeof(int));
  unsigned char* d_bid = (unsigned char*)pool.get("gee_bid", GP.bid.size());
  double* d_u = (double*)pool.get("gee_u", GP.u.size() * sizeof(double));
  Lineage* d_lin = (Lineage*)pool.get("lineage", sizeof(Lineage));
  GeeLineage* d_glin = (GeeLineage*)pool.get("gee_lineage", sizeof(GeeLineage));
  double* d_off = nullptr;
  CK(cudaMemcpyAsync(d_cells, GP.cells.data(), GP.cells.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_bid, GP.bid.data(), GP.bid.size(), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_u, GP.u.data(), GP.u.size() * sizeof(double), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_lin, &L, sizeof(Lineage), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_glin, &GP.GL, sizeof(GeeLineage), cudaMemcpyHostToDevice, s));
  if (!GP.off.empty()) {
    d_off = (double*)pool.get("gee_off", GP.off.size() * sizeof(double));
    CK(cudaMemcpyAsync(d_off, GP.off.data(), GP.off.size() * sizeof(double), cudaMemcpyHostToDevice, s));
  }
  static std::once_flag attr_once[SCLANE_MAX_GPUS];
  std::call_once(attr_once[device % SCLANE_MAX_GPUS], [] {
    CK(cudaFuncSetAttribute(k_gee, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GeeShared)));
  });
  const int64_t G = g1 - g0;
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  const size_t per_gene = (size_t)n_cells * 4 + (size_t)L.Npad * 4 + sizeof(GeeGeneOutDev) + sizeof(GeneStats);
  size_t budget = (size_t)((double)(free_b + pool.held()) * 0.6);
  const size_t cap = (size_t)24 << 30;
  if (budget > cap) budget = cap;
  int64_t B = (int64_t)(budget / per_gene);
  // a few batches (>= 1024 genes each) so that the H2D copy of batch k + 1 runs on the copy stream under k_gee of batch k
  if ((double)G * (double)n_cells * 4.0 >= 128e6) { const int64_t want = std::max<int64_t>(1024, (G + 3) / 4); if (want < B) B = want; }
  if (o.genes_per_batch > 0 && o.genes_per_batch < B) B = o.genes_per_batch;
  if (B > G) B = G;
  if (B < 1) throw CudaFail{"not enough free device memory for a single gene"};
  const int64_t n_batches = (G + B - 1) / B;
  if (!pool.copy_stream) CK(cudaStreamCreateWithFlags(&pool.copy_stream, cudaStreamNonBlocking));
  cudaStream_t cs = pool.copy_stream;
  int* d_raw2[2];
  d_raw2[0] = (int*)pool.get("raw", (size_t)B * (size_t)n_cells * 4);
  d_raw2[1] = n_batches > 1 ? (int*)pool.get("raw_b", (size_t)B * (size_t)n_cells * 4) : d_raw2[0];
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  for (int i = 0; i < 2; ++i) {
    CK(cudaEventCreateWithFlags(&ev_h2d[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev_free[i], cudaEventDisableTiming));
  }
  struct EventGuard {
    cudaEvent_t* a; cudaEvent_t* b;
    ~EventGuard() { for (int i = 0; i < 2; ++i) { if (a[i]) cudaEventDestroy(a[i]); if (b[i]) cudaEventDestroy(b[i]); } }
  } ev_guard{ev_h2d, ev_free};
  auto issue_h2d = [&](int64_t k) {
    const int buf = (int)(k & 1);
    const int64_t kb0 = g0 + k * B, kb1 = std::min<int64_t>(kb0 + B, g1);
    if (k >= 2) CK(cudaStreamWaitEvent(cs, ev_free[buf], 0));          // the gather of batch k - 2 has consumed this buffer
    CK(cudaMemcpyAsync(d_raw2[buf], h_counts + (size_t)kb0 * (size_t)n_cells, (size_t)(kb1 - kb0) * (size_t)n_cells * 4,
                       cudaMemcpyHostToDevice, cs));
    CK(cudaEventRecord(ev_h2d[buf], cs));
    prof.h2d_bytes += (double)(kb1 - kb0) * (double)n_cells * 4.0;
    prof.launches[SCLANE_STAGE_COPY]++;
  };
  CK(cudaStreamSynchronize(s));                                         // plan tables are in place before anything overlaps
  issue_h2d(0);
  int* d_ys = (int*)pool.get("ys", (size_t)B * (size_t)L.Npad * 4);
  GeneStats* d_stats = (GeneStats*)pool.get("stats", (size_t)B * sizeof(GeneStats));
  GeeGeneOutDev* d_out = (GeeGeneOutDev*)pool.get("gee_out", (size_t)B * sizeof(GeeGeneOutDev));
  GeeGeneOutDev* h_out = (GeeGeneOutDev*)pool.get("gee_h_out", (size_t)B * sizeof(GeeGeneOutDev), true);
  EventTimer tm(s);
  for (int64_t b0 = g0, kb = 0; b0 < g1; b0 += B, ++kb) {
    const int64_t b1 = std::min<int64_t>(b0 + B, g1);
    const int nb = (int)(b1 - b0);
    tm.mark(SCLANE_STAGE_COPY);             // exposed wait for the batch's H2D
    CK(cudaStreamWaitEvent(s, ev_h2d[kb & 1], 0));
    const int* d_raw = d_raw2[kb & 1];
    tm.mark(SCLANE_STAGE_GATHER);
    k_gather_stats<int><<<nb, kThreads, 0, s>>>(d_raw, (long long)n_cells, d_cells, L.N, L.Npad, d_ys, d_stats, nullptr, nb);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ev_free[kb & 1], s));
    g_launches.fetch_add(1);
    prof.launches[SCLANE_STAGE_GATHER]++;
    GeeArgs A;
    A.lineage = d_lin; A.glin = d_glin; A.ys = d_ys; A.bid = d_bid; A.u = d_u; A.off = d_off; A.stats = d_stats; A.out = d_out;
    A.n_genes = nb; A.M = o.M; A.use_offset = use_offset ? 1 : 0; A.cor = o.gee_cor; A.gee_test = o.gee_test; A.sandwich = o.gee_bias_correction != 0 ? 1 : 0; A.tols_score = o.tols_score; A.mean_off = GP.mean_off;
    tm.mark(SCLANE_STAGE_GEE);
    k_gee<<<nb, kThreads, sizeof(GeeShared), s>>>(A);
    CK(cudaGetLastError());
    g_launches.fetch_add(1);
    prof.launches[SCLANE_STAGE_GEE]++;
    tm.mark(SCLANE_STAGE_COPY);
    CK(cudaMemcpyAsync(h_out, d_out, (size_t)nb * sizeof(GeeGeneOutDev), cudaMemcpyDeviceToHost, s));
    prof.d2h_bytes += (double)nb * sizeof(GeeGeneOutDev);
    prof.launches[SCLANE_STAGE_COPY]++;
    tm.mark(-1);
    if (kb + 1 < n_batches) issue_h2d(kb + 1);                // next batch's transfer while the GPU works on this one
    CK(cudaStreamSynchronize(s));
    tm.collect(prof);
    for (int i = 0; i < nb; ++i) {
      GeeGeneOut go;
      go.st = h_out[i].st; go.alt = h_out[i].alt; go.null = h_out[i].null;
      finalize_record_gee(go, L, o.gee_test, o.gee_bias_correction, GP.GL.S, out[(b0 + i) * n_lineages + lin_idx]);
    }
  }
}

void launch_sweep(const StageArgs& A, bool first, cudaStream_t s, EventTimer* tm, sclane_profile* prof) {
  if (tm) tm->mark(SCLANE_STAGE_SWEEP);
  if (first) k_sweep_moments<true><<<A.n_genes, kSweepThreads, 0, s>>>(A.lineage, A.ys, A.mu, A.u, A.out, A.mom, A.n_genes);
  else k_sweep_moments<false><<<A.n_genes, kSweepThreads, 0, s>>>(A.lineage, A.ys, A.mu, A.u, A.out, A.mom, A.n_genes);
  CK(cudaGetLastError());
  g_launches.fetch_add(1);
  if (prof) prof->launches[SCLANE_STAGE_SWEEP]++;
  if (tm) tm->mark(SCLANE_STAGE_SELECT);
  k_sweep_select<<<(A.n_genes + kSelectWarps - 1) / kSelectWarps, kSelectWarps * 32, 0, s>>>(A);
  CK(cudaGetLastError());
  g_launches.fetch_add(1);
  if (prof) prof->launches[SCLANE_STAGE_SELECT]++;
}

int check_opts(const sclane_opts* o, char* errbuf, size_t errlen) {
  if (!o) { set_err(errbuf, errlen, "opts is NULL"); return SCLANE_ERR_ARG; }
  if (o->abi_version != SCLANE_ABI_VERSION) { set_err(errbuf, errlen, "ABI version mismatch: caller %d, library %d", o->abi_version, SCLANE_ABI_VERSION); return SCLANE_ERR_ARG; }
  if (o->mode != 0 && o->mode != 1) { set_err(errbuf, errlen, "mode %d is not built; 0 = GLM, 1 = GEE", o->mode); return SCLANE_ERR_UNSUPPORTED; }
  if (o->M < 3 || o->M > SCLANE_PMAX) { set_err(errbuf, errlen, "M must be in [3, %d]", SCLANE_PMAX); return SCLANE_ERR_ARG; }
  if (o->h2d_mode < 0 || o->h2d_mode > 2) { set_err(errbuf, errlen, "h2d_mode must be 0, 1 or 2"); return SCLANE_ERR_ARG; }
  if (o->batch_streams < 0 || o->batch_streams > 4) { set_err(errbuf, errlen, "batch_streams must be in [0, 4]"); return SCLANE_ERR_ARG; }
  if (!o->approx_knot && o->mode != 0) { set_err(errbuf, errlen, "approx_knot = 0 is built for GLM mode only"); return SCLANE_ERR_UNSUPPORTED; }
  if (o->n_knot_max < 1 || o->n_knot_max > SCLANE_TMAX) { set_err(errbuf, errlen, "n_knot_max must be in [1, %d]", SCLANE_TMAX); return SCLANE_ERR_ARG; }
  return SCLANE_OK;
}

int usable_devices() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void merge_profile(const sclane_profile& p) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < SCLANE_NSTAGES; ++i) { g_prof.ms[i] += p.ms[i]; g_prof.launches[i] += p.launches[i]; }
  g_prof.sweep_bytes += p.sweep_bytes;
  g_prof.sweep_gene_iters += p.sweep_gene_iters;
  g_prof.h2d_bytes += p.h2d_bytes;
  g_prof.d2h_bytes += p.d2h_bytes;
}

// ---- sparse ingestion: dgCMatrix (genes x cells, compressed by cell) -> dense gene-major int32 shard on the device ----
struct CscInput {
  const int32_t* indptr;     // [n_cells + 1]   dgCMatrix@p
  const int32_t* indices;    // [nnz] gene row  dgCMatrix@i
  const double* data;        // [nnz]           dgCMatrix@x
  int64_t nnz;
};

// one thread per cell column: scatter its non-zeros that fall into [g0, g1) into the dense shard
__global__ void k_csc_expand(const int* __restrict__ indptr, const int* __restrict__ indices, const double* __restrict__ data,
                             long long n_cells, int g0, int g1, int* __restrict__ dense) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_cells) return;
  const int k1 = indptr[j + 1];
  for (int k = indptr[j]; k < k1; ++k) {
    const int g = indices[k];
    if (g >= g0 && g < g1) dense[(size_t)(g - g0) * (size_t)n_cells + (size_t)j] = (int)llrint(data[k]);
  }
}

const int32_t* expand_csc_shard(int device, const CscInput& c, int64_t n_cells, int64_t g0, int64_t g1, double& h2d_bytes) {
  CK(cudaSetDevice(device));
  DevicePool& pool = pool_for(device);
  if (!pool.stream) CK(cudaStreamCreateWithFlags(&pool.stream, cudaStreamNonBlocking));
  cudaStream_t s = pool.stream;
  int* d_ptr = (int*)pool.get("csc_indptr", (size_t)(n_cells + 1) * 4);
  int* d_idx = (int*)pool.get("csc_indices", (size_t)c.nnz * 4);
  double* d_val = (double*)pool.get("csc_data", (size_t)c.nnz * 8);
  int* dense = (int*)pool.get("csc_dense", (size_t)(g1 - g0) * (size_t)n_cells * 4);
  CK(cudaMemcpyAsync(d_ptr, c.indptr, (size_t)(n_cells + 1) * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_idx, c.indices, (size_t)c.nnz * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_val, c.data, (size_t)c.nnz * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(dense, 0, (size_t)(g1 - g0) * (size_t)n_cells * 4, s));
  k_csc_expand<<<(unsigned)((n_cells + 255) / 256), 256, 0, s>>>(d_ptr, d_idx, d_val, (long long)n_cells, (int)g0, (int)g1, dense);
  CK(cudaGetLastError());
  g_launches.fetch_add(1);
  CK(cudaStreamSynchronize(s));
  h2d_bytes = (double)(n_cells + 1) * 4.0 + (double)c.nnz * 12.0;
  return dense;
}

void drain_device(int device) {
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return; }
  DevicePool& pool = pool_for(device);
  if (pool.copy_stream) cudaStreamSynchronize(pool.copy_stream);
  if (pool.stream) cudaStreamSynchronize(pool.stream);
  for (auto e : pool.extra_streams) if (e) cudaStreamSynchronize(e);
  cudaGetLastError();
}

int test_dynamic_impl(const int32_t* h_counts, const int32_t* d_counts, int device_for_d, int64_t n_cells, int64_t n_genes,
                      const double* pt, int32_t n_lineages, const double* size_factor, const sclane_opts* opts,
                      sclane_record* out, sclane_lineage_info* lineages, char* errbuf, size_t errlen,
                      const int32_t* subject_id = nullptr, const CscInput* csc = nullptr) {
  int rc = check_opts(opts, errbuf, errlen);
  if (rc) return rc;
  const bool gee = opts->mode == 1;
  if (gee) {
    if (!subject_id || !h_counts) { set_err(errbuf, errlen, "GEE mode needs subject ids and host counts: call sclane_test_dynamic_gee()"); return SCLANE_ERR_ARG; }
    if (opts->M > SCLANE_GEE_PMAX) { set_err(errbuf, errlen, "GEE mode: M must be <= %d", SCLANE_GEE_PMAX); return SCLANE_ERR_ARG; }
    if (opts->gee_test != 0 && opts->gee_test != 1) { set_err(errbuf, errlen, "gee_test must be 0 (wald) or 1 (score)"); return SCLANE_ERR_ARG; }
    if (opts->gee_bias_correction < 0 || opts->gee_bias_correction > 2) { set_err(errbuf, errlen, "Unsupported bias correction method in waldTestGEE()."); return SCLANE_ERR_ARG; }
    if (opts->gee_bias_correction != 0 && opts->gee_cor == SCLANE_COR_INDEPENDENCE) { set_err(errbuf, errlen, "Unrecognized correlation structure in biasCorrectGEE()."); return SCLANE_ERR_ARG; }
    if (opts->gee_cor < SCLANE_COR_INDEPENDENCE || opts->gee_cor > SCLANE_COR_AR1) { set_err(errbuf, errlen, "Currently unsupported correlation structure."); return SCLANE_ERR_ARG; }
  } else if (subject_id) { set_err(errbuf, errlen, "subject ids given but opts->mode is not 1 (GEE)"); return SCLANE_ERR_ARG; }
  if ((!h_counts && !d_counts && !csc) || !pt || !out || n_cells <= 0 || n_genes <= 0 || n_lineages <= 0) {
    set_err(errbuf, errlen, "bad arguments (null pointer or non-positive size)");
    return SCLANE_ERR_ARG;
  }
  if (n_cells > (int64_t)2000000000) { set_err(errbuf, errlen, "n_cells too large"); return SCLANE_ERR_ARG; }
  const int ndev = usable_devices();
  if (ndev <= 0) { set_err(errbuf, errlen, "no usable CUDA device (this library has no CPU fallback)"); return SCLANE_ERR_CUDA; }
  std::vector<int> devs;
  if (d_counts) devs.push_back(device_for_d);
  else if (opts->n_gpus > 0) { for (int i = 0; i < opts->n_gpus && i < SCLANE_MAX_GPUS; ++i) devs.push_back(opts->gpu_ids[i]); }
  else { for (int i = 0; i < ndev; ++i) devs.push_back(i); }
  for (int d : devs) if (d < 0 || d >= ndev) { set_err(errbuf, errlen, "device id %d out of range (0..%d)", d, ndev - 1); return SCLANE_ERR_ARG; }
  const auto t_start = std::chrono::steady_clock::now();
  g_trace.start();
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::memset(&g_prof, 0, sizeof(g_prof));
  }
  // sparse input: every device expands ITS gene shard to the dense gene-major layout once (all lineages reuse it)
  std::vector<const int32_t*> shard_base(devs.size(), nullptr);
  if (csc) {
    const int nd0 = (int)devs.size();
    std::vector<std::string> errs0((size_t)nd0);
    std::vector<double> h2d0((size_t)nd0, 0.0);
    auto expand = [&](int di) {
      const int64_t g0 = n_genes * di / nd0, g1 = n_genes * (di + 1) / nd0;
      try { shard_base[(size_t)di] = expand_csc_shard(devs[(size_t)di], *csc, n_cells, g0, g1, h2d0[(size_t)di]); }
      catch (const CudaFail& f) { errs0[(size_t)di] = f.msg; }
      catch (const std::exception& e) { errs0[(size_t)di] = e.what(); }
      catch (...) { errs0[(size_t)di] = "unknown exception"; }
    };
    if (nd0 == 1) expand(0);
    else { ThreadGroup tg; for (int di = 0; di < nd0; ++di) tg.run(expand, di); tg.join(); }
    for (int di = 0; di < nd0; ++di) {
      if (!errs0[(size_t)di].empty()) { set_err(errbuf, errlen, "device %d: %s", devs[(size_t)di], errs0[(size_t)di].c_str()); return SCLANE_ERR_CUDA; }
      std::lock_guard<std::mutex> lk(g_prof_mu);
      g_prof.h2d_bytes += h2d0[(size_t)di];
    }
  }
  // every lineage's plan is built on a background thread, all started now: lineage 0's overlaps the first H2D, the others
  // are ready long before their turn
  // plan objects are pooled across calls (calls are serialised by the library lock): their vectors keep their capacity
  static std::vector<LineagePlan> plans;
  if (plans.size() < (size_t)n_lineages) plans.resize((size_t)n_lineages);
  std::vector<std::shared_future<bool>> plans_ready((size_t)n_lineages);
  // ... unless the plan can be built on the device (K0 kernels, sclane_k0_device.cuh): then the host only scans the column once
  std::vector<K0Host> k0s((size_t)n_lineages);
  for (int l = 0; l < n_lineages; ++l) {
    LineagePlan* Pl = &plans[(size_t)l];
    if (!gee) k0s[(size_t)l] = k0_prescan(pt + (int64_t)l * n_cells, n_cells, *opts, size_factor != nullptr);
    if (k0s[(size_t)l].eligible) { std::promise<bool> pr; pr.set_value(true); plans_ready[(size_t)l] = pr.get_future().share(); Pl->error.clear(); continue; }
    plans_ready[(size_t)l] = std::async(std::launch::async, [Pl, pt, l, n_cells, size_factor, opts]() {
      return build_lineage_plan(pt + (int64_t)l * n_cells, n_cells, size_factor, *opts, *Pl);
    }).share();
  }
  for (int l = 0; l < n_lineages; ++l) {
    LineagePlan& P = plans[(size_t)l];
    std::shared_future<bool> plan_ready = plans_ready[(size_t)l];
    GeePlan GPl;
    if (gee) {
      if (!plan_ready.get()) { set_err(errbuf, errlen, "lineage %d: %s", l, P.error.c_str()); return SCLANE_ERR_ARG; }
      if (!build_gee_plan(P, pt + (int64_t)l * n_cells, n_cells, subject_id, size_factor, GPl)) {
        set_err(errbuf, errlen, "lineage %d: %s", l, GPl.error.c_str());
        return SCLANE_ERR_ARG;
      }
    }
    const int nd = (int)devs.size();
    std::vector<std::string> errs((size_t)nd);
    std::vector<sclane_profile> profs((size_t)nd);
    for (auto& p : profs) std::memset(&p, 0, sizeof(p));
    auto work = [&](int di) {
      // contiguous gene blocks per device (gene-major matrix => contiguous memory per shard)
      const int64_t g0 = n_genes * di / nd, g1 = n_genes * (di + 1) / nd;
      try {
        if (gee) run_shard_gee(devs[(size_t)di], h_counts, n_cells, g0, g1, P, GPl, *opts, size_factor != nullptr, out, n_lineages, l, profs[(size_t)di]);
        else {
          // a CSC shard is dense on the device with local gene index 0 = global gene g0
          const int32_t* dc = csc ? shard_base[(size_t)di] - (ptrdiff_t)g0 * (ptrdiff_t)n_cells : d_counts;
          PlanSource PS;
          PS.plan = &P; PS.ready = plan_ready;
          if (k0s[(size_t)l].eligible) { PS.k0 = &k0s[(size_t)l]; PS.pt_col = pt + (int64_t)l * n_cells; PS.header = &P; PS.publish = g0 == 0 && g1 > 0; }
          run_shard(devs[(size_t)di], h_counts, dc, n_cells, g0, g1, PS, *opts, size_factor != nullptr, out, 0, n_lineages, l, profs[(size_t)di]);
        }
      } catch (const PlanFail&) {
        // reported once below, as an argument error
      } catch (const CudaFail& f) {
        errs[(size_t)di] = f.msg;
      } catch (const std::exception& e) {
        errs[(size_t)di] = e.what();
      } catch (...) {
        errs[(size_t)di] = "unknown exception";
      }
      // whatever happened, nothing of this shard may still be in flight (an async H2D out of the caller's buffer, a kernel
      // writing the pooled buffers) when the caller sees the return code
      drain_device(devs[(size_t)di]);
    };
    if (nd == 1) work(0);
    else {
      ThreadGroup tg;
      for (int di = 0; di < nd; ++di) tg.run(work, di);
      tg.join();
    }
    if (!plan_ready.get() || (k0s[(size_t)l].eligible && !P.error.empty())) { set_err(errbuf, errlen, "lineage %d: %s", l, P.error.c_str()); return SCLANE_ERR_ARG; }
    if (lineages) {
      sclane_lineage_info& li = lineages[l];
      std::memset(&li, 0, sizeof(li));
      li.n_cells = P.L.N; li.n_knots = P.L.T; li.minspan = P.minspan;
      for (int k = 0; k < P.L.T; ++k) li.knots[k] = P.L.knots[k];
      if (P.exact) {          // approx.knot = FALSE: n_knots = number of candidate abscissae, knots[] = the first SCLANE_TMAX of them
        li.n_knots = (int32_t)P.cand.size();
        for (size_t k = 0; k < P.cand.size() && k < (size_t)SCLANE_TMAX; ++k) li.knots[k] = P.px[(size_t)P.cand[k]];
      }
      li.pt_min = P.pt_min; li.pt_max = P.pt_max;
    }
    for (int di = 0; di < nd; ++di) {
      if (!errs[(size_t)di].empty()) {
        set_err(errbuf, errlen, "device %d: %s", devs[(size_t)di], errs[(size_t)di].c_str());
        return SCLANE_ERR_CUDA;
      }
      merge_profile(profs[(size_t)di]);
    }
  }
  const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
  g_trace.at("return");
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.wall_ms = wall;
  }
  const double per_gene = wall / 1000.0 / (double)(n_genes * n_lineages);
  for (int64_t i = 0; i < n_genes * n_lineages; ++i) out[i].gene_time = per_gene;
  return SCLANE_OK;
}

// ---- small helper kernels -------------------------------------------------------------------------
__global__ void k_matmul(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, int m, int k, int n) {
  // column-major: A m x k, B k x n, C m x n; 16x16 shared tiles
  __shared__ double As[16][17], Bs[16][17];
  const int row = blockIdx.x * 16 + threadIdx.x, col = blockIdx.y * 16 + threadIdx.y;
  double acc = 0.0;
  for (int t = 0; t < k; t += 16) {
    As[threadIdx.y][threadIdx.x] = (row < m && t + threadIdx.y < k) ? A[(size_t)(t + threadIdx.y) * m + row] : 0.0;
    Bs[threadIdx.y][threadIdx.x] = (t + threadIdx.x < k && col < n) ? B[(size_t)col * k + t + threadIdx.x] : 0.0;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 16; ++q) acc += As[q][threadIdx.x] * Bs[threadIdx.y][q];
    __syncthreads();
  }
  if (row < m && col < n) C[(size_t)col * m + row] = acc;
}

// Eigen-style unblocked LLT + solve against identity, one CTA, n <= 64, column-major in/out.
__global__ void k_llt_inverse(const double* __restrict__ A, double* __restrict__ Ainv, int n) {
  __shared__ double Lm[64][65];
  __shared__ int stop_at;
  const int tid = threadIdx.x;
  for (int i = tid; i < n * n; i += blockDim.x) { const int r = i % n, c = i / n; Lm[r][c] = A[(size_t)c * n + r]; }
  if (tid == 0) stop_at = -1;
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    if (tid == 0) {
      double x = Lm[k][k];
      for (int j = 0; j < k; ++j) x -= Lm[k][j] * Lm[k][j];
      if (x <= 0.0) stop_at = k;                 // Eigen: return k, matrix left as is from here on
      else Lm[k][k] = sqrt(x);
    }
    __syncthreads();
    if (stop_at >= 0) break;
    const double piv = Lm[k][k];
    for (int i = k + 1 + tid; i < n; i += blockDim.x) {
      double s = Lm[i][k];
      for (int j = 0; j < k; ++j) s -= Lm[i][j] * Lm[k][j];
      Lm[i][k] = s / piv;
    }
    __syncthreads();
  }
  // solve L y = e_c, L^T x = y for every column c (lower triangle only is read)
  for (int c = tid; c < n; c += blockDim.x) {
    double yv[64];
    for (int i = 0; i < n; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int j = 0; j < i; ++j) s -= Lm[i][j] * yv[j];
      yv[i] = s / Lm[i][i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = yv[i];
      for (int j = i + 1; j < n; ++j) s -= Lm[j][i] * yv[j];
      yv[i] = s / Lm[i][i];
    }
    for (int i = 0; i < n; ++i) Ainv[(size_t)c * n + i] = yv[i];
  }
}

// One-sided Jacobi SVD pseudo-inverse, one CTA; works on W = A (m x n, m >= n) in shared memory.
// out = V diag(1/s_i if s_i > tol) U^T  (n x m), column-major.
__global__ void k_pinv_tall(const double* __restrict__ A, double* __restrict__ out, int m, int n, double tol) {
  __shared__ double Wm[48][49];   // columns of A being orthogonalised: Wm[row][col]
  __shared__ double Vm[48][49];
  __shared__ double red[3];
  __shared__ int rotated;
  const int tid = threadIdx.x;
  for (int i = tid; i < m * n; i += blockDim.x) { const int r = i % m, c = i / m; Wm[r][c] = A[(size_t)c * m + r]; }
  for (int i = tid; i < n * n; i += blockDim.x) { const int r = i % n, c = i / n; Vm[r][c] = (r == c) ? 1.0 : 0.0; }
  __syncthreads();
  for (int sweep = 0; sweep < 60; ++sweep) {
    if (tid == 0) rotated = 0;
    __syncthreads();
    for (int p = 0; p < n - 1; ++p) {
      for (int q = p + 1; q < n; ++q) {
        if (tid == 0) {
          double a = 0.0, b = 0.0, g = 0.0;
          for (int r = 0; r < m; ++r) { a += Wm[r][p] * Wm[r][p]; b += Wm[r][q] * Wm[r][q]; g += Wm[r][p] * Wm[r][q]; }
          red[0] = a; red[1] = b; red[2] = g;
        }
        __syncthreads();
        const double a = red[0], b = red[1], g = red[2];
        const bool rot = fabs(g) > 1e-15 * sqrt(a * b) && g != 0.0;
        if (rot) {
          const double zeta = (b - a) / (2.0 * g);
          const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
          for (int r = tid; r < m; r += blockDim.x) {
            const double wp = Wm[r][p], wq = Wm[r][q];
            Wm[r][p] = cs * wp - sn * wq;
            Wm[r][q] = sn * wp + cs * wq;
          }
          for (int r = tid; r < n; r += blockDim.x) {
            const double vp = Vm[r][p], vq = Vm[r][q];
            Vm[r][p] = cs * vp - sn * vq;
            Vm[r][q] = sn * vp + cs * vq;
          }
          if (tid == 0) rotated = 1;
        }
        __syncthreads();
      }
    }
    if (!rotated) break;
    __syncthreads();
  }
  // singular values = column norms of W; pinv[i][r] = sum_c V[i][c] * W[r][c] / s_c^2  (U = W / s)
  __shared__ double sv[48];
  for (int c = tid; c < n; c += blockDim.x) {
    double a = 0.0;
    for (int r = 0; r < m; ++r) a += Wm[r][c] * Wm[r][c];
    sv[c] = sqrt(a);
  }
  __syncthreads();
  for (int idx = tid; idx < n * m; idx += blockDim.x) {
    const int i = idx % n, r = idx / n;
    double s = 0.0;
    for (int c = 0; c < n; ++c) if (sv[c] > tol) s += Vm[i][c] * Wm[r][c] / (sv[c] * sv[c]);
    out[(size_t)r * n + i] = s;
  }
}

struct PredModel {
  int n_alt;
  int kind[PMAX];
  double knot[PMAX];
  double coef[PMAX];
  double cov[PMAX * PMAX];   // row-major n_alt x n_alt
  int has_null;
  double null_coef, null_var;
};

// predict(type = "link", se.fit = TRUE) per cell (pullMARGESummary.R:66-69, pullNullSummary.R:69-72)
__global__ void k_link_preds(PredModel M, const double* __restrict__ x, const double* __restrict__ sf, long long n,
                             double* fit1, double* se1, double* fit0, double* se0) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double X = r_round(x[i], 1e4);                 // marge2.R:154
  const double off = sf ? log(1.0 / sf[i]) : 0.0;
  if (M.n_alt > 0) {
    double col[PMAX];
    double eta = 0.0;
    for (int j = 0; j < M.n_alt; ++j) {
      double v;
      if (M.kind[j] == SCLANE_TERM_INTERCEPT) v = 1.0;
      else if (M.kind[j] == SCLANE_TERM_TP1) v = X > M.knot[j] ? X - M.knot[j] : 0.0;
      else v = X < M.knot[j] ? M.knot[j] - X : 0.0;
      col[j] = v;
      eta += v * M.coef[j];
    }
    double q = 0.0;
    for (int j = 0; j < M.n_alt; ++j)
      for (int l = 0; l < M.n_alt; ++l) q += col[j] * M.cov[j * M.n_alt + l] * col[l];
    if (fit1) fit1[i] = eta + off;
    if (se1) se1[i] = sqrt(q);
  } else {
    if (fit1) fit1[i] = nan("");
    if (se1) se1[i] = nan("");
  }
  if (M.has_null) {
    if (fit0) fit0[i] = M.null_coef + off;
    if (se0) se0[i] = sqrt(M.null_var);
  } else {
    if (fit0) fit0[i] = nan("");
    if (se0) se0[i] = nan("");
  }
}

// smoothedCountsMatrix(): fitted means of every cell under every gene's final model (smoothedCountsMatrix.R:44-67)
struct SmoothModel {
  int n_alt;                 // 0 = model failed -> NaN column
  int kind[PMAX];
  double knot[PMAX];
  double coef[PMAX];
};
__global__ void k_smoothed_counts(const SmoothModel* __restrict__ models, const double* __restrict__ x, const double* __restrict__ sf,
                                  long long n, int add_offset, int log1p_norm, double* __restrict__ out) {
  __shared__ SmoothModel M;
  const int g = blockIdx.y;
  if (threadIdx.x == 0) M = models[g];
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = nan("");
  if (M.n_alt > 0) {
    const double X = r_round(x[i], 1e4);                 // marge2.R:154
    double eta = 0.0;
    for (int j = 0; j < M.n_alt; ++j) {
      double c;
      if (M.kind[j] == SCLANE_TERM_INTERCEPT) c = 1.0;
      else if (M.kind[j] == SCLANE_TERM_TP1) c = X > M.knot[j] ? X - M.knot[j] : 0.0;
      else c = X < M.knot[j] ? M.knot[j] - X : 0.0;
      eta += c * M.coef[j];
    }
    if (sf && add_offset) eta += log(1.0 / sf[i]);       // marge_link_fit of a GLM record includes the offset
    v = exp(eta);
    if (sf) v *= sf[i];
    if (log1p_norm) v = log1p(v);
  }
  out[(size_t)g * (size_t)n + (size_t)i] = v;
}

// explicit-knot lineage plan for the single-step entry points
bool plan_from_knots(const double* x, int64_t n, const double* knots, int n_knots, LineagePlan& P, char* errbuf, size_t errlen) {
  if (n_knots < 1 || n_knots > SCLANE_TMAX) { set_err(errbuf, errlen, "n_knots must be in [1, %d]", SCLANE_TMAX); return false; }
  for (int k = 1; k < n_knots; ++k) if (!(knots[k] >= knots[k - 1])) { set_err(errbuf, errlen, "knots must be ascending"); return false; }
  Lineage& L = P.L;
  std::memset(&L, 0, sizeof(L));
  L.N = (int)n;
  L.Npad = (int)((n + 127) / 128 * 128);
  L.T = n_knots;
  for (int k = 0; k < n_knots; ++k) L.knots[k] = knots[k];
  L.ref[0] = L.knots[0];
  for (int b = 1; b <= L.T; ++b) L.ref[b] = L.knots[b - 1];
  std::vector<double> X((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    if (std::isnan(x[i])) { set_err(errbuf, errlen, "x must not contain NaN"); return false; }
    X[(size_t)i] = r_round(x[i], 1e4);
  }
  std::vector<int32_t> order((size_t)n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return X[(size_t)a] < X[(size_t)b]; });
  P.perm.assign(order.begin(), order.end());
  P.u.assign((size_t)L.Npad, 0.0);
  P.off.clear();
  std::vector<int> cnt((size_t)L.T + 1, 0);
  const double* kb = L.knots;
  for (int64_t i = 0; i < n; ++i) {
    const double xv = X[(size_t)order[(size_t)i]];
    const int b = (int)(std::upper_bound(kb, kb + L.T, xv) - kb);
    cnt[(size_t)b]++;
    const double u = xv - L.ref[b];
    P.u[(size_t)i] = u;
    L.U0[b] += 1.0; L.U1[b] += u; L.U2[b] += u * u;
  }
  int acc = 0;
  for (int b = 0; b <= L.T; ++b) { L.bstart[b] = acc; acc += cnt[(size_t)b]; }
  for (int b = L.T + 1; b <= NBMAX; ++b) L.bstart[b] = acc;
  return true;
}

bool model_from_terms(const int32_t* kind, const int32_t* knot_idx, int n_terms, int n_knots, Model& m, char* errbuf, size_t errlen) {
  if (n_terms < 1 || n_terms > PMAX || kind[0] != SCLANE_TERM_INTERCEPT) { set_err(errbuf, errlen, "terms: 1..%d terms, intercept first", PMAX); return false; }
  m.n = n_terms;
  for (int j = 0; j < n_terms; ++j) {
    m.kind[j] = kind[j];
    m.knot[j] = j == 0 ? 0 : knot_idx[j];
    if (j > 0 && (kind[j] < 1 || kind[j] > 2 || knot_idx[j] < 0 || knot_idx[j] >= n_knots)) { set_err(errbuf, errlen, "bad term %d", j); return false; }
  }
  return true;
}

}  // namespace

// =====================================================================================================
// entry point bodies (the exported symbols at the end of the file wrap them in guarded(): lock + try/catch)
// =====================================================================================================
namespace impl {


void sclane_default_opts(sclane_opts* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->abi_version = SCLANE_ABI_VERSION;
  o->mode = 0;
  o->gee_cor = SCLANE_COR_AR1;
  o->M = 5;
  o->approx_knot = 1;
  o->n_knot_max = 25;
  o->minspan = -1;
  o->tols_score = 1e-5;
}

void sclane_get_profile(sclane_profile* out) {
  if (!out) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  *out = g_prof;
}

void sclane_release(void) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  for (auto& kv : g_pools) {
    if (cudaSetDevice(kv.first) == cudaSuccess) kv.second.release();
  }
  g_pools.clear();
}

int32_t sclane_test_dynamic(const int32_t* counts, int64_t n_cells, int64_t n_genes, const double* pt, int32_t n_lineages,
                            const double* size_factor, const sclane_opts* opts, sclane_record* out,
                            sclane_lineage_info* lineages, char* errbuf, size_t errlen) {
  return test_dynamic_impl(counts, nullptr, -1, n_cells, n_genes, pt, n_lineages, size_factor, opts, out, lineages, errbuf, errlen);
}

int32_t sclane_marge_basis(const int32_t* counts, int64_t n_cells, int64_t n_genes, const double* pt, int32_t n_lineages,
                           const sclane_opts* opts, sclane_record* out, sclane_lineage_info* lineages, char* errbuf, size_t errlen) {
  if (!opts) { set_err(errbuf, errlen, "opts is NULL"); return SCLANE_ERR_ARG; }
  if (opts->mode != 0) { set_err(errbuf, errlen, "sclane_marge_basis() runs the GLM forward / backward pass (opts->mode = 0)"); return SCLANE_ERR_ARG; }
  sclane_opts o = *opts;
  o.basis_only = 1;
  return test_dynamic_impl(counts, nullptr, -1, n_cells, n_genes, pt, n_lineages, nullptr, &o, out, lineages, errbuf, errlen);
}

int32_t sclane_test_dynamic_gee(const int32_t* counts, int64_t n_cells, int64_t n_genes, const double* pt, int32_t n_lineages,
                                const int32_t* subject_id, const double* size_factor, const sclane_opts* opts, sclane_record* out,
                                sclane_lineage_info* lineages, char* errbuf, size_t errlen) {
  if (!subject_id) { set_err(errbuf, errlen, "subject_id is NULL"); return SCLANE_ERR_ARG; }
  if (opts && opts->mode != 1) { set_err(errbuf, errlen, "sclane_test_dynamic_gee() needs opts->mode = 1"); return SCLANE_ERR_ARG; }
  return test_dynamic_impl(counts, nullptr, -1, n_cells, n_genes, pt, n_lineages, size_factor, opts, out, lineages, errbuf, errlen, subject_id);
}

int32_t sclane_test_dynamic_csc(const int32_t* indptr, const int32_t* indices, const double* data, int64_t n_cells, int64_t n_genes,
                                const double* pt, int32_t n_lineages, const double* size_factor, const sclane_opts* opts,
                                sclane_record* out, sclane_lineage_info* lineages, char* errbuf, size_t errlen) {
  if (!indptr || !indices || !data || n_cells <= 0) { set_err(errbuf, errlen, "bad arguments (null CSC pointer)"); return SCLANE_ERR_ARG; }
  if (opts && opts->mode != 0) { set_err(errbuf, errlen, "sparse input is built for GLM mode only"); return SCLANE_ERR_UNSUPPORTED; }
  if (indptr[0] != 0 || indptr[n_cells] < 0) { set_err(errbuf, errlen, "indptr must start at 0"); return SCLANE_ERR_ARG; }
  const int64_t nnz = indptr[n_cells];
  for (int64_t j = 0; j < n_cells; ++j) if (indptr[j + 1] < indptr[j]) { set_err(errbuf, errlen, "indptr must be non-decreasing"); return SCLANE_ERR_ARG; }
  for (int64_t k = 0; k < nnz; ++k) if (indices[k] < 0 || indices[k] >= n_genes) { set_err(errbuf, errlen, "row index %d out of range at entry %lld", indices[k], (long long)k); return SCLANE_ERR_ARG; }
  CscInput c{indptr, indices, data, nnz};
  return test_dynamic_impl(nullptr, nullptr, -1, n_cells, n_genes, pt, n_lineages, size_factor, opts, out, lineages, errbuf, errlen, nullptr, &c);
}

int32_t sclane_test_dynamic_device(const int32_t* d_counts, int32_t device, int64_t n_cells, int64_t n_genes, const double* pt,
                                   int32_t n_lineages, const double* size_factor, const sclane_opts* opts, sclane_record* out,
                                   sclane_lineage_info* lineages, char* errbuf, size_t errlen) {
  return test_dynamic_impl(nullptr, d_counts, device, n_cells, n_genes, pt, n_lineages, size_factor, opts, out, lineages, errbuf, errlen);
}

int32_t sclane_candidate_knots(const double* x, int64_t n, const sclane_opts* opts, double* knots_out, int32_t* minspan_out) {
  if (!x || !opts || !knots_out || n < 1) return -SCLANE_ERR_ARG;
  std::vector<double> xv(x, x + n), X;
  for (double v : xv) if (std::isnan(v)) return -SCLANE_ERR_ARG;
  int ms = 0;
  std::vector<double> k = candidate_knots(xv, opts->approx_knot != 0, opts->n_knot_max, opts->minspan, &X, &ms);
  if (minspan_out) *minspan_out = ms;
  if ((int)k.size() > SCLANE_TMAX) return -SCLANE_ERR_UNSUPPORTED;
  std::sort(k.begin(), k.end());
  for (size_t i = 0; i < k.size(); ++i) knots_out[i] = k[i];
  return (int32_t)k.size();
}

int32_t sclane_lineage_plan(const double* pt, int64_t n_cells, const sclane_opts* opts, int32_t on_device, int32_t device,
                            int32_t* perm, double* u, double* knots, int32_t* bstart, double* U, int32_t* pstart, double* pn,
                            double* pu, uint8_t* pb, int32_t* pbstart, int32_t* dims, char* errbuf, size_t errlen) {
  if (int rc = check_opts(opts, errbuf, errlen)) return rc;
  if (!pt || n_cells < 1 || n_cells > (int64_t)2000000000) { set_err(errbuf, errlen, "bad arguments"); return SCLANE_ERR_ARG; }
  auto put_lineage = [&](const Lineage& L, int D, int minspan) {
    const int T = L.T;
    if (knots) for (int k = 0; k < T; ++k) knots[k] = L.knots[k];
    if (bstart) for (int b = 0; b <= T + 1; ++b) bstart[b] = L.bstart[b];
    if (U) for (int b = 0; b <= T; ++b) { U[b] = L.U0[b]; U[(T + 1) + b] = L.U1[b]; U[2 * (T + 1) + b] = L.U2[b]; }
    if (dims) { dims[0] = L.N; dims[1] = T; dims[2] = D; dims[3] = minspan; }
  };
  if (!on_device) {
    LineagePlan P;
    if (!build_lineage_plan(pt, n_cells, nullptr, *opts, P)) { set_err(errbuf, errlen, "%s", P.error.c_str()); return SCLANE_ERR_ARG; }
    const int N = P.L.N, D = P.D, T = P.L.T;
    put_lineage(P.L, D, P.minspan);
    if (perm) std::memcpy(perm, P.perm.data(), (size_t)N * 4);
    if (u) std::memcpy(u, P.u.data(), (size_t)N * 8);
    if (pstart) std::memcpy(pstart, P.pstart.data(), ((size_t)D + 1) * 4);
    if (pn) std::memcpy(pn, P.pn.data(), (size_t)D * 8);
    if (pu) std::memcpy(pu, P.pu.data(), (size_t)D * 8);
    if (pb) std::memcpy(pb, P.pb.data(), (size_t)D);
    if (pbstart) for (int b = 0; b <= T + 1; ++b) pbstart[b] = P.pbstart[b];
    return SCLANE_OK;
  }
  const int ndev = usable_devices();
  if (ndev <= 0) { set_err(errbuf, errlen, "no usable CUDA device (this library has no CPU fallback)"); return SCLANE_ERR_CUDA; }
  if (device < 0 || device >= ndev) { set_err(errbuf, errlen, "device id %d out of range (0..%d)", device, ndev - 1); return SCLANE_ERR_ARG; }
  const K0Host H = k0_prescan(pt, n_cells, *opts, false);
  if (!H.eligible) { set_err(errbuf, errlen, "this input keeps the host plan (approx_knot = 0, host_plan set, fewer than 30 cells, or a pseudotime range beyond the radix key)"); return SCLANE_ERR_UNSUPPORTED; }
  CK(cudaSetDevice(device));
  DevicePool& pool = pool_for(device);
  if (!pool.stream) CK(cudaStreamCreateWithFlags(&pool.stream, cudaStreamNonBlocking));
  cudaStream_t s = pool.stream;
  LineagePlan hdr;
  DevicePlan D;
  try { device_k0(pool, s, pt, n_cells, H, *opts, hdr, D, nullptr); }
  catch (const PlanFail& f) { set_err(errbuf, errlen, "%s", f.msg.c_str()); return SCLANE_ERR_ARG; }
  const int N = hdr.L.N, Dn = hdr.D, T = hdr.L.T;
  put_lineage(hdr.L, Dn, hdr.minspan);
  if (perm) CK(cudaMemcpyAsync(perm, D.perm, (size_t)N * 4, cudaMemcpyDeviceToHost, s));
  if (u) CK(cudaMemcpyAsync(u, D.u, (size_t)N * 8, cudaMemcpyDeviceToHost, s));
  if (pstart) CK(cudaMemcpyAsync(pstart, D.pstart, ((size_t)Dn + 1) * 4, cudaMemcpyDeviceToHost, s));
  if (pn) CK(cudaMemcpyAsync(pn, D.pn, (size_t)Dn * 8, cudaMemcpyDeviceToHost, s));
  if (pu) CK(cudaMemcpyAsync(pu, D.pu, (size_t)Dn * 8, cudaMemcpyDeviceToHost, s));
  if (pb) CK(cudaMemcpyAsync(pb, D.pb, (size_t)Dn, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (pbstart) for (int b = 0; b <= T + 1; ++b) pbstart[b] = hdr.pbstart[b];
  return SCLANE_OK;
}

int32_t sclane_null_stats_glm(const int32_t* counts, int64_t n_cells, int64_t n_genes, const double* x, const double* knots,
                              int32_t n_knots, const int32_t* term_kind, const int32_t* term_knot_idx, int32_t n_terms,
                              int32_t device, double* beta_out, double* sigma_out, double* mu_out, char* errbuf, size_t errlen) {
  if (!counts || !x || !knots || !term_kind || !term_knot_idx || !beta_out || !sigma_out || n_cells < 1 || n_genes < 1) {
    set_err(errbuf, errlen, "bad arguments"); return SCLANE_ERR_ARG;
  }
  if (usable_devices() <= 0) { set_err(errbuf, errlen, "no usable CUDA device (this library has no CPU fallback)"); return SCLANE_ERR_CUDA; }
  LineagePlan P;
  Model m;
  if (!plan_from_knots(x, n_cells, knots, n_knots, P, errbuf, errlen)) return SCLANE_ERR_ARG;
  if (!model_from_terms(term_kind, term_knot_idx, n_terms, n_knots, m, errbuf, errlen)) return SCLANE_ERR_ARG;
  try {
    CK(cudaSetDevice(device));
    DevicePool& pool = pool_for(device);
    if (!pool.stream) CK(cudaStreamCreateWithFlags(&pool.stream, cudaStreamNonBlocking));
    cudaStream_t s = pool.stream;
    DevicePlan D;
    D.upload(P, pool, s);
    const Lineage& L = P.L;
    const int nb = (int)n_genes;
    int* d_raw = (int*)pool.get("raw", (size_t)nb * n_cells * 4);
    int* d_ys = (int*)pool.get("ys", (size_t)nb * L.Npad * 4);
    double* d_mu = (double*)pool.get("mu", (size_t)nb * L.Npad * 8);
    GeneStats* d_stats = (GeneStats*)pool.get("stats", (size_t)nb * sizeof(GeneStats));
    GeneOutDev* d_out = (GeneOutDev*)pool.get("out", (size_t)nb * sizeof(GeneOutDev));
    CK(cudaMemcpyAsync(d_raw, counts, (size_t)nb * n_cells * 4, cudaMemcpyHostToDevice, s));
    k_gather_stats<int><<<nb, kThreads, 0, s>>>(d_raw, (long long)n_cells, D.perm, L.N, L.Npad, d_ys, d_stats, d_out, nb);
    CK(cudaGetLastError()); g_launches.fetch_add(1);
    std::vector<GeneOutDev> h((size_t)nb);
    CK(cudaMemcpyAsync(h.data(), d_out, (size_t)nb * sizeof(GeneOutDev), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (auto& g : h) g.st.fwd = m;                       // preset basis; fwd_iters = 0 => cold start
    CK(cudaMemcpyAsync(d_out, h.data(), (size_t)nb * sizeof(GeneOutDev), cudaMemcpyHostToDevice, s));
    StageArgs A;
    A.lineage = D.lin; A.ys = d_ys; A.u = D.u; A.off = nullptr; A.mu = d_mu; A.stats = d_stats; A.out = d_out;
    A.scores = nullptr; A.mom = nullptr; A.n_genes = nb; A.M = 5; A.tols_score = 1e-5; A.use_offset = 0; A.order = nullptr; A.cg = nullptr; A.tail = nullptr; A.big = nullptr; A.exact = 0; A.pstart = nullptr;
    launch_stage<STAGE_FWD_FIT>(A, s);
    CK(cudaMemcpyAsync(h.data(), d_out, (size_t)nb * sizeof(GeneOutDev), cudaMemcpyDeviceToHost, s));
    std::vector<double> hmu;
    if (mu_out) { hmu.resize((size_t)nb * L.Npad); CK(cudaMemcpyAsync(hmu.data(), d_mu, hmu.size() * 8, cudaMemcpyDeviceToHost, s)); }
    CK(cudaStreamSynchronize(s));
    for (int g = 0; g < nb; ++g) {
      const bool bad = (h[(size_t)g].st.error != 0);
      for (int j = 0; j < n_terms; ++j) beta_out[(size_t)g * n_terms + j] = bad ? std::nan("") : h[(size_t)g].st.beta[j];
      sigma_out[g] = bad ? std::nan("") : h[(size_t)g].st.sigma;
      if (mu_out) for (int i = 0; i < L.N; ++i) mu_out[(size_t)g * n_cells + P.perm[(size_t)i]] = bad ? std::nan("") : hmu[(size_t)g * L.Npad + i];
    }
  } catch (const CudaFail& f) { set_err(errbuf, errlen, "%s", f.msg.c_str()); return SCLANE_ERR_CUDA; }
  return SCLANE_OK;
}

int32_t sclane_score_glm(const int32_t* counts, const double* mu, int64_t n_cells, int64_t n_genes, const double* x,
                         const double* knots, int32_t n_knots, const int32_t* term_kind, const int32_t* term_knot_idx,
                         int32_t n_terms, const double* sigma, int32_t device, double* scores_out, char* errbuf, size_t errlen) {
  if (!counts || !mu || !x || !knots || !term_kind || !term_knot_idx || !sigma || !scores_out || n_cells < 1 || n_genes < 1) {
    set_err(errbuf, errlen, "bad arguments"); return SCLANE_ERR_ARG;
  }
  if (usable_devices() <= 0) { set_err(errbuf, errlen, "no usable CUDA device (this library has no CPU fallback)"); return SCLANE_ERR_CUDA; }
  LineagePlan P;
  Model m;
  if (!plan_from_knots(x, n_cells, knots, n_knots, P, errbuf, errlen)) return SCLANE_ERR_ARG;
  if (!model_from_terms(term_kind, term_knot_idx, n_terms, n_knots, m, errbuf, errlen)) return SCLANE_ERR_ARG;
  try {
    CK(cudaSetDevice(device));
    DevicePool& pool = pool_for(device);
    if (!pool.stream) CK(cudaStreamCreateWithFlags(&pool.stream, cudaStreamNonBlocking));
    cudaStream_t s = pool.stream;
    DevicePlan D;
    D.upload(P, pool, s);
    const Lineage& L = P.L;
    const int nb = (int)n_genes;
    int* d_raw = (int*)pool.get("raw", (size_t)nb * n_cells * 4);
    int* d_ys = (int*)pool.get("ys", (size_t)nb * L.Npad * 4);
    double* d_mu_raw = (double*)pool.get("mu_raw", (size_t)nb * n_cells * 8);
    double* d_mu = (double*)pool.get("mu", (size_t)nb * L.Npad * 8);
    double* d_mom = (double*)pool.get("mom", (size_t)nb * NBMAX * NACC * 8);
    double* d_scores = (double*)pool.get("scores", (size_t)nb * 2 * L.T * 8);
    GeneStats* d_stats = (GeneStats*)pool.get("stats", (size_t)nb * sizeof(GeneStats));
    GeneOutDev* d_out = (GeneOutDev*)pool.get("out", (size_t)nb * sizeof(GeneOutDev));
    CK(cudaMemcpyAsync(d_raw, counts, (size_t)nb * n_cells * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_mu_raw, mu, (size_t)nb * n_cells * 8, cudaMemcpyHostToDevice, s));
    k_gather_stats<int><<<nb, kThreads, 0, s>>>(d_raw, (long long)n_cells, D.perm, L.N, L.Npad, d_ys, d_stats, d_out, nb);
    CK(cudaGetLastError()); g_launches.fetch_add(1);
    k_gather_f64<<<nb, 256, 0, s>>>(d_mu_raw, (long long)n_cells, D.perm, L.N, L.Npad, d_mu, nb);
    CK(cudaGetLastError()); g_launches.fetch_add(1);
    std::vector<GeneOutDev> h((size_t)nb);
    CK(cudaMemcpyAsync(h.data(), d_out, (size_t)nb * sizeof(GeneOutDev), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int g = 0; g < nb; ++g) { h[(size_t)g].st.fwd = m; h[(size_t)g].st.sigma = sigma[g]; h[(size_t)g].st.error = 0; }
    CK(cudaMemcpyAsync(d_out, h.data(), (size_t)nb * sizeof(GeneOutDev), cudaMemcpyHostToDevice, s));
    StageArgs A;
    A.lineage = D.lin; A.ys = d_ys; A.u = D.u; A.off = nullptr; A.mu = d_mu; A.stats = d_stats; A.out = d_out;
    A.scores = d_scores; A.mom = d_mom; A.n_genes = nb; A.M = 99; A.tols_score = 1e-5; A.use_offset = 0; A.order = nullptr; A.cg = nullptr; A.tail = nullptr; A.big = nullptr; A.exact = 0; A.pstart = nullptr;
    launch_sweep(A, true, s, nullptr, nullptr);
    std::vector<double> hs((size_t)nb * 2 * L.T);
    CK(cudaMemcpyAsync(hs.data(), d_scores, hs.size() * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    std::memcpy(scores_out, hs.data(), hs.size() * 8);
  } catch (const CudaFail& f) { set_err(errbuf, errlen, "%s", f.msg.c_str()); return SCLANE_ERR_CUDA; }
  return SCLANE_OK;
}

int32_t sclane_link_preds(const sclane_record* rec, const double* x, int64_t n, const double* size_factor, int32_t device,
                          double* marge_link_fit, double* marge_link_se, double* null_link_fit, double* null_link_se,
                          char* errbuf, size_t errlen) {
  if (!rec || !x || n < 1) { set_err(errbuf, errlen, "bad arguments"); return SCLANE_ERR_ARG; }
  if (usable_devices() <= 0) { set_err(errbuf, errlen, "no usable CUDA device (this library has no CPU fallback)"); return SCLANE_ERR_CUDA; }
  PredModel M;
  std::memset(&M, 0, sizeof(M));
  if (rec->status & SCLANE_MARGE_OK) {
    M.n_alt = rec->n_terms;
    for (int j = 0; j < rec->n_terms; ++j) { M.kind[j] = rec->term_kind[j]; M.knot[j] = rec->term_knot[j]; M.coef[j] = rec->coef[j]; }
    for (int j = 0; j < rec->n_terms * rec->n_terms; ++j) M.cov[j] = rec->cov[j];
  }
  if (rec->status & SCLANE_NULL_OK) { M.has_null = 1; M.null_coef = rec->null_coef; M.null_var = rec->null_se * rec->null_se; }
  try {
    CK(cudaSetDevice(device));
    DevBuf<double> dx, dsf, o[4];
    dx.alloc((size_t)n);
    CK(cudaMemcpy(dx.p, x, (size_t)n * 8, cudaMemcpyHostToDevice));
    if (size_factor) { dsf.alloc((size_t)n); CK(cudaMemcpy(dsf.p, size_factor, (size_t)n * 8, cudaMemcpyHostToDevice)); }
    double* host[4] = {marge_link_fit, marge_link_se, null_link_fit, null_link_se};
    for (int k = 0; k < 4; ++k) if (host[k]) o[k].alloc((size_t)n);
    k_link_preds<<<(unsigned)((n + 255) / 256), 256>>>(M, dx.p, dsf.p, (long long)n, o[0].p, o[1].p, o[2].p, o[3].p);
    CK(cudaGetLastError()); g_launches.fetch_add(1);
    for (int k = 0; k < 4; ++k) if (host[k]) CK(cudaMemcpy(host[k], o[k].p, (size_t)n * 8, cudaMemcpyDeviceToHost));
  } catch (const CudaFail& f) { set_err(errbuf, errlen, "%s", f.msg.c_str()); return SCLANE_ERR_CUDA; }
  return SCLANE_OK;
}

int32_t sclane_smoothed_counts(const sclane_record* recs, int64_t n_genes, int32_t n_lineages, int32_t lineage, const double* x, int64_t n,
                               const double* size_factor, int32_t is_gee, int32_t log1p_norm, int32_t device, double* out,
                               char* errbuf, size_t errlen) {
  if (!recs || !x || !out || n < 1 || n_genes < 1 || n_lineages < 1 || lineage < 0 || lineage >= n_lineages) { set_err(errbuf, errlen, "bad arguments"); return SCLANE_ERR_ARG; }
  if (n_genes > 65535) { set_err(errbuf, errlen, "at most 65535 genes per call"); return SCLANE_ERR_ARG; }
  if (usable_devices() <= 0) { set_err(errbuf, errlen, "no usable CUDA device (this library has no CPU fallback)"); return SCLANE_ERR_CUDA; }
  std::vector<SmoothModel> models((size_t)n_genes);
  for (int64_t g = 0; g < n_genes; ++g) {
    const sclane_record& r = recs[g * n_lineages + lineage];
    SmoothModel& M = models[(size_t)g];
    std::memset(&M, 0, sizeof(M));
    if (r.status & SCLANE_MARGE_OK) {
      M.n_alt = r.n_terms;
      for (int j = 0; j < r.n_terms; ++j) { M.kind[j] = r.term_kind[j]; M.knot[j] = r.term_knot[j]; M.coef[j] = r.coef[j]; }
    }
  }
  try {
    CK(cudaSetDevice(device));
    DevBuf<double> dx, dsf, dout;
    DevBuf<SmoothModel> dm;
    dx.alloc((size_t)n); dout.alloc((size_t)n * (size_t)n_genes); dm.alloc((size_t)n_genes);
    CK(cudaMemcpy(dx.p, x, (size_t)n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dm.p, models.data(), (size_t)n_genes * sizeof(SmoothModel), cudaMemcpyHostToDevice));
    if (size_factor) { dsf.alloc((size_t)n); CK(cudaMemcpy(dsf.p, size_factor, (size_t)n * 8, cudaMemcpyHostToDevice)); }
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)n_genes);
    k_smoothed_counts<<<grid, 256>>>(dm.p, dx.p, dsf.p, (long long)n, is_gee ? 0 : 1, log1p_norm, dout.p);
    CK(cudaGetLastError()); g_launches.fetch_add(1);
    CK(cudaMemcpy(out, dout.p, (size_t)n * (size_t)n_genes * 8, cudaMemcpyDeviceToHost));
  } catch (const CudaFail& f) { set_err(errbuf, errlen, "%s", f.msg.c_str()); return SCLANE_ERR_CUDA; }
  return SCLANE_OK;
}

int32_t sclane_matmul(const double* A, const double* B, double* C, int64_t m, int64_t k, int64_t n, int32_t device, char* errbuf, size_t errlen) {
  if (!A || !B || !C || m < 1 || k < 1 || n < 1) { set_err(errbuf, errlen, "bad arguments"); return SCLANE_ERR_ARG; }
  if (m > 0x7fffffff || k > 0x7fffffff || n > 0x7fffffff || (m + 15) / 16 > 0x7fffffff || (n + 15) / 16 > 65535) {
    set_err(errbuf, errlen, "matrix dimensions out of range (m, k < 2^31; n <= 1,048,560)"); return SCLANE_ERR_ARG;
  }
  if (usable_devices() <= 0) { set_err(errbuf, errlen, "no usable CUDA device (this library has no CPU fallback)"); return SCLANE_ERR_CUDA; }
  try {
    CK(cudaSetDevice(device));
    DevBuf<double> dA, dB, dC;
    dA.alloc((size_t)(m * k)); dB.alloc((size_t)(k * n)); dC.alloc((size_t)(m * n));
    CK(cudaMemcpy(dA.p, A, (size_t)(m * k) * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB.p, B, (size_t)(k * n) * 8, cudaMemcpyHostToDevice));
    dim3 grid((unsigned)((m + 15) / 16), (unsigned)((n + 15) / 16)), block(16, 16);
    k_matmul<<<grid, block>>>(dA.p, dB.p, dC.p, (int)m, (int)k, (int)n);
    CK(cudaGetLastError()); g_launches.fetch_add(1);
    CK(cudaMemcpy(C, dC.p, (size_t)(m * n) * 8, cudaMemcpyDeviceToHost));
  } catch (const CudaFail& f) { set_err(errbuf, errlen, "%s", f.msg.c_str()); return SCLANE_ERR_CUDA; }
  return SCLANE_OK;
}

int32_t sclane_llt_inverse(const double* A, double* Ainv, int64_t n, int32_t device, char* errbuf, size_t errlen) {
  if (!A || !Ainv || n < 1 || n > 64) { set_err(errbuf, errlen, "bad arguments (1 <= n <= 64)"); return SCLANE_ERR_ARG; }
  if (usable_devices() <= 0) { set_err(errbuf, errlen, "no usable CUDA device (this library has no CPU fallback)"); return SCLANE_ERR_CUDA; }
  try {
    CK(cudaSetDevice(device));
    DevBuf<double> dA, dI;
    dA.alloc((size_t)(n * n)); dI.alloc((size_t)(n * n));
    CK(cudaMemcpy(dA.p, A, (size_t)(n * n) * 8, cudaMemcpyHostToDevice));
    k_llt_inverse<<<1, 64>>>(dA.p, dI.p, (int)n);
    CK(cudaGetLastError()); g_launches.fetch_add(1);
    CK(cudaMemcpy(Ainv, dI.p, (size_t)(n * n) * 8, cudaMemcpyDeviceToHost));
  } catch (const CudaFail& f) { set_err(errbuf, errlen, "%s", f.msg.c_str()); return SCLANE_ERR_CUDA; }
  return SCLANE_OK;
}

int32_t sclane_pinv(const double* A, double* Apinv, int64_t m, int64_t n, double tolerance, int32_t device, char* errbuf, size_t errlen) {
  if (!A || !Apinv || m < 1 || n < 1 || m > 48 || n > 48) { set_err(errbuf, errlen, "bad arguments (1 <= m, n <= 48)"); return SCLANE_ERR_ARG; }
  if (usable_devices() <= 0) { set_err(errbuf, errlen, "no usable CUDA device (this library has no CPU fallback)"); return SCLANE_ERR_CUDA; }
  try {
    CK(cudaSetDevice(device));
    // the kernel wants a tall matrix; for a wide one use pinv(A) = pinv(A^T)^T
    const bool wide = m < n;
    const int64_t mm = wide ? n : m, nn = wide ? m : n;
    std::vector<double> At;
    const double* src = A;
    if (wide) {
      At.resize((size_t)(m * n));
      for (int64_t r = 0; r < m; ++r) for (int64_t c = 0; c < n; ++c) At[(size_t)(r * n + c)] = A[(size_t)(c * m + r)];   // A^T (n x m) column-major
      src = At.data();
    }
    DevBuf<double> dA, dP;
    dA.alloc((size_t)(mm * nn)); dP.alloc((size_t)(mm * nn));
    CK(cudaMemcpy(dA.p, src, (size_t)(mm * nn) * 8, cudaMemcpyHostToDevice));
    k_pinv_tall<<<1, 64>>>(dA.p, dP.p, (int)mm, (int)nn, tolerance);
    CK(cudaGetLastError()); g_launches.fetch_add(1);
    std::vector<double> P((size_t)(mm * nn));
    CK(cudaMemcpy(P.data(), dP.p, P.size() * 8, cudaMemcpyDeviceToHost));   // nn x mm column-major
    if (!wide) std::memcpy(Apinv, P.data(), P.size() * 8);                  // n x m
    else for (int64_t r = 0; r < n; ++r) for (int64_t c = 0; c < m; ++c) Apinv[(size_t)(c * n + r)] = P[(size_t)(r * nn + c)];   // transpose back: (m x n)^T
  } catch (const CudaFail& f) { set_err(errbuf, errlen, "%s", f.msg.c_str()); return SCLANE_ERR_CUDA; }
  return SCLANE_OK;
}

}  // namespace impl

// =====================================================================================================
// exported C ABI: every body runs under the library lock and inside a try block (see guarded())
// =====================================================================================================
extern "C" {

int32_t sclane_abi_version(void) { return SCLANE_ABI_VERSION; }
int32_t sclane_device_count(void) { return usable_devices(); }
size_t sclane_record_size(void) { return sizeof(sclane_record); }
int64_t sclane_launch_count(void) { return g_launches.load(); }
void sclane_reset_launch_count(void) { g_launches.store(0); }
void sclane_default_opts(sclane_opts* o) { impl::sclane_default_opts(o); }
void sclane_get_profile(sclane_profile* out) { impl::sclane_get_profile(out); }
void sclane_release(void) {
  std::lock_guard<std::mutex> lk(g_call_mu);
  try { impl::sclane_release(); } catch (...) {}
}

int32_t sclane_test_dynamic(const int32_t* counts, int64_t n_cells, int64_t n_genes, const double* pt, int32_t n_lineages,
                            const double* size_factor, const sclane_opts* opts, sclane_record* out,
                            sclane_lineage_info* lineages, char* errbuf, size_t errlen) {
  return guarded(errbuf, errlen, [&]() -> int32_t {
    return impl::sclane_test_dynamic(counts, n_cells, n_genes, pt, n_lineages, size_factor, opts, out, lineages, errbuf, errlen);
  });
}
int32_t sclane_marge_basis(const int32_t* counts, int64_t n_cells, int64_t n_genes, const double* pt, int32_t n_lineages,
                           const sclane_opts* opts, sclane_record* out, sclane_lineage_info* lineages, char* errbuf, size_t errlen) {
  return guarded(errbuf, errlen, [&]() -> int32_t {
    return impl::sclane_marge_basis(counts, n_cells, n_genes, pt, n_lineages, opts, out, lineages, errbuf, errlen);
  });
}
int32_t sclane_test_dynamic_gee(const int32_t* counts, int64_t n_cells, int64_t n_genes, const double* pt, int32_t n_lineages,
                                const int32_t* subject_id, const double* size_factor, const sclane_opts* opts, sclane_record* out,
                                sclane_lineage_info* lineages, char* errbuf, size_t errlen) {
  return guarded(errbuf, errlen, [&]() -> int32_t {
    return impl::sclane_test_dynamic_gee(counts, n_cells, n_genes, pt, n_lineages, subject_id, size_factor, opts, out, lineages, errbuf, errlen);
  });
}
int32_t sclane_test_dynamic_csc(const int32_t* indptr, const int32_t* indices, const double* data, int64_t n_cells, int64_t n_genes,
                                const double* pt, int32_t n_lineages, const double* size_factor, const sclane_opts* opts,
                                sclane_record* out, sclane_lineage_info* lineages, char* errbuf, size_t errlen) {
  return guarded(errbuf, errlen, [&]() -> int32_t {
    return impl::sclane_test_dynamic_csc(indptr, indices, data, n_cells, n_genes, pt, n_lineages, size_factor, opts, out, lineages, errbuf, errlen);
  });
}
int32_t sclane_test_dynamic_device(const int32_t* d_counts, int32_t device, int64_t n_cells, int64_t n_genes, const double* pt,
                                   int32_t n_lineages, const double* size_factor, const sclane_opts* opts, sclane_record* out,
                                   sclane_lineage_info* lineages, char* errbuf, size_t errlen) {
  return guarded(errbuf, errlen, [&]() -> int32_t {
    return impl::sclane_test_dynamic_device(d_counts, device, n_cells, n_genes, pt, n_lineages, size_factor, opts, out, lineages, errbuf, errlen);
  });
}
int32_t sclane_candidate_knots(const double* x, int64_t n, const sclane_opts* opts, double* knots_out, int32_t* minspan_out) {
  try { return impl::sclane_candidate_knots(x, n, opts, knots_out, minspan_out); }
  catch (const std::bad_alloc&) { return -SCLANE_ERR_NOMEM; }
  catch (...) { return -SCLANE_ERR_INTERNAL; }
}
int32_t sclane_lineage_plan(const double* pt, int64_t n_cells, const sclane_opts* opts, int32_t on_device, int32_t device,
                            int32_t* perm, double* u, double* knots, int32_t* bstart, double* U, int32_t* pstart, double* pn,
                            double* pu, uint8_t* pb, int32_t* pbstart, int32_t* dims, char* errbuf, size_t errlen) {
  return guarded(errbuf, errlen, [&]() -> int32_t {
    return impl::sclane_lineage_plan(pt, n_cells, opts, on_device, device, perm, u, knots, bstart, U, pstart, pn, pu, pb, pbstart, dims, errbuf, errlen);
  });
}
int32_t sclane_null_stats_glm(const int32_t* counts, int64_t n_cells, int64_t n_genes, const double* x, const double* knots,
                              int32_t n_knots, const int32_t* term_kind, const int32_t* term_knot_idx, int32_t n_terms,
                              int32_t device, double* beta_out, double* sigma_out, double* mu_out, char* errbuf, size_t errlen) {
  return guarded(errbuf, errlen, [&]() -> int32_t {
    return impl::sclane_null_stats_glm(counts, n_cells, n_genes, x, knots, n_knots, term_kind, term_knot_idx, n_terms, device, beta_out, sigma_out, mu_out, errbuf, errlen);
  });
}
int32_t sclane_score_glm(const int32_t* counts, const double* mu, int64_t n_cells, int64_t n_genes, const double* x,
                         const double* knots, int32_t n_knots, const int32_t* term_kind, const int32_t* term_knot_idx,
                         int32_t n_terms, const double* sigma, int32_t device, double* scores_out, char* errbuf, size_t errlen) {
  return guarded(errbuf, errlen, [&]() -> int32_t {
    return impl::sclane_score_glm(counts, mu, n_cells, n_genes, x, knots, n_knots, term_kind, term_knot_idx, n_terms, sigma, device, scores_out, errbuf, errlen);
  });
}
int32_t sclane_link_preds(const sclane_record* rec, const double* x, int64_t n, const double* size_factor, int32_t device,
                          double* marge_link_fit, double* marge_link_se, double* null_link_fit, double* null_link_se,
                          char* errbuf, size_t errlen) {
  return guarded(errbuf, errlen, [&]() -> int32_t {
    return impl::sclane_link_preds(rec, x, n, size_factor, device, marge_link_fit, marge_link_se, null_link_fit, null_link_se, errbuf, errlen);
  });
}
int32_t sclane_smoothed_counts(const sclane_record* recs, int64_t n_genes, int32_t n_lineages, int32_t lineage, const double* x, int64_t n,
                               const double* size_factor, int32_t is_gee, int32_t log1p_norm, int32_t device, double* out,
                               char* errbuf, size_t errlen) {
  return guarded(errbuf, errlen, [&]() -> int32_t {
    return impl::sclane_smoothed_counts(recs, n_genes, n_lineages, lineage, x, n, size_factor, is_gee, log1p_norm, device, out, errbuf, errlen);
  });
}
int32_t sclane_matmul(const double* A, const double* B, double* C, int64_t m, int64_t k, int64_t n, int32_t device, char* errbuf, size_t errlen) {
  return guarded(errbuf, errlen, [&]() -> int32_t { return impl::sclane_matmul(A, B, C, m, k, n, device, errbuf, errlen); });
}
int32_t sclane_llt_inverse(const double* A, double* Ainv, int64_t n, int32_t device, char* errbuf, size_t errlen) {
  return guarded(errbuf, errlen, [&]() -> int32_t { return impl::sclane_llt_inverse(A, Ainv, n, device, errbuf, errlen); });
}
int32_t sclane_pinv(const double* A, double* Apinv, int64_t m, int64_t n, double tolerance, int32_t device, char* errbuf, size_t errlen) {
  return guarded(errbuf, errlen, [&]() -> int32_t { return impl::sclane_pinv(A, Apinv, m, n, tolerance, device, errbuf, errlen); });
}

}  // extern "C"
