// sclane_comp_device.cuh -- CUDA side of the X-compressed GLM path (see sclane_comp.cuh for the algebra).
//
//   k_aggregate      one CTA per gene: reads the gene's pseudotime-sorted counts once (4 B / cell) and reduces them to
//                    the per-point sums s_d = sum y, q_d = sum y^2, the gene's tail counts C_k = #{y > k}, the per-bucket
//                    totals and the glm.fit mustart sums.  All sums run in a fixed order (integer atomics for the
//                    histogram), so the result is bit-reproducible.
//   k_comp_stage<S>  one CTA per gene: the same per-gene algorithm as k_stage (sclane_core.cuh), driven by CompDevExec,
//                    whose pass<MODE>() walks the D distinct abscissae (not the N cells) plus ymax histogram terms, with a
//                    closed form for buckets where the linear predictor is flat.  S = forward pass (all iterations: null
//                    fit, score sweep, selection), backward pass, final + null glm.nb.
// Measured and rejected in k_aggregate (11.4 ms at the default workload, 0.55 instructions issued per cycle and scheduler, ~60
// instructions per cell): staging the cells of 256 points through shared memory with coalesced loads (same 11.4 ms: the strided
// per-thread reads were not the limit; 62 KB of shared memory); counting a point's zeros locally and updating hist[0] once per
// point (13.8 vs 13.6 ms with the node weights: the contended bin is not the limit either).
// Measured and rejected: aggregating straight from the caller's unsorted rows (no sorted copy) with integer / fixed-point
// shared-memory atomics per non-zero cell -- exact and order independent, but 107 ms at the default workload against 27.7 ms
// for k_gather_copy + k_aggregate (64-bit shared atomics and the contended per-bucket slots run at ~0.5 per clock per SM).
// Counts >= kHistCap are kept as a short sorted list per gene; a gene with more than kBigCap such cells keeps route = 0 and
// runs through the per-cell kernels instead.
#pragma once
#include "sclane_comp.cuh"
#include "sclane_device.cuh"

namespace sclane {

struct CompArgs {
  const Lineage* lineage;
  const int* pstart;            // [D + 1]
  const double* pn;             // [D]
  const double* pu;             // [D]
  const unsigned char* pb;      // [D] bucket of the point
  const int* pbstart;           // [NBMAX + 1] point range per bucket
  int D, Dpad;
  const int* ys;                // [B][Npad] sorted counts (k_aggregate input)
  double* s1;                   // [B][Dpad]
  double* q;                    // [B][Dpad]
  int* tail;                    // [B][kHistCap]
  int* big;                     // [B][kBigCap] counts >= kHistCap of the gene, ascending
  int* big_tmp;                 // [B][kBigCap] unsorted staging
  CompGeneDev* cg;              // [B]
  GeneStats* stats;             // written by k_aggregate when compute_stats != 0 (k_gather_copy route)
  GeneOutDev* out;
  int n_genes, M, n_iter;
  int compute_stats;
  double tols_score;
  const int* order;             // optional CTA -> gene map of the final stage (k_order_final), or nullptr
  int theta_nodes;              // k_comp_stage<FINAL>: theta.ml on Chebyshev nodes (needs sizeof(ThetaNodes) bytes of dynamic shared memory)
  double* a0p;                  // optional [B][Dpad]: glm.fit start sums per point (sum w0, sum w0 z0) for the exact-knot kernels
  double* b0p;
  const BucketNodes* bn = nullptr;   // bucket nodes (sclane_comp.cuh) or nullptr: every sloped bucket walks its points
  double* bn_s = nullptr;            // [B][NBMAX][16] node weights of sum y
  double* bn_q = nullptr;            // [B][NBMAX][16] node weights of sum y^2
};

// Chebyshev moments of val[] over the points [p0, p1) of one bucket -> the 16 node weights; lanes j and j + 16 both hold node
// j's weight of `val_lo` (lanes 0-15) / `val_hi` (lanes 16-31).  All 32 lanes must call.
__device__ __forceinline__ double bucket_node_weight(const double* __restrict__ val_lo, const double* __restrict__ val_hi, const double* __restrict__ pu,
                                                     int p0, int p1, double ulo, double uw) {
  const int lane = threadIdx.x & 31, half = lane >> 4, j = lane & 15;
  const double* __restrict__ val = half ? val_hi : val_lo;
  const double sc2 = uw > 0.0 ? 2.0 / uw : 0.0;
  double cm[kOffJ];
#pragma unroll
  for (int m = 0; m < kOffJ; ++m) cm[m] = 0.0;
  if (val != nullptr) {
    for (int d = p0 + j; d < p1; d += 16) {
      const double v = val[d];
      double xi = sc2 * (pu[d] - ulo) - 1.0;
      xi = fmin(1.0, fmax(-1.0, xi));
      double t0 = 1.0, t1 = xi;
      cm[0] += v; cm[1] += v * xi;
#pragma unroll
      for (int m = 2; m < kOffJ; ++m) { const double t2 = 2.0 * xi * t1 - t0; cm[m] += v * t2; t0 = t1; t1 = t2; }
    }
  }
#pragma unroll
  for (int m = 0; m < kOffJ; ++m) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) cm[m] += __shfl_xor_sync(0xffffffffu, cm[m], o);
  }
  double sacc = 0.0;
#pragma unroll
  for (int m = kOffJ - 1; m >= 1; --m) sacc += cospi((double)(m * (2 * j + 1)) / (double)(2 * kOffJ)) * cm[m];
  return (cm[0] + 2.0 * sacc) * (1.0 / (double)kOffJ);
}

// gene independent part of the bucket nodes, once per lineage: one warp per bucket
__global__ void __launch_bounds__(kThreads) k_bucket_nodes(const Lineage* __restrict__ lineage, const double* __restrict__ pn, const double* __restrict__ pu,
                                                           const int* __restrict__ pbstart, BucketNodes* __restrict__ bn) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int T = lineage->T;
  for (int b = warp; b < NBMAX; b += kWarps) {
    const int p0 = b <= T ? pbstart[b] : 0, p1 = b <= T ? pbstart[b + 1] : 0;
    const int P = p1 - p0;
    const double ulo = P > 0 ? pu[p0] : 0.0, uw = P > 0 ? pu[p1 - 1] - ulo : 0.0;     // points are sorted by u inside a bucket
    const bool ok = P >= kBnMinPoints && uw > 0.0;
    const double w = ok ? bucket_node_weight(pn, nullptr, pu, p0, p1, ulo, uw) : 0.0;
    if (lane == 0) { bn->elig[b] = ok ? 1 : 0; bn->ulo[b] = ulo; bn->uw[b] = uw; }
    if (lane < kOffJ) {
      bn->nn[b][lane] = w;
      bn->un[b][lane] = ulo + 0.5 * (cospi((double)(2 * lane + 1) / (double)(2 * kOffJ)) + 1.0) * uw;
    }
  }
}

#ifndef SCLANE_COMP_WARPS
#define SCLANE_COMP_WARPS 4
#endif
#ifndef SCLANE_COMP_MINB
#define SCLANE_COMP_MINB 8              // forward / backward stage: resident CTAs per SM the launch bounds ask for
#endif
#ifndef SCLANE_COMP_MINB_FINAL
#define SCLANE_COMP_MINB_FINAL 4        // final stage (glm.nb: more live state; 128 registers per thread instead of 64)
#endif
// CTA shape of k_comp_stage.  With the bucket nodes (sclane_comp.cuh) a pass is 16 evaluations per sloped bucket and the kernels are
// bound by the serial algebra between the passes and its barriers (ncu: 10-15 barrier stalls per issued instruction at 8 warps), so
// many small CTAs per SM win.  Measured on B200, C5 ms of forward / backward / final stage (C2 step ms): 8 warps x 3 CTAs/SM
// 13.0 / 12.7 / 19.9 (8.3); 4 x 6 11.5 / 8.7 / 19.8 (7.5); 4 x 4 11.7 / 10.4 / 16.4 (7.5); 2 x 12 11.2 / 9.0 / 21.9 (9.4);
// 4 x 8 for forward + backward with 4 x 4 for the final stage 11.1 / 8.1 / 16.4 (6.8) <- default; 4 x 6 + 4 x 3 11.6 / 8.8 / 19.1;
// 2 x 12 + 2 x 8 11.1 / 9.0 / 20.2 (14.1).
// BEFORE the bucket nodes the passes were throughput bound and the opposite held (C5 / C2 ms per step): 8 warps x 3 CTAs/SM 144.4 / 12.7, 4 x 6 150.0 / 13.6,
// 8 x 4 (64 registers) 143.3 / 13.2, 4 x 4 154.8 / 14.1, 8 x 2 167.8 / 13.7, 2 x 12 178.6 / 17.6 -- the passes are throughput bound, not leader-latency bound.
// Also measured and rejected (same units; baseline 144.7 / 12.7): an exactly balanced split of the sloped points over the
// warps with per-warp partial sums 153.9 / 13.9; two points per lane per iteration 161.4 / 14.9 (80 registers) and
// 170.0 / 14.0 (128 registers, 2 CTAs/SM); backward + final fused into one launch 162.0 / 15.2;
// flat buckets evaluated in parallel by the lanes of warp 0 + least-loaded assignment of the sloped buckets 149.7 / 12.7;
// means cached in global memory across the Newton iterations of theta.ml (41 of 56 passes per gene) 146.9 / 12.9.
#ifndef SCLANE_COMP_UNROLL
#define SCLANE_COMP_UNROLL 2            // independent points per lane in flight in the sloped-bucket loops
#endif
constexpr int kCompUnroll = SCLANE_COMP_UNROLL;
constexpr int kCompWarps = SCLANE_COMP_WARPS;
constexpr int kCompThreads = kCompWarps * 32;
constexpr int kAggV = 9;        // per-bucket totals gathered by k_aggregate

// add NV per-lane values into slot[b] (warp-private), lanes of a 32-point group usually share the bucket
template <int NV>
__device__ __forceinline__ void warp_bucket_add(const double (&v)[NV], int b, bool valid, double* myslot, int stride) {
  const int lane = threadIdx.x & 31;
  const int b0 = __shfl_sync(0xffffffffu, b, 0);
  if (__all_sync(0xffffffffu, !valid || b == b0)) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const double t = warp_sum(valid ? v[k] : 0.0);
      if (lane == 0) myslot[b0 * stride + k] += t;
    }
  } else {
    unsigned rem = __ballot_sync(0xffffffffu, valid);
    while (rem) {
      const int ld = __ffs(rem) - 1;
      const int bb = __shfl_sync(0xffffffffu, b, ld);
      const bool mine = valid && b == bb;
      const unsigned same = __ballot_sync(0xffffffffu, mine);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const double t = warp_sum(mine ? v[k] : 0.0);
        if (lane == 0) myslot[bb * stride + k] += t;
      }
      rem &= ~same;
    }
  }
}

__global__ void __launch_bounds__(kThreads) k_aggregate(CompArgs A) {
  __shared__ int hist[kHistCap];
  __shared__ double slot[kWarps][NBMAX][kAggV];
  __shared__ double wtab[64], wztab[64];
  __shared__ int chunk_sum[kThreads];
  __shared__ int n_big_s, ymax_s, neg_s;
  __shared__ double red_s[kWarps][2];
  const int g = blockIdx.x;
  if (g >= A.n_genes) return;
  if (threadIdx.x == 0) { n_big_s = 0; ymax_s = 0; neg_s = 0; }
  int* __restrict__ big_tmp = A.big_tmp + (size_t)g * (size_t)kBigCap;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int T = A.lineage->T, Npad = A.lineage->Npad;
  int my_max = 0;
  const int* __restrict__ y = A.ys + (size_t)g * (size_t)Npad;
  double* __restrict__ s1 = A.s1 + (size_t)g * (size_t)A.Dpad;
  double* __restrict__ q = A.q + (size_t)g * (size_t)A.Dpad;
  for (int i = threadIdx.x; i < kHistCap; i += kThreads) hist[i] = 0;
  for (int i = threadIdx.x; i < kWarps * NBMAX * kAggV; i += kThreads) (&slot[0][0][0])[i] = 0.0;
  if (threadIdx.x < 64) start_cell(threadIdx.x, wtab[threadIdx.x], wztab[threadIdx.x]);
  __syncthreads();
  double* myslot = &slot[warp][0][0];
  for (int base = warp * 32; base < A.D; base += kThreads) {
    const int d = base + lane;
    const bool valid = d < A.D;
    double v[kAggV];
#pragma unroll
    for (int k = 0; k < kAggV; ++k) v[k] = 0.0;
    int b = 0;
    if (valid) {
      const int i0 = A.pstart[d], i1 = A.pstart[d + 1];
      long long s = 0, qq = 0;
      double a0 = 0.0, b0 = 0.0;
      for (int i = i0; i < i1; ++i) {
        int yi = y[i];
        if (yi < 0) { yi = 0; neg_s = 1; }               // negative count: the gene fails (SCLANE_NOTE_BAD_COUNTS), summed as 0
        my_max = max(my_max, yi);
        s += yi;
        qq += (long long)yi * yi;
        if (yi < kHistCap) atomicAdd(&hist[yi], 1);
        else { const int slot_i = atomicAdd(&n_big_s, 1); if (slot_i < kBigCap) big_tmp[slot_i] = yi; }
        double w, wz;
        if (yi < 64) { w = wtab[yi]; wz = wztab[yi]; } else start_cell(yi, w, wz);
        a0 += w; b0 += wz;
      }
      const double sd = (double)s, qd = (double)qq, u = A.pu[d];
      s1[d] = sd; q[d] = qd;
      if (A.a0p) { A.a0p[(size_t)g * (size_t)A.Dpad + d] = a0; A.b0p[(size_t)g * (size_t)A.Dpad + d] = b0; }
      b = A.pb[d];
      v[0] = sd; v[1] = sd * u; v[2] = sd * u * u; v[3] = qd;
      v[4] = a0; v[5] = a0 * u; v[6] = a0 * u * u; v[7] = b0; v[8] = b0 * u;
    }
    warp_bucket_add<kAggV>(v, b, valid, myslot, kAggV);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my_max = max(my_max, __shfl_xor_sync(0xffffffffu, my_max, o));
  if (lane == 0) atomicMax(&ymax_s, my_max);
  __syncthreads();
  const int ymax = ymax_s;
  CompGeneDev& cg = A.cg[g];
  for (int idx = threadIdx.x; idx < (T + 1) * kAggV; idx += kThreads) {
    const int b = idx / kAggV, k = idx - b * kAggV;
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += slot[w][b][k];
    if (k == 0) cg.hdr.SB[b] = s;
    else if (k == 1) cg.hdr.SBu[b] = s;
    else if (k == 2) cg.hdr.SBuu[b] = s;
    else if (k == 3) cg.hdr.QB[b] = s;
    else cg.ST[b][k - 4] = s;
  }
  // tail counts: tail[k] = #{y > k} = sum_{j > k} hist[j]; suffix sums in chunks of kHistCap / kThreads
  constexpr int CH = kHistCap / kThreads;
  int local = 0;
  for (int j = 0; j < CH; ++j) local += hist[threadIdx.x * CH + j];
  chunk_sum[threadIdx.x] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int t = kThreads - 1; t >= 0; --t) { const int c = chunk_sum[t]; chunk_sum[t] = run; run += c; }   // sum of later chunks
    cg.hdr.ymax = ymax;
    cg.hdr.n_zero = hist[0];
    cg.hdr.route = n_big_s <= kBigCap ? 1 : 0;
    cg.hdr.n_big = n_big_s;
  }
  __syncthreads();
  const int n_big = n_big_s;
  const bool routed = n_big <= kBigCap;                   // block-uniform; otherwise the gene takes the per-cell kernels
  if (routed) {
    // tail[k] = #{k < y < kHistCap} for k <= hmax, the largest count that goes through the histogram
    const int hmax = ymax < kHistCap ? ymax : kHistCap - 1;
    int* __restrict__ tail = A.tail + (size_t)g * (size_t)kHistCap;
    int run = chunk_sum[threadIdx.x];
    for (int j = CH - 1; j >= 0; --j) {
      const int k = threadIdx.x * CH + j;
      if (k <= hmax) tail[k] = run;
      run += hist[k];
    }
  }
  if (routed && n_big > 0) {
    // the append order above is arbitrary: rank sort (ties are equal values) makes the list -- and every sum taken
    // over it -- reproducible
    __threadfence_block();
    int* __restrict__ big = A.big + (size_t)g * (size_t)kBigCap;
    for (int i = threadIdx.x; i < n_big; i += kThreads) {
      const int v = big_tmp[i];
      int rank = 0;
      for (int j = 0; j < n_big; ++j) { const int w = big_tmp[j]; rank += (w < v || (w == v && j < i)) ? 1 : 0; }
      big[rank] = v;
    }
  }
  if (!A.compute_stats) return;
  // per-gene constants (what k_gather_stats provides on the per-cell route) from the histogram and the sorted big list,
  // plus the initial per-gene state
  __syncthreads();
  {
    const int hmax = ymax < kHistCap ? ymax : kHistCap - 1;
    double t0 = 0.0, t1 = 0.0;
    if (routed) {
      for (int yv = 2 + (int)threadIdx.x; yv <= hmax; yv += kThreads) {
        const double c = (double)hist[yv], d = (double)yv;
        if (c != 0.0) { t0 += c * d * log(d); t1 += c * lgamma(d + 1.0); }
      }
      if (n_big > 0) {
        const int* __restrict__ big = A.big + (size_t)g * (size_t)kBigCap;
        for (int e = threadIdx.x; e < n_big; e += kThreads) { const double d = (double)big[e]; t0 += d * log(d); t1 += lgamma(d + 1.0); }
      }
    } else {
      const int N = A.lineage->N;
      for (int i = threadIdx.x; i < N; i += kThreads) {
        const int v = y[i];
        if (v > 1) { const double d = (double)v; t0 += d * log(d); t1 += lgamma(d + 1.0); }
      }
    }
    t0 = warp_sum(t0); t1 = warp_sum(t1);
    if (lane == 0) { red_s[warp][0] = t0; red_s[warp][1] = t1; }
    __syncthreads();
    if (threadIdx.x == 0) {
      GeneStats gs;
      gs.sum_y = 0.0; gs.sum_y2 = 0.0; gs.sum_ylogy = 0.0; gs.sum_lgamma_y1 = 0.0;
      for (int b = 0; b <= T; ++b) { gs.sum_y += cg.hdr.SB[b]; gs.sum_y2 += cg.hdr.QB[b]; }
      for (int w = 0; w < kWarps; ++w) { gs.sum_ylogy += red_s[w][0]; gs.sum_lgamma_y1 += red_s[w][1]; }
      gs.ymax = ymax; gs.pad = neg_s;
      A.stats[g] = gs;
      GeneState st;
      int* z = reinterpret_cast<int*>(&st);
      for (int i = 0; i < (int)(sizeof(GeneState) / sizeof(int)); ++i) z[i] = 0;
      st.fwd.n = 1;
      st.fwd.kind[0] = SCLANE_TERM_INTERCEPT;
      st.m_count = 1;
      if (gs.sum_y == 0.0) st.error |= SCLANE_NOTE_ALL_ZERO;
      if (neg_s) st.error |= SCLANE_NOTE_BAD_COUNTS;
      A.out[g].st = st;
      A.out[g].alt.ok = 0;
      A.out[g].null.ok = 0;
    }
  }
}

// Bucket nodes of every gene (its own kernel: inside k_aggregate the 16 moment registers cost that kernel two resident CTAs per
// SM): one warp per (gene, bucket) builds the Chebyshev moments of s_d and q_d over the bucket's points -> 2 x 16 node weights.
__global__ void __launch_bounds__(kThreads) k_gene_bucket_nodes(CompArgs A) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int T = A.lineage->T;
  const long long unit = (long long)blockIdx.x * kWarps + warp;
  const int g = (int)(unit / (T + 1)), b = (int)(unit - (long long)g * (T + 1));
  if (g >= A.n_genes) return;
  if (!A.bn->elig[b] || A.cg[g].hdr.route == 0) return;              // warp-uniform
  const double* __restrict__ s1 = A.s1 + (size_t)g * (size_t)A.Dpad;
  const double* __restrict__ q = A.q + (size_t)g * (size_t)A.Dpad;
  const double w = bucket_node_weight(s1, q, A.pu, A.pbstart[b], A.pbstart[b + 1], A.bn->ulo[b], A.bn->uw[b]);
  double* dst = (lane < kOffJ ? A.bn_s : A.bn_q) + (size_t)g * (size_t)(NBMAX * kOffJ) + b * kOffJ + (lane & 15);
  *dst = w;
}

// Pure permutation copy (the compressed route's replacement of k_gather_stats): K CTAs per gene, so that the rows being
// gathered at any moment fit in L2 and the 4-byte random reads of a row are served from there instead of costing a
// 32-byte DRAM sector each.
// ncu (profiles/r02_ncu_final_c5.txt): DRAM traffic == the algorithmic 8 B/cell, 1.9 TB/s; the warps wait on the load / store unit
// (lg_throttle 77, long scoreboard 41 per issued instruction): a warp's gather touches 32 different sectors, and 4e9 such
// accesses per 20,000 x 200,000 step at ~1 sector per clock and SM are ~14 ms.  Measured and rejected: the same permutation as
// a scatter (coalesced reads of the caller's row, 4-byte stores to the sorted positions): 29.6 ms against 16.6.
// RAW = int (the caller's counts) or unsigned short (counts narrowed on the host for the PCIe transfer).
template <class RAW>
__global__ void __launch_bounds__(kThreads) k_gather_copy(const RAW* __restrict__ raw, long long ld, const int* __restrict__ perm, int N,
                                                          int Npad, int* __restrict__ ys, int n_genes, int K) {
  const int g = blockIdx.x / K, part = blockIdx.x - g * K;
  if (g >= n_genes) return;
  const int chunk = ((Npad / K + 127) / 128) * 128;
  const int i0 = part * chunk, i1 = min(Npad, i0 + chunk);
  const RAW* src = raw + (size_t)g * (size_t)ld;
  int* dst = ys + (size_t)g * (size_t)Npad;
  for (int i = i0 + threadIdx.x; i < i1; i += kThreads) dst[i] = i < N ? (int)__ldg(src + __ldg(perm + i)) : 0;
}

template <int MODE> struct CompTraits;
template <> struct CompTraits<PASS_FIT>        { static constexpr int NA = 7, NS = 3; };
template <> struct CompTraits<PASS_IRLS>       { static constexpr int NA = 5, NS = 3; };
template <> struct CompTraits<PASS_IRLS_START> { static constexpr int NA = 5, NS = 2; };
template <> struct CompTraits<PASS_THETA>      { static constexpr int NA = 0, NS = 2; };
template <> struct CompTraits<PASS_SWEEP>      { static constexpr int NA = 5, NS = 0; };
template <> struct CompTraits<PASS_EMIT>       { static constexpr int NA = 0, NS = 0; };

struct CompDevExec {
  static constexpr bool kTabs = true;      // term tables attached to Work (sclane_core.cuh term_coef)
  Work* W;
  const Lineage* L;
  const CompGeneDev* cg;            // shared-memory copy
  const GeneStats* gs;
  int D;
  const double* __restrict__ pn;
  const double* __restrict__ pu;
  const int* pbstart;               // [NBMAX + 1] point range per bucket (shared memory)
  const double* __restrict__ s1;
  const double* __restrict__ q;
  const int* __restrict__ tail;
  const int* __restrict__ big;
  double (*scw)[NSC];
  const FastTabs* ft;               // exp / log1p tables in shared memory (sclane_fastmath.cuh)
  const double* __restrict__ px;    // exact-knot mode: abscissa per point, u = px[d] - ref[bucket] (the buckets are per gene); else nullptr
  ThetaNodes* tn;                   // theta.ml node list (shared memory; sclane_comp.cuh) or nullptr: theta passes walk the points
  const double (*ctab)[kOffJ];      // T_m(xi_j) table (shared memory) when tn is set
  const BucketNodes* bn = nullptr;  // bucket nodes (global memory) or nullptr
  const double* __restrict__ bs = nullptr;      // this gene's node weights of sum y   [NBMAX][16]
  const double* __restrict__ bq = nullptr;      // ... of sum y^2
  const double* bnuw = nullptr;     // [NBMAX] shared-memory copy of bn->uw (-1 = bucket without nodes): the test below runs per bucket,
                                    // per pass and per warp, and as two global loads it was 32 % of the final stage's stall samples
  __device__ __forceinline__ bool on_bucket_nodes(int b, double c) const {
    if (bnuw == nullptr || px != nullptr) return false;
    const double w = bnuw[b];
    return w > 0.0 && fabs(c) * w <= kThBinWidth;
  }

  __device__ __forceinline__ double u_of(int b, int d) const { return px ? px[d] - L->ref[b] : pu[d]; }

  // theta.ml is about to iterate at the mean defined by W.a / W.c: condense the points into the node list (sclane_comp.cuh).
  // Units (flat bucket / raw bucket / one bin) are dealt round-robin to the warps; a bin is built by one warp: lanes 0-15
  // accumulate the Chebyshev moments of n, lanes 16-31 those of s (both halves walk all points of the bin, stride 16).
  __device__ __noinline__ void theta_prepare() {
    if (!tn) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int T = L->T;
    __syncthreads();
    if (bn != nullptr && px == nullptr) {
      // every sloped bucket on its bucket nodes: the Newton steps need no list of their own
      if (threadIdx.x == 0) {
        int ok = 1;
        for (int b = 0; b <= T; ++b) { const double c = W->c[b]; if (c != 0.0 && !on_bucket_nodes(b, c)) ok = 0; }
        tn->skip = ok;
        if (ok) { tn->on = 0; tn->n = 0; }
      }
      __syncthreads();
      if (tn->skip) return;
    }
    if (threadIdx.x == 0) theta_nodes_layout(*tn, T, W->c, pbstart, [&](int b, int d) { return u_of(b, d); });
    __syncthreads();
    if (!tn->on) return;
    int unit = 0;
#pragma unroll 1
    for (int b = 0; b <= T; ++b) {
      const int kb = tn->bins[b], p0 = pbstart[b], p1 = pbstart[b + 1], e0 = tn->first[b];
      const double a = W->a[b], c = W->c[b];
      if (kb == 0) {
        if ((unit++ & (kCompWarps - 1)) == warp && lane == 0) { tn->en[e0] = L->U0[b]; tn->es[e0] = cg->hdr.SB[b]; tn->eeta[e0] = a; }
        continue;
      }
      if (kb < 0) {
        if ((unit++ & (kCompWarps - 1)) == warp)
          for (int d = p0 + lane; d < p1; d += 32) { const int e = e0 + (d - p0); tn->en[e] = pn[d]; tn->es[e] = s1[d]; tn->eeta[e] = a + c * u_of(b, d); }
        continue;
      }
      const double h = tn->h[b], ulo = tn->ulo[b];
#pragma unroll 1
      for (int k = 0; k < kb; ++k) {
        if ((unit++ & (kCompWarps - 1)) != warp) continue;
        const double u0 = ulo + (double)k * h, u1 = ulo + (double)(k + 1) * h;
        // point range of the bin: [first point with u >= u0, first point with u >= u1); the outer bins take the bucket's ends
        int lo = p0, hi = p1;
        if (k > 0) { int x = p0, y = p1; while (x < y) { const int m = (x + y) >> 1; if (u_of(b, m) < u0) x = m + 1; else y = m; } lo = x; }
        if (k + 1 < kb) { int x = p0, y = p1; while (x < y) { const int m = (x + y) >> 1; if (u_of(b, m) < u1) x = m + 1; else y = m; } hi = x; }
        const int half = lane >> 4, j = lane & 15;
        const double* __restrict__ val = half ? s1 : pn;
        const double sc2 = h > 0.0 ? 2.0 / h : 0.0;
        double cm[kOffJ];
#pragma unroll
        for (int m = 0; m < kOffJ; ++m) cm[m] = 0.0;
        for (int d = lo + j; d < hi; d += 16) {
          const double v = val[d];
          double xi = sc2 * (u_of(b, d) - u0) - 1.0;
          xi = fmin(1.0, fmax(-1.0, xi));
          double t0 = 1.0, t1 = xi;
          cm[0] += v; cm[1] += v * xi;
#pragma unroll
          for (int m = 2; m < kOffJ; ++m) { const double t2 = 2.0 * xi * t1 - t0; cm[m] += v * t2; t0 = t1; t1 = t2; }
        }
#pragma unroll
        for (int m = 0; m < kOffJ; ++m) {
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) cm[m] += __shfl_xor_sync(0xffffffffu, cm[m], o);
        }
        double sacc = 0.0;
#pragma unroll
        for (int m = kOffJ - 1; m >= 1; --m) sacc += ctab[m][j] * cm[m];
        const double wgt = (cm[0] + 2.0 * sacc) * (1.0 / (double)kOffJ);
        const int e = e0 + k * kOffJ + j;
        if (half) tn->es[e] = wgt;
        else { tn->en[e] = wgt; tn->eeta[e] = a + c * (u0 + 0.5 * (ctab[1][j] + 1.0) * h); }
      }
    }
    __syncthreads();
  }

  __device__ __forceinline__ bool leader() const { return threadIdx.x == 0; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  __device__ __forceinline__ PassTabs* tabs() const { return nullptr; }
  template <class F>
  __device__ __forceinline__ void par_for(int n, F f) const {
    for (int i = threadIdx.x; i < n; i += kCompThreads) f(i);
  }

  // One pass.  Work unit = one knot bucket, owned by one warp: a bucket with a flat linear predictor is a single closed-form
  // evaluation (lane 0); otherwise the bucket's points are walked lane-strided with independent iterations (the loads of
  // the next points are in flight while the current ones are evaluated), reduced with one shuffle tree per accumulator
  // and written straight to W.acc[b] -- no cross-warp combine, fixed summation order.  Non-flat buckets are dealt
  // round-robin to the 8 warps.  The mu-free histogram terms are thread-strided; the scalars are reduced across warps.
  template <int MODE>
  __device__ __noinline__ void pass() {
    if (MODE == PASS_EMIT) return;                 // the compressed sweep recomputes mu from the eta parameters
    constexpr int NA = CompTraits<MODE>::NA;
    constexpr int NS = CompTraits<MODE>::NS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int T = L->T;
    const double sigma = W->sigma, theta = W->theta;
    const bool want_ll = W->want_ll != 0;
    __syncthreads();
    if (MODE == PASS_IRLS_START) {
      for (int idx = threadIdx.x; idx < (T + 1) * 5; idx += kCompThreads) { const int b = idx / 5, k = idx - b * 5; W->acc[b][k] = cg->ST[b][k]; }
      if (threadIdx.x == 0) {
        W->sc[0] = -gs->sum_ylogy + (double)cg->hdr.n_zero * theta * log1p(sigma / 6.0);
        W->sc[1] = (double)cg->hdr.n_zero;
        W->sc[2] = 0.0;
      }
      __syncthreads();
      (void)NA; (void)NS; (void)lane; (void)warp; (void)want_ll;
    } else {
      pass_points<MODE>();
    }
  }

  template <int MODE>
  __device__ __forceinline__ void pass_points() {
    constexpr int NA = CompTraits<MODE>::NA;
    constexpr int NS = CompTraits<MODE>::NS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int T = L->T;
    const double sigma = W->sigma, theta = W->theta;
    const bool want_ll = W->want_ll != 0;
    double sc[NSC];
#pragma unroll
    for (int k = 0; k < NSC; ++k) sc[k] = 0.0;
    const bool on_nodes = MODE == PASS_THETA && tn != nullptr && tn->on;     // block-uniform
    if (on_nodes) {
      const int ne = tn->n;
      double dummy[1];
      for (int e = threadIdx.x; e < ne; e += kCompThreads)
        point_contrib<MODE>(tn->en[e], tn->es[e], 0.0, 0.0, tn->eeta[e], 0.0, sigma, theta, want_ll, ft, dummy, sc);
    }
    // Since the bucket nodes a pass is short and its latency is what counts (the kernels wait at the barrier behind it), so the
    // work of a pass is spread as widely as the warp allows:
    //   flat buckets     one closed-form evaluation each, flat bucket f on lane f / kCompWarps of warp f % kCompWarps (all at once);
    //   node buckets     16 evaluations each, two buckets per warp step (one per half warp, half-warp shuffle trees);
    //   the rest         (slope too steep for the nodes in this pass, few points, exact-knot mode) walks its points, warp per bucket.
    // Every sum still runs in a fixed order -> bit-reproducible.
    const int half = lane >> 4, hl = lane & 15;
    auto node_pair = [&](int b0, int b1) {           // warp-uniform call; b1 may be -1
      const int mb = half ? b1 : b0;
      double acc[NACC];
#pragma unroll
      for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
      if (mb >= 0)
        point_contrib<MODE>(__ldg(&bn->nn[mb][hl]), bs[mb * kOffJ + hl], MODE == PASS_IRLS ? bq[mb * kOffJ + hl] : 0.0, __ldg(&bn->un[mb][hl]), W->a[mb], W->c[mb],
                            sigma, theta, want_ll, ft, acc, sc);
#pragma unroll
      for (int k = 0; k < NA; ++k) {
        double t = acc[k];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (hl == 0 && mb >= 0) W->acc[mb][k] = t;
      }
    };
    // The buckets are classified in parallel (lane l looks at buckets l and l + 32; a scalar loop over the <= 33 buckets in every
    // warp was half of the final stage's stall samples) and handed out by rank among their class: flat bucket number f to lane
    // f / kCompWarps of warp f % kCompWarps, node buckets 2p and 2p + 1 to warp p % kCompWarps, walked bucket number s to warp
    // s % kCompWarps.
    unsigned long long flat_m = 0ull, node_m = 0ull, walk_m = 0ull;
    if (!on_nodes) {
      int cls0 = 0, cls1 = 0;                        // 1 flat, 2 on its nodes, 3 walks its points
      if (lane <= T) { const double c = W->c[lane]; cls0 = c == 0.0 ? 1 : (on_bucket_nodes(lane, c) ? 2 : 3); }
      if (lane + 32 <= T) { const double c = W->c[lane + 32]; cls1 = c == 0.0 ? 1 : (on_bucket_nodes(lane + 32, c) ? 2 : 3); }
      flat_m = (unsigned long long)__ballot_sync(0xffffffffu, cls0 == 1) | ((unsigned long long)__ballot_sync(0xffffffffu, cls1 == 1) << 32);
      node_m = (unsigned long long)__ballot_sync(0xffffffffu, cls0 == 2) | ((unsigned long long)__ballot_sync(0xffffffffu, cls1 == 2) << 32);
      walk_m = (unsigned long long)__ballot_sync(0xffffffffu, cls0 == 3) | ((unsigned long long)__ballot_sync(0xffffffffu, cls1 == 3) << 32);
    }
    auto nth = [](unsigned long long m, int n) -> int {    // position of the n-th (0-based) set bit of m, -1 when there is none
      const unsigned lo = (unsigned)m, hi = (unsigned)(m >> 32);
      const int nlo = __popc(lo);
      if (n < nlo) return (int)__fns(lo, 0, n + 1);
      n -= nlo;
      if (n < __popc(hi)) return 32 + (int)__fns(hi, 0, n + 1);
      return -1;
    };
    const int n_node = __popcll(node_m), n_walk = __popcll(walk_m);
#pragma unroll 1
    for (int pr = warp; 2 * pr < n_node; pr += kCompWarps) node_pair(nth(node_m, 2 * pr), nth(node_m, 2 * pr + 1));
#pragma unroll 1
    for (int sidx = warp; sidx < n_walk; sidx += kCompWarps) {
      const int b = nth(walk_m, sidx);
      const double a = W->a[b], c = W->c[b];
      double acc[NACC];
#pragma unroll
      for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
      const int p1 = pbstart[b + 1];
      if (px) {
        const double refb = L->ref[b];
#pragma unroll kCompUnroll
        for (int d = pbstart[b] + lane; d < p1; d += 32)
          point_contrib<MODE>(pn[d], s1[d], MODE == PASS_IRLS ? q[d] : 0.0, px[d] - refb, a, c, sigma, theta, want_ll, ft, acc, sc);
      } else {
#pragma unroll kCompUnroll
        for (int d = pbstart[b] + lane; d < p1; d += 32)
          point_contrib<MODE>(pn[d], s1[d], MODE == PASS_IRLS ? q[d] : 0.0, pu[d], a, c, sigma, theta, want_ll, ft, acc, sc);
      }
#pragma unroll
      for (int k = 0; k < NA; ++k) {
        const double t = warp_sum(acc[k]);
        if (lane == 0) W->acc[b][k] = t;
      }
    }
    const int my_flat = nth(flat_m, lane * kCompWarps + warp);
    if (my_flat >= 0) {
      double fa[NACC];
#pragma unroll
      for (int k = 0; k < NACC; ++k) fa[k] = 0.0;
      flat_bucket_contrib<MODE>(L->U0[my_flat], L->U1[my_flat], L->U2[my_flat], cg->hdr.SB[my_flat], cg->hdr.SBu[my_flat], cg->hdr.SBuu[my_flat],
                                cg->hdr.QB[my_flat], W->a[my_flat], sigma, theta, want_ll, ft, fa, sc);
#pragma unroll
      for (int k = 0; k < NA; ++k) W->acc[my_flat][k] = fa[k];
    }
    // the mu-free histogram terms
    if (NS > 0 && MODE != PASS_SWEEP) {
      const int hmax = cg->hdr.ymax < kHistCap ? cg->hdr.ymax : kHistCap - 1;
      for (int k = threadIdx.x; k < hmax; k += kCompThreads)
        hist_contrib<MODE>(k, (double)tail[k], (double)tail[k + 1], sigma, theta, want_ll, sc);
      const int n_big = cg->hdr.n_big;
      for (int e = threadIdx.x; e < n_big; e += kCompThreads) big_contrib<MODE>(big[e], sigma, theta, want_ll, sc);
    }
    if (NS > 0) {
#pragma unroll
      for (int k = 0; k < NS; ++k) {
        const double t = warp_sum(sc[k]);
        if (lane == 0) scw[warp][k] = t;
      }
    }
    if (MODE == PASS_SWEEP)
      for (int b = threadIdx.x; b <= T; b += kCompThreads) { W->acc[b][5] = cg->hdr.SB[b]; W->acc[b][6] = cg->hdr.SBu[b]; }
    __syncthreads();
    if (NS > 0 && threadIdx.x < NS) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kCompWarps; ++w) s += scw[w][threadIdx.x];
      W->sc[threadIdx.x] = s;
    }
    __syncthreads();
  }
};

enum CompStage : int { CSTAGE_FORWARD = 0, CSTAGE_BACKWARD = 1, CSTAGE_FINAL = 2 };

template <int STAGE>
__global__ void __launch_bounds__(kCompThreads, STAGE == CSTAGE_FINAL ? SCLANE_COMP_MINB_FINAL : SCLANE_COMP_MINB) k_comp_stage(CompArgs A) {
  __shared__ Work W;
  __shared__ Lineage Ls;
  __shared__ GeneState st;
  __shared__ GeneStats gs;
  __shared__ CompGeneDev cg;
  __shared__ int pbs[NBMAX + 1];
  __shared__ double scw[kCompWarps][NSC];
  __shared__ FastTabs ftab;
  __shared__ TermTabs ttab;
  __shared__ double ctab[STAGE == CSTAGE_FINAL ? kOffJ : 1][kOffJ];
  __shared__ double bnuw_s[NBMAX];
  extern __shared__ double comp_dyn_smem[];                // CSTAGE_FINAL: the theta.ml node list (sizeof(ThetaNodes) bytes)
  if ((int)blockIdx.x >= A.n_genes) return;
  const int g = ((STAGE == CSTAGE_FINAL || STAGE == CSTAGE_BACKWARD) && A.order) ? A.order[blockIdx.x] : (int)blockIdx.x;
  if (A.cg[g].hdr.route == 0) return;                     // block-uniform: this gene takes the per-cell kernels
  ThetaNodes* tn = (STAGE == CSTAGE_FINAL && A.theta_nodes) ? reinterpret_cast<ThetaNodes*>(comp_dyn_smem) : nullptr;
  if (STAGE == CSTAGE_FINAL) {
    for (int i = threadIdx.x; i < kOffJ * kOffJ; i += kCompThreads) {
      const int m = i / kOffJ, j = i - m * kOffJ;
      ctab[m][j] = cospi((double)(m * (2 * j + 1)) / (double)(2 * kOffJ));
    }
    if (tn && threadIdx.x == 0) { tn->on = 0; tn->n = 0; }
  }
  {
    const int* src = reinterpret_cast<const int*>(A.lineage);
    int* dst = reinterpret_cast<int*>(&Ls);
    for (int i = threadIdx.x; i < (int)(sizeof(Lineage) / sizeof(int)); i += kCompThreads) dst[i] = src[i];
    const int* s2 = reinterpret_cast<const int*>(&A.out[g].st);
    int* d2 = reinterpret_cast<int*>(&st);
    for (int i = threadIdx.x; i < (int)(sizeof(GeneState) / sizeof(int)); i += kCompThreads) d2[i] = s2[i];
    const int* s3 = reinterpret_cast<const int*>(&A.stats[g]);
    int* d3 = reinterpret_cast<int*>(&gs);
    for (int i = threadIdx.x; i < (int)(sizeof(GeneStats) / sizeof(int)); i += kCompThreads) d3[i] = s3[i];
    const int* s4 = reinterpret_cast<const int*>(&A.cg[g]);
    int* d4 = reinterpret_cast<int*>(&cg);
    for (int i = threadIdx.x; i < (int)(sizeof(CompGeneDev) / sizeof(int)); i += kCompThreads) d4[i] = s4[i];
    int* w = reinterpret_cast<int*>(&W);
    for (int i = threadIdx.x; i < (int)(sizeof(Work) / sizeof(int)); i += kCompThreads) w[i] = 0;
    for (int i = threadIdx.x; i <= NBMAX; i += kCompThreads) pbs[i] = A.pbstart[i];
    for (int i = threadIdx.x; i < NBMAX; i += kCompThreads) bnuw_s[i] = (A.bn != nullptr && A.bn->elig[i]) ? A.bn->uw[i] : -1.0;
    fast_tabs_load(&ftab, threadIdx.x, kCompThreads);
  }
  __syncthreads();
  if (threadIdx.x == 0) work_attach_tabs(W, &ttab);
  __syncthreads();
  const long long t_start = clock64();
  CompDevExec ex{&W, &Ls, &cg, &gs, A.D, A.pn, A.pu, pbs, A.s1 + (size_t)g * (size_t)A.Dpad, A.q + (size_t)g * (size_t)A.Dpad,
                 A.tail + (size_t)g * (size_t)kHistCap, A.big + (size_t)g * (size_t)kBigCap, scw, &ftab, nullptr, tn, ctab};
  if (A.bn != nullptr) { ex.bn = A.bn; ex.bnuw = bnuw_s; ex.bs = A.bn_s + (size_t)g * (size_t)(NBMAX * kOffJ); ex.bq = A.bn_q + (size_t)g * (size_t)(NBMAX * kOffJ); }
  if (STAGE == CSTAGE_FORWARD) {
    for (int it = 0; it < A.n_iter; ++it) {
      gene_forward_fit(ex, Ls, gs, st, W);
      gene_sweep(ex, Ls, gs, st, W, A.M, A.tols_score);
    }
    __syncthreads();
    if (threadIdx.x == 0 && !st.fwd_done) { st.fwd_done = 1; st.fwd_stop = 4; }
  } else if (STAGE == CSTAGE_BACKWARD) {
    gene_backward(ex, Ls, gs, st, W);
    __syncthreads();
    if (threadIdx.x == 0) st.done |= 1;
  } else {
    gene_final(ex, Ls, gs, st, W, false, &A.out[g].alt, &A.out[g].null, 3);
    __syncthreads();
    if (threadIdx.x == 0) st.done |= 2;
  }
  __syncthreads();
  if (threadIdx.x == 0) { st.passes += W.passes; st.cycles[STAGE] += clock64() - t_start; }
  __syncthreads();
  {
    int* d2 = reinterpret_cast<int*>(&A.out[g].st);
    const int* s2 = reinterpret_cast<const int*>(&st);
    for (int i = threadIdx.x; i < (int)(sizeof(GeneState) / sizeof(int)); i += kCompThreads) d2[i] = s2[i];
  }
}

}  // namespace sclane
