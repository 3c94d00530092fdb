// sclane_device.cuh -- the CUDA executor of a pass and the per-stage kernels (sm_100a).
//
// Mapping: one CTA per gene (grid = genes of the batch, >> 148 SMs for every config of interest),
// 8 warps per CTA.  The gene's cells are sorted by pseudotime, so the <= 33 knot buckets are
// contiguous index ranges: each warp streams a contiguous run of 128-cell tiles with 128-bit
// coalesced loads (int4 counts, double2 x2 coordinates / offsets / means), keeps the running
// bucket's moments in registers, and only at a bucket boundary (<= 33 + 8 times per pass) does a
// warp-shuffle reduction into its private shared-memory slot.  Slots are combined across warps in a
// fixed order, so every sum is bit-reproducible run to run and across GPU counts.
#pragma once
#include <type_traits>
#include <cuda_runtime.h>

#include "sclane_comp.cuh"

namespace sclane {

// two independent cells per lane per iteration of the per-cell pass loops.  Measured on B200 (C4-shaped final stage, 8,000
// gene-lineage fits): 208 -> 163 ms once fast_exp() was made branch free (before that the rare-path call split the loop body
// into basic blocks and the two cells were not interleaved: no gain)
#ifndef SCLANE_PASS_ILP2
#define SCLANE_PASS_ILP2 1
#endif
#ifndef SCLANE_STAGE_MINB
#define SCLANE_STAGE_MINB 3
#endif
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int MODE> struct PassTraits;
template <> struct PassTraits<PASS_FIT>        { static constexpr int NA = 7, NS = 3; };
template <> struct PassTraits<PASS_IRLS>       { static constexpr int NA = 5, NS = 3; };
template <> struct PassTraits<PASS_IRLS_START> { static constexpr int NA = 5, NS = 3; };
template <> struct PassTraits<PASS_THETA>      { static constexpr int NA = 0, NS = 2; };
template <> struct PassTraits<PASS_SWEEP>      { static constexpr int NA = 7, NS = 0; };
template <> struct PassTraits<PASS_EMIT>       { static constexpr int NA = 0, NS = 0; };

// HYB = false: the per-cell executor (every term per cell, count-indexed special-function tables).
// HYB = true:  the hybrid executor of the fits WITH an offset (alternatives with hinges): the mu-dependent terms per cell --
//              mu_i = exp(a_b + c_b u_i + o_i) differs from cell to cell -- but every mu-free term (psi / trigamma / lgamma
//              differences, (y + theta) log1p(sigma y)) from the gene's count histogram, exactly as the X-compressed passes do
//              (sclane_comp.cuh): no per-pass tables, no table look-ups and no large-count branch in the cell loop.
// Cluster barrier (all threads of all CTAs of the cluster): release / acquire, so that shared-memory writes made before it
// are visible to remote reads (DSMEM) made after it.
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_cta_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// generic address of `p` (a shared-memory address of this CTA) in the shared memory of CTA `rank` of the cluster
__device__ __forceinline__ const double* cluster_map(const double* p, unsigned rank) {
  unsigned long long out;
  asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((unsigned long long)p), "r"(rank));
  return reinterpret_cast<const double*>(out);
}

constexpr int kPartN = NBMAX * NACC + NSC;     // one CTA's partial sums of a pass (cluster mode)

// theta.ml of a fit WITH a size-factor offset on Chebyshev nodes (hybrid final stage; same idea as ThetaNodes in sclane_comp.cuh).
// mu is fixed while theta.ml iterates and both of its sums are  sum_i [A(z_i; theta) + y_i B(z_i; theta)]  with
// z_i = a_b + c_b u_i + o_i.  The offset quadrature's bin-ordered cell list (oord: by offset bin, then pseudotime) makes the
// cells of (offset bin k, knot bucket b) a contiguous, pseudotime-sorted range; inside it z spans at most |c_b| du + w (w <= 1/2
// the bin width), so the range -- cut further along u when the slope is steep -- is one interval of width <= kThBinWidth in z and
// its cells collapse into 16 node weights (Chebyshev moments of 1 and of y, one warp per interval, fixed summation order).
// A Newton step then costs  (#bins x #buckets) x 16  node evaluations instead of N cell evaluations.
constexpr int kHtUnits = 512;          // (offset bin, bucket) groups the unit table holds
constexpr int kHtCap = 4096;           // entries of the node list (global scratch, 3 doubles each)
struct HybTheta {                      // shared memory
  int on, n;
  int first[kHtUnits + 1];
  short sub[kHtUnits];                 // 0 = empty group, -1 = its cells as they are, S >= 1 = S intervals along u
  int wt[kWarps + 1];
  double ctab[kOffJ][kOffJ];           // T_m(xi_j)
};
static_assert(kHtUnits == 2 * kThreads, "theta_prepare() lays out two groups per thread");
struct HybThetaArgs {
  const OffsetNodes* on;
  const int* oord;                     // [N] sorted-cell positions by (offset bin, position)
  const double* off_ord;               // [N] offsets in that order
  const int* ogrp;                     // [K (T + 2)] group boundaries inside oord
  double* en; double* es; double* eeta;      // [kHtCap] each: this gene's node list
};

// CL > 1: a thread-block CLUSTER of CL CTAs fits one gene.  Every CTA keeps its own copy of the gene's state and runs the
// identical control flow (the serial algebra is redundant, deterministic and therefore identical); a pass splits the cells
// over the CL x 8 warps of the cluster and the CTAs exchange their partial moments through distributed shared memory:
// each CTA writes part[buf], ONE cluster barrier, each CTA sums the CL partials in rank order (-> bit-identical in every
// CTA and for every launch).  part[] is double buffered, which makes a second barrier per pass unnecessary.
template <bool HYB, int CL = 1>
struct DevExecT {
  static constexpr bool kTabs = true;      // term tables attached to Work (sclane_core.cuh term_coef)
  Work* W;
  const Lineage* L;
  const int* __restrict__ y;        // sorted counts of this gene, row padded to Npad
  const double* __restrict__ u;     // [Npad] bucket-local coordinate (shared by all genes)
  const double* __restrict__ off;   // [Npad] offset or nullptr
  double* mu;                       // [Npad] fitted means of this gene (sweep input / emit output)
  double (*slot)[NBMAX][NACC];      // [kWarps] warp-private partial moments
  double (*scw)[NSC];               // [kWarps] warp-private partial scalars
  PassTabs* tb;                     // count-indexed special-function tables of the current pass
  const FastTabs* ft;               // exp / log1p tables in shared memory (sclane_fastmath.cuh)
  // hybrid mode only: the gene's aggregates of k_aggregate
  const CompGeneDev* cg;
  const GeneStats* gs;
  const int* __restrict__ tail;
  const int* __restrict__ big;
  const double* wtab;               // [64] glm.fit start weights by count (shared memory)
  const double* wztab;
  const double* uoff;               // exact-knot mode: u[] is relative to the lineage's smallest abscissa and the buckets are the gene's
                                    // own, so the bucket-local coordinate is u[i] - uoff[bucket]; nullptr otherwise
  double* part;                     // cluster mode: [2][kPartN] this CTA's partial sums (shared memory), else nullptr
  int crank;                        // cluster mode: rank of this CTA in its cluster
  int pbuf;                         // cluster mode: which half of part[] the running pass writes
  HybTheta* ht = nullptr;           // hybrid mode, CL == 1: theta.ml node list bookkeeping (shared memory) or nullptr
  HybThetaArgs hta = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

  // theta.ml is about to iterate at the mean defined by W.a / W.c (+ offset): condense the cells into the node list
  __device__ __noinline__ void theta_prepare() {
    if (!HYB || CL > 1 || ht == nullptr) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Lineage& Lr = *L;
    const int T = Lr.T;
    const OffsetNodes& O = *hta.on;
    const int K = O.K, nU = K * (T + 1);
    const double ow = O.w;
    __syncthreads();
    if (!(W->use_offset && off != nullptr) || nU > kHtUnits || !(ow <= kOffBinWidth + 1e-9)) {
      if (threadIdx.x == 0) ht->on = 0;
      __syncthreads();
      return;
    }
    // entries per group: two consecutive groups per thread (kHtUnits = 2 x kThreads), exclusive scan in group order
    const double room = kThBinWidth - ow;               // what the slope may use of the interval's width
    int cnt[2] = {0, 0};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int unit = 2 * (int)threadIdx.x + j;
      if (unit >= nU) continue;
      const int k = unit / (T + 1), b = unit - k * (T + 1);
      const int r0 = hta.ogrp[k * (T + 2) + b], r1 = hta.ogrp[k * (T + 2) + b + 1], P = r1 - r0;
      int S = 0;
      if (P > 0) {
        const double c = W->c[b];
        const double du = c != 0.0 ? u[hta.oord[r1 - 1]] - u[hta.oord[r0]] : 0.0;
        const double span = fabs(c) * du;
        S = (int)ceil(span / room);
        if (S < 1) S = 1;
        if (!(span == span) || S > 256 || 3 * S * kOffJ > 2 * P) { S = -1; cnt[j] = P; }
        else cnt[j] = S * kOffJ;
      }
      ht->sub[unit] = (short)S;
    }
    {
      const int v = cnt[0] + cnt[1];
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
      if (lane == 31) ht->wt[warp] = inc;
      __syncthreads();
      if (threadIdx.x == 0) {
        int runw = 0;
        for (int w = 0; w < kWarps; ++w) { const int t = ht->wt[w]; ht->wt[w] = runw; runw += t; }
        ht->wt[kWarps] = runw;
        ht->n = runw;
        ht->on = (runw <= kHtCap && 2 * runw <= Lr.N) ? 1 : 0;       // the list must fit and must at least halve the work
      }
      __syncthreads();
      const int base = ht->wt[warp] + inc - v;
      if (2 * (int)threadIdx.x < nU) ht->first[2 * threadIdx.x] = base;
      if (2 * (int)threadIdx.x + 1 < nU) ht->first[2 * threadIdx.x + 1] = base + cnt[0];
    }
    __syncthreads();
    if (!ht->on) return;
    // build: one group per warp at a time
#pragma unroll 1
    for (int unit = warp; unit < nU; unit += kWarps) {
      const int S = ht->sub[unit];
      if (S == 0) continue;
      const int k = unit / (T + 1), b = unit - k * (T + 1);
      const int r0 = hta.ogrp[k * (T + 2) + b], r1 = hta.ogrp[k * (T + 2) + b + 1];
      const double a = W->a[b], c = W->c[b];
      const int e0 = ht->first[unit];
      if (S < 0) {
        for (int r = r0 + lane; r < r1; r += 32) {
          const int i = hta.oord[r];
          const int e = e0 + (r - r0);
          hta.en[e] = 1.0; hta.es[e] = (double)y[i]; hta.eeta[e] = a + c * u[i] + hta.off_ord[r];
        }
        continue;
      }
      const double olo = O.omin + (double)k * ow, ohi = olo + ow;
      const double ulo = u[hta.oord[r0]], uhi = u[hta.oord[r1 - 1]];
      const double hu = (uhi - ulo) / (double)S;
      const int half = lane >> 4, j = lane & 15;
#pragma unroll 1
      for (int sidx = 0; sidx < S; ++sidx) {
        const double u0 = ulo + (double)sidx * hu, u1 = sidx + 1 == S ? uhi : ulo + (double)(sidx + 1) * hu;
        int lo = r0, hi = r1;
        if (sidx > 0) { int x = r0, yv = r1; while (x < yv) { const int m = (x + yv) >> 1; if (u[hta.oord[m]] < u0) x = m + 1; else yv = m; } lo = x; }
        if (sidx + 1 < S) { int x = r0, yv = r1; while (x < yv) { const int m = (x + yv) >> 1; if (u[hta.oord[m]] < u1) x = m + 1; else yv = m; } hi = x; }
        const double zA = a + c * u0, zB = a + c * u1;
        const double zmin = fmin(zA, zB) + olo, zmax = fmax(zA, zB) + ohi;
        const double zw = zmax - zmin;
        const double sc2 = zw > 0.0 ? 2.0 / zw : 0.0;
        double cm[kOffJ];
#pragma unroll
        for (int m = 0; m < kOffJ; ++m) cm[m] = 0.0;
        for (int r = lo + j; r < hi; r += 16) {
          const int i = hta.oord[r];
          const double z = a + c * u[i] + hta.off_ord[r];
          const double v = half ? (double)y[i] : 1.0;
          double xi = sc2 * (z - zmin) - 1.0;
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
        for (int m = kOffJ - 1; m >= 1; --m) sacc += ht->ctab[m][j] * cm[m];
        const double wgt = (cm[0] + 2.0 * sacc) * (1.0 / (double)kOffJ);
        const int e = e0 + sidx * kOffJ + j;
        if (half) hta.es[e] = wgt;
        else { hta.en[e] = wgt; hta.eeta[e] = zmin + 0.5 * (ht->ctab[1][j] + 1.0) * zw; }
      }
    }
    __syncthreads();
  }

  // a Newton step of theta.ml on the node list built by theta_prepare()
  __device__ __noinline__ void pass_theta_nodes() {
    constexpr int NS = PassTraits<PASS_THETA>::NS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double sigma = W->sigma, theta = W->theta;
    if (lane < NSC) scw[warp][lane] = 0.0;
    __syncwarp();
    double sc[NSC], dummy[1];
#pragma unroll
    for (int k = 0; k < NSC; ++k) sc[k] = 0.0;
    const int n = ht->n;
    for (int e = threadIdx.x; e < n; e += kThreads)
      point_contrib<PASS_THETA>(hta.en[e], hta.es[e], 0.0, 0.0, hta.eeta[e], 0.0, sigma, theta, false, ft, dummy, sc);
    const int ymax = cg->hdr.ymax;
    const int hmax = ymax < kHistCap ? ymax : kHistCap - 1;
    for (int k = threadIdx.x; k < hmax; k += kThreads) hist_contrib<PASS_THETA>(k, (double)tail[k], (double)tail[k + 1], sigma, theta, false, sc);
    const int n_big = cg->hdr.n_big;
    for (int e = threadIdx.x; e < n_big; e += kThreads) big_contrib<PASS_THETA>(big[e], sigma, theta, false, sc);
    flush_uniform<NS>(sc, 0, &scw[warp][0], 0);
    combine<0, NS>(L->T);
  }

  __device__ __forceinline__ bool leader() const { return threadIdx.x == 0; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  __device__ __forceinline__ PassTabs* tabs() const { return HYB ? nullptr : tb; }
  template <class F>
  __device__ __forceinline__ void par_for(int n, F f) const {
    for (int i = threadIdx.x; i < n; i += kThreads) f(i);
  }

  // Register accumulators -> the warp's shared-memory slot.  flush_uniform: every lane belongs to bucket b (shuffle
  // tree, lane 0 adds).  flush_mixed: the lanes of a boundary group belong to buckets b_lo..b_hi (usually two): one
  // masked shuffle tree per bucket.  Both run <= ~T + 1 times per pass per warp and sum in a fixed order.
  template <int NV>
  __device__ __noinline__ void flush_uniform(const double* v, int b, double* dst, int stride) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const double t = warp_sum(v[k]);
      if (lane == 0) dst[b * stride + k] += t;
    }
  }
  template <int NV>
  __device__ __noinline__ void flush_mixed(const double* v, int my_b, bool valid, int b_lo, int b_hi, double* dst, int stride) {
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int bb = b_lo; bb <= b_hi; ++bb) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const double t = warp_sum((valid && my_b == bb) ? v[k] : 0.0);
        if (lane == 0) dst[bb * stride + k] += t;
      }
    }
  }

  // Cross-warp combination of the warp-private partial sums, in a fixed order (bit-reproducible).
  template <int NA, int NS>
  __device__ __forceinline__ void combine(int T) {
    __syncthreads();
    if (CL > 1) {
      double* mine = part + pbuf * kPartN;
      if (NA > 0) {
        for (int i = threadIdx.x; i < (T + 1) * NACC; i += kThreads) {
          double s = 0.0;
#pragma unroll
          for (int w = 0; w < kWarps; ++w) s += (&slot[w][0][0])[i];
          mine[i] = s;
        }
      }
      if (NS > 0 && threadIdx.x < NS) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += scw[w][threadIdx.x];
        mine[NBMAX * NACC + threadIdx.x] = s;
      }
      cluster_sync_all();
      if (NA > 0) {
        double* out = &W->acc[0][0];
        for (int i = threadIdx.x; i < (T + 1) * NACC; i += kThreads) {
          double v[CL];
#pragma unroll
          for (int r = 0; r < CL; ++r) v[r] = cluster_map(mine, r)[i];
          double s = 0.0;
#pragma unroll
          for (int r = 0; r < CL; ++r) s += v[r];
          out[i] = s;
        }
      }
      if (NS > 0 && threadIdx.x < NS) {
        double v[CL];
#pragma unroll
        for (int r = 0; r < CL; ++r) v[r] = cluster_map(mine, r)[NBMAX * NACC + threadIdx.x];
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < CL; ++r) s += v[r];
        W->sc[threadIdx.x] = s;
      }
      pbuf ^= 1;
      __syncthreads();
      return;
    }
    if (NA > 0) {
      double* out = &W->acc[0][0];
      for (int i = threadIdx.x; i < (T + 1) * NACC; i += kThreads) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += (&slot[w][0][0])[i];
        out[i] = s;
      }
    }
    if (NS > 0 && threadIdx.x < NS) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += scw[w][threadIdx.x];
      W->sc[threadIdx.x] = s;
    }
    __syncthreads();
  }

  // one cell of a pass
  template <int MODE>
  __device__ __forceinline__ void cell(int yi, double ui, double oi, double a, double c, double sigma, double theta, bool want_ll,
                                       double* acc, double* sc, double& mo) const {
    if (HYB) {
      if (MODE == PASS_IRLS_START) {
        double w, wz;
        if (yi < 64) { w = wtab[yi]; wz = wztab[yi]; } else start_cell(yi, w, wz);
        wz -= w * oi;                                  // z = (log(mustart) - offset) + (y - mu)/mu
        acc[0] += w; acc[1] += w * ui; acc[2] += w * ui * ui;
        acc[3] += wz; acc[4] += wz * ui;
      } else if (MODE == PASS_IRLS || MODE == PASS_THETA) {
        const double y = (double)yi;
        point_contrib<MODE>(1.0, y, y * y, ui, a, c, sigma, theta, want_ll, ft, acc, sc, oi);
      }
    } else {
      cell_contrib<MODE>(yi, ui, oi, a, c, sigma, theta, *tb, ft, 0.0, want_ll, acc, sc, mo);
    }
  }

  // FP64-bound passes (fits, theta.ml, emit): ONE compact copy of the per-cell math, one cell per lane
  // per iteration (a warp still reads 32 consecutive cells: 128 B of counts, 256 B of coordinates).
  // The loop body must stay small: these kernels were instruction-fetch bound when it was unrolled.
  template <int MODE>
  __device__ __noinline__ void pass_scalar() {
    constexpr int NA = PassTraits<MODE>::NA;
    constexpr int NS = PassTraits<MODE>::NS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Lineage& Lr = *L;
    const int N = Lr.N, T = Lr.T;
    const double sigma = W->sigma, theta = W->theta;
    const bool want_ll = W->want_ll != 0;
    const bool use_off = (MODE == PASS_IRLS || MODE == PASS_IRLS_START || MODE == PASS_THETA) && W->use_offset && off != nullptr;
    double* myslot = &slot[warp][0][0];
    if (NA > 0) {
      for (int i = lane; i < (T + 1) * NACC; i += 32) myslot[i] = 0.0;
    }
    if (NS > 0 && lane < NSC) scw[warp][lane] = 0.0;
    __syncwarp();
    double acc[NACC], sc[NSC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
#pragma unroll
    for (int k = 0; k < NSC; ++k) sc[k] = 0.0;
#if SCLANE_PASS_ILP2
    double acc2[NACC], sc2[NSC];                       // second, independent accumulation chain
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc2[k] = 0.0;
#pragma unroll
    for (int k = 0; k < NSC; ++k) sc2[k] = 0.0;
#endif
    // contiguous run of 32-cell groups per warp
    const int ngroups = (N + 31) >> 5;
    const int gw = CL > 1 ? crank * kWarps + warp : warp;      // this warp among the CL x 8 warps that share the gene
    const int g0 = (int)(((long long)ngroups * gw) / (kWarps * CL));
    const int g1 = (int)(((long long)ngroups * (gw + 1)) / (kWarps * CL));
    if (g0 < g1) {
      int cur_b = 0;                       // bucket the register accumulators belong to (warp uniform)
      const int first = g0 << 5;
      while (cur_b < T && Lr.bstart[cur_b + 1] <= first) ++cur_b;
      int b_hi = Lr.bstart[cur_b + 1];
      double a = W->a[cur_b], c = W->c[cur_b];
      double uo = uoff ? uoff[cur_b] : 0.0;
      // constant linear predictor inside the bucket (no slope term, no offset): hoist exp / log1p / reciprocals
      constexpr bool kCanFlat = (MODE == PASS_FIT || MODE == PASS_IRLS || MODE == PASS_THETA || MODE == PASS_EMIT);
      bool flat = !HYB && kCanFlat && !use_off && c == 0.0;
      FlatBucket fb;
      if (!HYB) fb = make_flat_bucket(flat ? a : 0.0, sigma);
      int g = g0;
#pragma unroll 1
      while (g < g1) {
        const int i = (g << 5) + lane;
#if SCLANE_PASS_ILP2
        if (g + 1 < g1 && (g << 5) + 64 <= b_hi) {
          // two groups inside the running bucket (b_hi <= N: every lane valid): two independent cells per lane
          const int y0 = __ldg(y + i), y1 = __ldg(y + i + 32);
          const double u0 = __ldg(u + i) - uo, u1 = __ldg(u + i + 32) - uo;
          const double o0 = use_off ? __ldg(off + i) : 0.0, o1 = use_off ? __ldg(off + i + 32) : 0.0;
          double m0, m1;
          cell<MODE>(y0, u0, o0, a, c, sigma, theta, want_ll, acc, sc, m0);
          cell<MODE>(y1, u1, o1, a, c, sigma, theta, want_ll, acc2, sc2, m1);
          if (MODE == PASS_EMIT) { mu[i] = m0; mu[i + 32] = m1; }
          g += 2;
          continue;
        }
#endif
        if ((g << 5) + 32 <= b_hi) {
          // the whole group lies inside the running bucket (b_hi <= N, so every lane is valid)
          const int yi = __ldg(y + i);
          const double ui = __ldg(u + i) - uo;
          double mo;
          if (flat) {
            if (!HYB) cell_contrib_flat<MODE>(yi, ui, fb, sigma, theta, *tb, want_ll, acc, sc, mo);
          } else {
            const double oi = use_off ? __ldg(off + i) : 0.0;
            cell<MODE>(yi, ui, oi, a, c, sigma, theta, want_ll, acc, sc, mo);
          }
          if (MODE == PASS_EMIT) mu[i] = mo;
          g += 1;
          continue;
        }
        // the group straddles a bucket boundary and/or the end of the gene (<= T + 1 times per pass)
        const bool valid = i < N;
        const int yi = __ldg(y + i);                   // rows are padded: always in bounds
        double ui = __ldg(u + i);
        const double oi = use_off ? __ldg(off + i) : 0.0;
        int b = cur_b;                                 // this lane's bucket (cells are sorted: b >= cur_b)
        while (valid && b < T && i >= Lr.bstart[b + 1]) ++b;
        if (uoff) ui -= uoff[b];
        double ct[NACC], mo = 0.0;
#pragma unroll
        for (int k = 0; k < NACC; ++k) ct[k] = 0.0;
        if (valid) {
          cell<MODE>(yi, ui, oi, W->a[b], W->c[b], sigma, theta, want_ll, ct, sc, mo);
          if (MODE == PASS_EMIT) mu[i] = mo;
        }
        if (NA > 0) {
#if SCLANE_PASS_ILP2
#pragma unroll
          for (int k = 0; k < NA; ++k) { acc[k] += acc2[k]; acc2[k] = 0.0; }
#endif
          flush_uniform<NA>(acc, cur_b, myslot, NACC);
#pragma unroll
          for (int k = 0; k < NA; ++k) acc[k] = 0.0;
        }
        int b_last = valid ? b : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b_last = max(b_last, __shfl_xor_sync(0xffffffffu, b_last, o));
        if (NA > 0) flush_mixed<NA>(ct, b, valid, cur_b, b_last, myslot, NACC);
        cur_b = b_last;
        b_hi = Lr.bstart[cur_b + 1];
        a = W->a[cur_b]; c = W->c[cur_b];
        uo = uoff ? uoff[cur_b] : 0.0;
        flat = !HYB && kCanFlat && !use_off && c == 0.0;
        if (flat) fb = make_flat_bucket(a, sigma);
        g += 1;
      }
#if SCLANE_PASS_ILP2
#pragma unroll
      for (int k = 0; k < NACC; ++k) acc[k] += acc2[k];
#pragma unroll
      for (int k = 0; k < NSC; ++k) sc[k] += sc2[k];
#endif
      if (NA > 0) flush_uniform<NA>(acc, cur_b, myslot, NACC);
    }
    if (HYB && NS > 0 && (MODE == PASS_IRLS || MODE == PASS_THETA)) {
      // the mu-free terms of the pass from the gene's tail counts and its list of large counts (sclane_comp.cuh)
      const int ymax = cg->hdr.ymax;
      const int hmax = ymax < kHistCap ? ymax : kHistCap - 1;
      const int gt = CL > 1 ? crank * kThreads + (int)threadIdx.x : (int)threadIdx.x;
      for (int k = gt; k < hmax; k += kThreads * CL) hist_contrib<MODE>(k, (double)tail[k], (double)tail[k + 1], sigma, theta, want_ll, sc);
      const int n_big = cg->hdr.n_big;
      for (int e = gt; e < n_big; e += kThreads * CL) big_contrib<MODE>(big[e], sigma, theta, want_ll, sc);
    }
    if (NS > 0) flush_uniform<NS>(sc, 0, &scw[warp][0], 0);
    combine<NA, NS>(T);
    if (HYB && MODE == PASS_IRLS_START) {
      // deviance and theta.ml start sums at mustart do not involve the offset: gene constants (as the compressed start pass)
      if (threadIdx.x == 0) {
        W->sc[0] = -gs->sum_ylogy + (double)cg->hdr.n_zero * theta * log1p(sigma / 6.0);
        W->sc[1] = (double)cg->hdr.n_zero;
        W->sc[2] = 0.0;
      }
      __syncthreads();
    }
  }

  template <int MODE>
  __device__ __forceinline__ void pass() {
    static_assert(MODE != PASS_SWEEP, "the sweep streams through k_sweep_moments, not through a CTA pass");
    if (HYB && CL == 1 && MODE == PASS_THETA) {
      if (ht != nullptr && ht->on) { pass_theta_nodes(); return; }      // block-uniform
    }
    pass_scalar<MODE>();
  }
};
using DevExec = DevExecT<false>;

// ---------------------------------------------------------------------------------------------
// kernel arguments
// ---------------------------------------------------------------------------------------------
struct GeneOutDev {
  GeneState st;
  NbResult alt, null;
};

struct StageArgs {
  const Lineage* lineage;      // device copy
  const int* ys;               // [B][Npad] sorted counts
  const double* u;             // [Npad]
  const double* off;           // [Npad] or nullptr
  double* mu;                  // [B][Npad]
  const GeneStats* stats;      // [B]
  GeneOutDev* out;             // [B]
  double* scores;              // optional [B][2][T] dump of the sweep scores (k_sweep_select), or nullptr
  double* mom;                 // [B][NBMAX][NACC] sweep moments (k_sweep_moments -> k_sweep_select)
  int n_genes;
  int M;
  double tols_score;
  int use_offset;
  // hybrid final stage (k_stage<STAGE_FINAL, true>): the gene aggregates of k_aggregate
  const CompGeneDev* cg;       // [B]
  const int* tail;             // [B][kHistCap]
  const int* big;              // [B][kBigCap]
  int exact;                   // approx.knot = FALSE: the buckets are the gene's own (GeneState.ex_*); needs pstart
  const int* pstart;           // [D + 1] first sorted cell of every abscissa
  const int* order;            // optional CTA -> gene map of the final stage (k_order_final: expensive genes first), or nullptr
  // hybrid final stage: theta.ml on Chebyshev nodes (HybTheta above); all null = theta.ml walks the cells
  const OffsetNodes* on = nullptr;
  const int* oord = nullptr;
  const double* off_ord = nullptr;
  const int* ogrp = nullptr;
  double* ht_scratch = nullptr;  // [B][3][kHtCap]
};

enum Stage : int { STAGE_FWD_FIT = 0, STAGE_BACKWARD = 2, STAGE_FINAL = 3 };

template <int STAGE, bool HYB = false, int CL = 1>
__global__ void __launch_bounds__(kThreads, SCLANE_STAGE_MINB) k_stage(StageArgs A) {
  __shared__ Work W;
  __shared__ Lineage Ls;
  __shared__ GeneState st;
  __shared__ GeneStats gs;
  __shared__ double slot[kWarps][NBMAX][NACC];
  __shared__ double scw[kWarps][NSC];
  __shared__ PassTabs tabs;
  __shared__ FastTabs ftab;
  __shared__ TermTabs ttab;
  __shared__ double wtab[HYB ? 64 : 1], wztab[HYB ? 64 : 1];
  __shared__ double uoff_s[NBMAX + 1];
  __shared__ double part_s[CL > 1 ? 2 * kPartN : 1];
  __shared__ NbResult dump_s[CL > 1 ? 2 : 1];        // cluster mode: CTAs other than rank 0 keep their (identical) results here
  constexpr bool kHt = HYB && CL == 1 && STAGE == STAGE_FINAL;
  using HtStore = typename std::conditional<kHt, HybTheta, int>::type;      // only the hybrid final stage pays for it
  __shared__ HtStore ht_store;
  HybTheta* const htp = kHt ? reinterpret_cast<HybTheta*>(&ht_store) : nullptr;
  static_assert(CL == 1 || (HYB && STAGE == STAGE_FINAL), "clusters are used by the hybrid final stage only");
  const int slot_g = (int)blockIdx.x / CL;            // every exit below is uniform over the cluster (gene-level conditions)
  const int crank = CL > 1 ? (int)cluster_cta_rank() : 0;
  if (slot_g >= A.n_genes) return;
  const int g = ((STAGE == STAGE_FINAL || STAGE == STAGE_BACKWARD) && A.order) ? A.order[slot_g] : slot_g;
  if (HYB && A.cg[g].hdr.route == 0) return;          // block-uniform: this gene has no histogram (plain per-cell kernel fits it)
  {
    if (HYB && threadIdx.x < 64) start_cell(threadIdx.x, wtab[threadIdx.x], wztab[threadIdx.x]);
    // cooperative word copies of the small per-lineage / per-gene structs
    const int* src = reinterpret_cast<const int*>(A.lineage);
    int* dst = reinterpret_cast<int*>(&Ls);
    for (int i = threadIdx.x; i < (int)(sizeof(Lineage) / sizeof(int)); i += kThreads) dst[i] = src[i];
    const int* s2 = reinterpret_cast<const int*>(&A.out[g].st);
    int* d2 = reinterpret_cast<int*>(&st);
    for (int i = threadIdx.x; i < (int)(sizeof(GeneState) / sizeof(int)); i += kThreads) d2[i] = s2[i];
    const int* s3 = reinterpret_cast<const int*>(&A.stats[g]);
    int* d3 = reinterpret_cast<int*>(&gs);
    for (int i = threadIdx.x; i < (int)(sizeof(GeneStats) / sizeof(int)); i += kThreads) d3[i] = s3[i];
    int* w = reinterpret_cast<int*>(&W);
    for (int i = threadIdx.x; i < (int)(sizeof(Work) / sizeof(int)); i += kThreads) w[i] = 0;
    fast_tabs_load(&ftab, threadIdx.x, kThreads);
    if (kHt && A.ht_scratch) {
      for (int i = threadIdx.x; i < kOffJ * kOffJ; i += kThreads) {
        const int m = i / kOffJ, j = i - m * kOffJ;
        htp->ctab[m][j] = cospi((double)(m * (2 * j + 1)) / (double)(2 * kOffJ));
      }
      if (threadIdx.x == 0) { htp->on = 0; htp->n = 0; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    work_attach_tabs(W, &ttab);
    if (A.exact) {
      // the gene's own buckets from the knots its forward pass selected (sclane_exact_device.cuh); u[] is relative to x0
      const double x0 = Ls.ref[0];
      const int T = st.ex_T;
      Ls.T = T;
      for (int k = 0; k < T; ++k) Ls.knots[k] = st.ex_knots[k];
      Ls.ref[0] = T > 0 ? Ls.knots[0] : x0;
      for (int b = 1; b <= T; ++b) Ls.ref[b] = Ls.knots[b - 1];
      Ls.bstart[0] = 0;
      for (int b = 1; b <= T; ++b) Ls.bstart[b] = A.pstart[st.ex_pidx[b - 1]];
      for (int b = T + 1; b <= NBMAX; ++b) Ls.bstart[b] = Ls.N;
      for (int b = 0; b <= T; ++b) uoff_s[b] = Ls.ref[b] - x0;
    }
  }
  __syncthreads();
  if (STAGE == STAGE_BACKWARD && (st.done & 1)) return;      // block-uniform: already handled by k_comp_stage
  if (STAGE == STAGE_FINAL && (st.done & 2)) return;
  const size_t row = (size_t)g * (size_t)Ls.Npad;
  const long long t_start = clock64();
  DevExecT<HYB, CL> ex{&W, &Ls, A.ys + row, A.u, A.off, A.mu ? A.mu + row : nullptr, slot, scw, &tabs, &ftab,
                       HYB ? A.cg + g : nullptr, &gs, HYB ? A.tail + (size_t)g * (size_t)kHistCap : nullptr, HYB ? A.big + (size_t)g * (size_t)kBigCap : nullptr,
                       wtab, wztab, A.exact ? uoff_s : nullptr, CL > 1 ? part_s : nullptr, crank, 0};
  if (kHt && A.ht_scratch && !A.exact) {
    double* sc0 = A.ht_scratch + (size_t)g * 3 * (size_t)kHtCap;
    ex.ht = htp;
    ex.hta = HybThetaArgs{A.on, A.oord, A.off_ord, A.ogrp, sc0, sc0 + kHtCap, sc0 + 2 * kHtCap};
  }
  if (STAGE == STAGE_FWD_FIT) {
    gene_forward_fit(ex, Ls, gs, st, W);
  } else if (STAGE == STAGE_BACKWARD) {
    gene_backward(ex, Ls, gs, st, W);
  } else if (STAGE == STAGE_FINAL) {
    // genes whose intercept-only fits already ran on the offset quadrature (done bit 2) only need the alternative with hinges
    NbResult* alt_out = crank == 0 ? &A.out[g].alt : &dump_s[0];
    NbResult* null_out = crank == 0 ? &A.out[g].null : &dump_s[CL > 1 ? 1 : 0];
    gene_final(ex, Ls, gs, st, W, A.use_offset != 0, alt_out, null_out, (st.done & 4) ? 2 : 3);
    __syncthreads();
    if (threadIdx.x == 0) st.done |= 2;
  }
  __syncthreads();
  if (threadIdx.x == 0) { st.passes += W.passes; st.cycles[STAGE == STAGE_FWD_FIT ? 0 : (STAGE == STAGE_BACKWARD ? 1 : 2)] += clock64() - t_start; }
  __syncthreads();
  if (crank == 0) {
    int* d2 = reinterpret_cast<int*>(&A.out[g].st);
    const int* s2 = reinterpret_cast<const int*>(&st);
    for (int i = threadIdx.x; i < (int)(sizeof(GeneState) / sizeof(int)); i += kThreads) d2[i] = s2[i];
  }
  if (CL > 1) cluster_sync_all();                      // no CTA leaves while another may still read its partial sums
}

// ---------------------------------------------------------------------------------------------
// K2a: the score sweep's streaming kernel.  One CTA per gene; reads the gene's sorted counts (int32) and fitted
// means (fp64) once -- 12 B per cell, the algorithmic traffic of score_fun_glm -- and reduces them
// to the per-bucket moments  sum w, sum w u, sum w u^2, sum r, sum r u  (w = mu/(1+mu sigma), r = (y-mu)/(1+mu sigma);
// stat_out_score_glm_null.R:27-30, score_fun_glm.R:40,44) and, on the first forward iteration (WITH_Y), sum y, sum y u
// for the OLS check of stat_out().  Nothing else happens here: the O(T p^2) algebra is k_sweep_select's.
//   * 128-bit coalesced streaming loads (ld.global.nc.L1::no_allocate) for counts and means; the bucket-local
//     coordinate u is shared by every gene and stays L1/L2 resident;
//   * four cells per lane per tile, bucket moments in registers, flushed at the <= T bucket boundaries.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int4 ld_stream_i4(const int* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ double2 ld_stream_d2(const double* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

// 1/x for x >= 1 (x = 1 + sigma mu): hardware seed + two Newton steps, no special-case handling needed.
__device__ __forceinline__ double rcp_ge1(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// Tuned on B200 with tools/sweep_ablation.sh (profiles/r01_sweep_ablation.txt): 4 warps per CTA at 8 CTAs/SM, no software
// prefetch and the plain masked boundary loop beat every other combination tried at 10,000 cells per gene.
#ifndef SCLANE_SWEEP_WARPS
#define SCLANE_SWEEP_WARPS 4
#endif
#ifndef SCLANE_SWEEP_MINB
#define SCLANE_SWEEP_MINB 8
#endif
#ifndef SCLANE_SWEEP_PREFETCH
#define SCLANE_SWEEP_PREFETCH 0
#endif
// ablation switches for tools/stream_probe.cu (all 1 in the product build)
#ifndef SCLANE_SWEEP_DO_FLUSH
#define SCLANE_SWEEP_DO_FLUSH 1
#endif
#ifndef SCLANE_SWEEP_DO_BOUNDARY
#define SCLANE_SWEEP_DO_BOUNDARY 1
#endif
#ifndef SCLANE_SWEEP_DO_EPILOGUE
#define SCLANE_SWEEP_DO_EPILOGUE 1
#endif
constexpr int kSweepWarps = SCLANE_SWEEP_WARPS;
constexpr int kSweepThreads = kSweepWarps * 32;

struct SweepTile {      // one lane's share of a 128-cell tile: 4 consecutive cells
  int4 y;
  double2 ma, mb, ua, ub;
};
__device__ __forceinline__ SweepTile sweep_load(const int* y, const double* mu, const double* u, int i0) {
  SweepTile t;
  t.y = ld_stream_i4(y + i0);
  t.ma = ld_stream_d2(mu + i0);
  t.mb = ld_stream_d2(mu + i0 + 2);
  t.ua = __ldg(reinterpret_cast<const double2*>(u + i0));
  t.ub = __ldg(reinterpret_cast<const double2*>(u + i0 + 2));
  return t;
}

#ifndef SCLANE_SWEEP_CELL2
#define SCLANE_SWEEP_CELL2 0
#endif
// One CTA per gene, kSweepWarps warps; each warp streams a contiguous run of 128-cell tiles with the running
// bucket's moments in registers and reduces them (shuffle tree) into its shared-memory slot at bucket boundaries.
// A tile that contains a boundary feeds two register sets in one sweep (running bucket / next bucket).
template <bool WITH_Y>
__global__ void __launch_bounds__(kSweepThreads, SCLANE_SWEEP_MINB)
k_sweep_moments(const Lineage* __restrict__ lineage, const int* __restrict__ ys, const double* __restrict__ mu_all,
                const double* __restrict__ u, const GeneOutDev* __restrict__ out, double* __restrict__ mom, int n_genes) {
  constexpr int NA = WITH_Y ? 7 : 5;
  __shared__ int bstart[NBMAX + 1];
  __shared__ double slot[kSweepWarps][NBMAX][NACC];
  const int g = blockIdx.x;
  if (g >= n_genes) return;
  const GeneState& st = out[g].st;
  if (st.fwd_done || st.error) return;                 // block-uniform
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int N = lineage->N, T = lineage->T, Npad = lineage->Npad;
  const double sigma = st.sigma;
  const int* __restrict__ y = ys + (size_t)g * (size_t)Npad;
  const double* __restrict__ mu = mu_all + (size_t)g * (size_t)Npad;
  const int ntiles = (N + 127) >> 7;
  const int t0 = (int)(((long long)ntiles * warp) / kSweepWarps);
  const int t1 = (int)(((long long)ntiles * (warp + 1)) / kSweepWarps);
  // the first tile's loads are issued before the shared-memory prologue so that their latency overlaps it
  SweepTile cur;
  if (t0 < t1) cur = sweep_load(y, mu, u, (t0 << 7) + (lane << 2));
  for (int i = threadIdx.x; i <= NBMAX; i += kSweepThreads) bstart[i] = lineage->bstart[i];
  double* myslot = &slot[warp][0][0];
  for (int i = lane; i < (T + 1) * NACC; i += 32) myslot[i] = 0.0;
  __syncthreads();
  double acc[NA];
#pragma unroll
  for (int k = 0; k < NA; ++k) acc[k] = 0.0;
  auto flush = [&](int b) {                            // bucket boundary: warp-shuffle tree, lane 0 owns the slot
#if SCLANE_SWEEP_DO_FLUSH
#pragma unroll
    for (int k = 0; k < NA; ++k) {
      const double v = warp_sum(acc[k]);
      if (lane == 0) myslot[b * NACC + k] += v;
      acc[k] = 0.0;
    }
#endif
  };
  auto cell = [&](int yi, double ui, double mi) {
    const double yd = (double)yi;
    const double inv = rcp_ge1(fma(sigma, mi, 1.0));
    const double w = mi * inv, r = (yd - mi) * inv;
    const double wu = w * ui;
    acc[0] += w; acc[1] += wu; acc[2] = fma(wu, ui, acc[2]);
    acc[3] += r; acc[4] = fma(r, ui, acc[4]);
    if (WITH_Y) { acc[5] += yd; acc[6] = fma(yd, ui, acc[6]); }
  };
  if (t0 < t1) {
    int cur_b = 0;
    const int first = t0 << 7;
    while (cur_b < T && bstart[cur_b + 1] <= first) ++cur_b;
    int b_hi = bstart[cur_b + 1];
#pragma unroll 1
    for (int t = t0; t < t1; ++t) {
      const int tile_lo = t << 7;
      const int i0 = tile_lo + (lane << 2);
#if SCLANE_SWEEP_PREFETCH
      SweepTile nxt = cur;
      if (t + 1 < t1) nxt = sweep_load(y, mu, u, i0 + 128);        // software pipeline: next tile in flight
#else
      if (t > t0) cur = sweep_load(y, mu, u, i0);
#endif
      if (!SCLANE_SWEEP_DO_BOUNDARY || (tile_lo + 128 <= b_hi && tile_lo + 128 <= N)) {
        cell(cur.y.x, cur.ua.x, cur.ma.x); cell(cur.y.y, cur.ua.y, cur.ma.y);
        cell(cur.y.z, cur.ub.x, cur.mb.x); cell(cur.y.w, cur.ub.y, cur.mb.y);
      } else {
        const int yv[4] = {cur.y.x, cur.y.y, cur.y.z, cur.y.w};
        const double uv[4] = {cur.ua.x, cur.ua.y, cur.ub.x, cur.ub.y};
        const double mv[4] = {cur.ma.x, cur.ma.y, cur.mb.x, cur.mb.y};
        const int tile_end = min(tile_lo + 128, N);
#if SCLANE_SWEEP_CELL2
        // one sweep of the tile: cells below `hi` go to the running bucket, cells in [hi, hi2) to the next one
        const int hi = min(b_hi, tile_end);
        const int hi2 = cur_b < T ? min(bstart[cur_b + 2], tile_end) : hi;
        double nx[NA];
#pragma unroll
        for (int k = 0; k < NA; ++k) nx[k] = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = i0 + j;
          const double ca = (i < hi) ? 1.0 : 0.0, cn = (i >= hi && i < hi2) ? 1.0 : 0.0;
          const double yd = (double)yv[j];
          const double inv = rcp_ge1(fma(sigma, mv[j], 1.0));
          const double w = mv[j] * inv, r = (yd - mv[j]) * inv;
          const double wu = w * uv[j], wuu = wu * uv[j], ru = r * uv[j];
          acc[0] = fma(ca, w, acc[0]); acc[1] = fma(ca, wu, acc[1]); acc[2] = fma(ca, wuu, acc[2]);
          acc[3] = fma(ca, r, acc[3]); acc[4] = fma(ca, ru, acc[4]);
          nx[0] = fma(cn, w, nx[0]); nx[1] = fma(cn, wu, nx[1]); nx[2] = fma(cn, wuu, nx[2]);
          nx[3] = fma(cn, r, nx[3]); nx[4] = fma(cn, ru, nx[4]);
          if (WITH_Y) {
            const double yu = yd * uv[j];
            acc[5] = fma(ca, yd, acc[5]); acc[6] = fma(ca, yu, acc[6]);
            nx[5] = fma(cn, yd, nx[5]); nx[6] = fma(cn, yu, nx[6]);
          }
        }
        if (hi < tile_end) {
          flush(cur_b);
#pragma unroll
          for (int k = 0; k < NA; ++k) acc[k] = nx[k];
          ++cur_b;
          b_hi = bstart[cur_b + 1];
          // buckets narrower than the rest of the tile (rare): masked sweeps, one bucket at a time
#pragma unroll 1
          while (b_hi < tile_end) {
            flush(cur_b);
            const int lo3 = b_hi;
            ++cur_b;
            b_hi = bstart[cur_b + 1];
            const int hi3 = min(b_hi, tile_end);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = i0 + j;
              if (i >= lo3 && i < hi3) cell(yv[j], uv[j], mv[j]);
            }
          }
        }
#else
#pragma unroll 1
        for (;;) {
          const int lo = bstart[cur_b];
          const int hi = min(b_hi, tile_end);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = i0 + j;
            if (i >= lo && i < hi) cell(yv[j], uv[j], mv[j]);
          }
          if (hi >= tile_end) break;
          flush(cur_b);
          ++cur_b;
          b_hi = bstart[cur_b + 1];
        }
#endif
      }
#if SCLANE_SWEEP_PREFETCH
      cur = nxt;
#endif
    }
#if SCLANE_SWEEP_DO_FLUSH
    flush(cur_b);
#else
    { double s = 0.0; for (int k = 0; k < NA; ++k) s += acc[k]; if (s == 123.456) myslot[0] = s; }
#endif
  }
#if SCLANE_SWEEP_DO_EPILOGUE
  __syncthreads();
  double* dst = mom + (size_t)g * (size_t)(NBMAX * NACC);
  for (int i = threadIdx.x; i < (T + 1) * NACC; i += kSweepThreads) {
    if ((i % NACC) < NA) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kSweepWarps; ++w) s += (&slot[w][0][0])[i];
      dst[i] = s;
    }
  }
#endif
}

// ---------------------------------------------------------------------------------------------
// K2b: one WARP per gene.  From the moments: A = (B'WB)^-1, the 2T scores (lane = knot), the selection tree of
// marge2.R:368-587, the stat_out check and the forward bookkeeping.
// ---------------------------------------------------------------------------------------------
struct WarpExec {
  static constexpr bool kTabs = false;      // term tables attached to Work (sclane_core.cuh term_coef)
  __device__ __forceinline__ PassTabs* tabs() const { return nullptr; }     // the select kernel runs no pass
  __device__ __forceinline__ bool leader() const { return (threadIdx.x & 31) == 0; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  template <class F>
  __device__ __forceinline__ void par_for(int n, F f) const {
    for (int i = threadIdx.x & 31; i < n; i += 32) f(i);
  }
};

constexpr int kSelectWarps = 4;

__global__ void __launch_bounds__(kSelectWarps * 32)
k_sweep_select(StageArgs A) {
  __shared__ Lineage Ls;
  __shared__ Work Wall[kSelectWarps];
  __shared__ GeneState stall[kSelectWarps];
  {
    const int* src = reinterpret_cast<const int*>(A.lineage);
    int* dst = reinterpret_cast<int*>(&Ls);
    for (int i = threadIdx.x; i < (int)(sizeof(Lineage) / sizeof(int)); i += kSelectWarps * 32) dst[i] = src[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = blockIdx.x * kSelectWarps + warp;
  if (g >= A.n_genes) return;                                 // warp-uniform; no block-wide barrier below
  Work& W = Wall[warp];
  GeneState& st = stall[warp];
  {
    const int* s2 = reinterpret_cast<const int*>(&A.out[g].st);
    int* d2 = reinterpret_cast<int*>(&st);
    for (int i = lane; i < (int)(sizeof(GeneState) / sizeof(int)); i += 32) d2[i] = s2[i];
  }
  __syncwarp();
  if (st.fwd_done || st.error) return;                        // warp-uniform
  const double* m = A.mom + (size_t)g * (size_t)(NBMAX * NACC);
  double* wacc = &W.acc[0][0];
  for (int i = lane; i < (Ls.T + 1) * NACC; i += 32) wacc[i] = m[i];
  if (lane == 0) { W.passes = 0; W.flag = 0; work_attach_tabs(W, nullptr); }      // no term tables: term_coef() on the fly
  __syncwarp();
  WarpExec ex;
  const GeneStats gs = A.stats[g];
  gene_sweep_select(ex, Ls, gs, st, W, A.M, A.tols_score);
  __syncwarp();
  if (A.scores != nullptr) {
    double* o = A.scores + (size_t)g * 2 * Ls.T;
    for (int i = lane; i < Ls.T; i += 32) { o[i] = W.sw_one[i]; o[Ls.T + i] = W.sw_both[i]; }
  }
  if (lane == 0) st.passes += 1;
  __syncwarp();
  {
    int* d2 = reinterpret_cast<int*>(&A.out[g].st);
    const int* s2 = reinterpret_cast<const int*>(&st);
    for (int i = lane; i < (int)(sizeof(GeneState) / sizeof(int)); i += 32) d2[i] = s2[i];
  }
}

// ---------------------------------------------------------------------------------------------
// K0b: apply the lineage's sort permutation to a batch of genes and gather the per-gene constants
// (replaces the per-gene `expr.mat[lineage_cells, i]` column extraction of testDynamic.R:190).
// raw: [B][ld] gene-major counts in the caller's cell order; ys: [B][Npad] sorted, zero padded.
// ---------------------------------------------------------------------------------------------
template <class RAW>
__global__ void __launch_bounds__(kThreads) k_gather_stats(const RAW* __restrict__ raw, long long ld, const int* __restrict__ perm,
                                                           int N, int Npad, int* __restrict__ ys, GeneStats* __restrict__ stats,
                                                           GeneOutDev* __restrict__ out, int n_genes) {
  const int g = blockIdx.x;
  if (g >= n_genes) return;
  const RAW* src = raw + (size_t)g * (size_t)ld;
  int* dst = ys + (size_t)g * (size_t)Npad;
  double s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0;
  int ymax = 0;
  for (int i = threadIdx.x; i < Npad; i += kThreads) {
    int v = 0;
    if (i < N) v = (int)__ldg(src + __ldg(perm + i));
    if (v < 0) { v = 0; ymax = 0x7fffffff; }          // negative count: flagged through ymax, stored as 0
    dst[i] = v;
    if (v > 0) {
      const double d = (double)v;
      s1 += d; s2 += d * d;
      if (v > 1) { s3 += d * log(d); s4 += lgamma(d + 1.0); }
      ymax = max(ymax, v);
    }
  }
  __shared__ double red[kWarps][4];
  __shared__ int redm[kWarps];
  s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3); s4 = warp_sum(s4);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[warp][0] = s1; red[warp][1] = s2; red[warp][2] = s3; red[warp][3] = s4; redm[warp] = ymax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    GeneStats gs;
    gs.sum_y = gs.sum_y2 = gs.sum_ylogy = gs.sum_lgamma_y1 = 0.0;
    gs.ymax = 0; gs.pad = 0;
    for (int w = 0; w < kWarps; ++w) {
      gs.sum_y += red[w][0]; gs.sum_y2 += red[w][1]; gs.sum_ylogy += red[w][2]; gs.sum_lgamma_y1 += red[w][3];
      gs.ymax = max(gs.ymax, redm[w]);
    }
    if (gs.ymax == 0x7fffffff) { gs.pad = 1; gs.ymax = 0; }   // pad = 1: the gene has a negative count (SCLANE_NOTE_BAD_COUNTS)
    stats[g] = gs;
    if (!out) return;                       // GEE mode initialises its own per-gene state (k_gee)
    GeneState st;
    int* z = reinterpret_cast<int*>(&st);
    for (int i = 0; i < (int)(sizeof(GeneState) / sizeof(int)); ++i) z[i] = 0;
    st.fwd.n = 1;
    st.fwd.kind[0] = SCLANE_TERM_INTERCEPT;
    st.m_count = 1;
    if (gs.sum_y == 0.0) st.error |= SCLANE_NOTE_ALL_ZERO;
    if (gs.pad) st.error |= SCLANE_NOTE_BAD_COUNTS;
    out[g].st = st;
    out[g].alt.ok = 0;
    out[g].null.ok = 0;
  }
}

// Launch order of the final stage.  The cost of glm.nb varies ~10x between genes and the few near-Poisson genes (theta runs
// away: up to 25 alternations of IRLS + theta.ml) take 5-15x the median; when such a gene starts late its CTA runs alone at
// the end of the launch (measured: SMs 45 % active in a 2,048-gene launch of k_stage<FINAL>).  The forward pass already
// knows them -- sigma of the last forward null fit is ~0 -- so the final stage visits those first and the rest by descending
// forward score (rank sort, ties by index: a deterministic permutation; which CTA fits which gene does not change any result).
__global__ void k_order_final(const GeneOutDev* __restrict__ out, int n, double* __restrict__ keys, int* __restrict__ order, int phase) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (phase == 2) {
    // backward pass: one NB refit per column of the forward model -> widest models first, then by forward score
    const GeneState& st = out[i].st;
    double sc = 0.0;
    for (int t = 0; t < st.fwd_iters && t < 4; ++t) if (st.fwd_score[t] == st.fwd_score[t] && st.fwd_score[t] > sc) sc = st.fwd_score[t];
    keys[i] = st.error ? 1e300 : -((double)st.fwd.n * 1e12 + (sc < 1e11 ? sc : 1e11));
    return;
  }
  if (phase == 3) {
    // final fits WITH an offset: the forward fits ran without it, so a gene whose forward sigma is small is one whose
    // dispersion is mostly depth variation -- with the offset its theta runs away and glm.nb goes to the alternation limit
    // (bench generator: Spearman -0.71 between forward sigma and the final stage's passes; the score only 0.43)
    // The offset quadrature has fitted the null model by now: a null fit that ran to the alternation limit (theta runs away
    // once the depth variation is in the offset) predicts the same for the alternative -- those genes cost 5-10x the median.
    const GeneState& st = out[i].st;
    const double sg = st.sigma < 1e5 ? st.sigma : 1e5;
    keys[i] = (st.error || st.fin.n <= 1) ? 1e300 : (out[i].null.ok ? -1e6 * (double)out[i].null.alternations : 0.0) + sg;
    return;
  }
  if (phase == 0) {
    const GeneState& st = out[i].st;
    double k;
    if (st.error) k = 1e300;                      // failed genes return at once: last
    else if (st.sigma < 1e-6) k = -1e300;         // Poisson regime / near-Poisson: theta runs away, up to 25 alternations: first
    else {
      // then by descending forward score: strongly dynamic genes need more alternations (Spearman 0.8 with the final
      // stage's iteration count on the bench generator; sigma alone 0.2)
      double sc = 0.0;
      for (int t = 0; t < st.fwd_iters && t < 4; ++t) if (st.fwd_score[t] == st.fwd_score[t] && st.fwd_score[t] > sc) sc = st.fwd_score[t];
      k = -sc;
    }
    keys[i] = k;
    return;
  }
  // NaN keys would break the strict weak order (and with it the permutation): they sort with the failed genes
  double k = keys[i];
  if (!(k == k)) k = 1e300;
  int rank = 0;
  for (int j = 0; j < n; ++j) {
    double kj = keys[j];
    if (!(kj == kj)) kj = 1e300;
    rank += (kj < k || (kj == k && j < i)) ? 1 : 0;
  }
  order[rank] = i;
}

// gather of a double matrix by the same permutation (sclane_score_glm: mu in caller order -> sorted)
__global__ void k_gather_f64(const double* __restrict__ raw, long long ld, const int* __restrict__ perm, int N, int Npad,
                             double* __restrict__ dst, int n_genes) {
  const int g = blockIdx.x;
  if (g >= n_genes) return;
  for (int i = threadIdx.x; i < Npad; i += blockDim.x)
    dst[(size_t)g * Npad + i] = i < N ? raw[(size_t)g * ld + perm[i]] : 0.0;
}

}  // namespace sclane
