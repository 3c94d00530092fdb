// sclane_k0_device.cuh -- K0 (the per-lineage plan) on the device.
//
// What build_lineage_plan() (sclane_host.hpp) does on the host -- drop the cells outside the lineage, X = round(x, 4), stable
// sort by X, candidate knots (marge2.R:150-166: min_span.R:24-34, max_span.R:20-25, the 5 % / 95 % quantiles of the
// unrounded pseudotime, thinning to n.knot.max by seq()), bucket table, bucket-local coordinates, unweighted bucket moments
// and the table of distinct abscissae -- as a short chain of kernels, so that a call whose counts are already on the device
// does not wait ~10 ms (200,000 cells) for host work before its first fit kernel starts, and so that eight ranks sharing one
// host do not build the same plan on the same cores.  The host version stays as the checker: every table produced here is
// BIT-IDENTICAL to the host's (tests/test_parity_r02.py::test_device_plan_bit_identical), which is why
//   * sums that the host takes serially are taken serially here (one CTA per bucket, thread 0 adds from shared memory),
//   * every a * b + c of the host source is spelled __dadd_rn(__dmul_rn(a, b), c): nvcc would contract it to an FMA, g++ on
//     x86-64 does not.
// Used for approx.knot = TRUE without size factors (log(1 / sf) from the device's log is not libm's bit for bit) and a key
// range that a 24-bit radix sort covers; everything else keeps the host plan.
#pragma once
#include "sclane_core.cuh"

namespace sclane {

constexpr int kK0Threads = 1024;
constexpr int kK0Warps = kK0Threads / 32;
constexpr int kK0KnotCap = TMAX;

struct K0Quant {               // stats::quantile(type = 7) bookkeeping done on the host (depends on N only)
  int lo[2];                   // 1-based floor(index) for p = 0.05 / 0.95
  int has_hi[2];               // ceil(index) > floor(index)
  int above[2];                // index > floor(index)
  double h[2], one_minus_h[2];
};

struct K0Result {              // what the host needs back (one small D2H)
  Lineage L;
  int D;                       // distinct abscissae
  int n_red;                   // candidates after min_span / max_span / quantile filter (before thinning)
  int pbstart[NBMAX + 1];
  double q05, q95;
};

struct K0Args {
  const double* pt;            // [n_cells] pseudotime column on the device, NaN = not in the lineage
  int n_cells, N, Npad;
  long long key0;              // llround(1e4 * min X)
  int step;                    // minspan + 1
  int maxspan;
  int n_knot_max;
  K0Quant Q;
  // work arrays
  int* blk;                    // [n_blocks + 1] per-block counts / offsets (compaction, points)
  int* cells;                  // [N] compact index -> cell
  double* xc;                  // [N] unrounded pseudotime, compact order
  double* Xc;                  // [N] rounded pseudotime, compact order
  unsigned* key[2];            // [N] radix ping-pong
  int* idx[2];                 // [N]
  int* hist;                   // [256 * n_tiles]
  double* Xs;                  // [N] sorted rounded pseudotime
  double* xs;                  // [N] unrounded pseudotime in the same (X-sorted) order
  unsigned* key_s;             // -> key[...] holding the sorted keys
  int* idx_s;
  double* qstat;               // [4] order statistics
  int* scal;                   // [4]: D
  // outputs
  int* perm;                   // [N]
  double* u;                   // [Npad]
  int* pstart;                 // [N + 1] (D + 1 used)
  double* pn;                  // [N]
  double* pu;                  // [N]
  unsigned char* pb;           // [N]
  int* pbstart;                // [NBMAX + 1]
  Lineage* lin;
  K0Result* res;
};

// exclusive prefix of one int per thread over a 1024-thread block; returns the block total through *total (all threads)
__device__ __forceinline__ int k0_block_scan(int v, int* warp_tot /*[kK0Warps + 1]*/, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
    warp_tot[lane] = winc - w;
    if (lane == 31) warp_tot[kK0Warps] = winc;
  }
  __syncthreads();
  const int r = warp_tot[warp] + inc - v;
  *total = warp_tot[kK0Warps];
  __syncthreads();
  return r;
}

// in-place exclusive scan of a[0..n) by ONE block; a[n] = total
__global__ void __launch_bounds__(kK0Threads) k0_scan(int* a, int n) {
  __shared__ int wt[kK0Warps + 1];
  const int chunk = (n + kK0Threads - 1) / kK0Threads;
  const int i0 = min(n, (int)threadIdx.x * chunk), i1 = min(n, i0 + chunk);
  int s = 0;
  for (int i = i0; i < i1; ++i) s += a[i];
  int total;
  int run = k0_block_scan(s, wt, &total);
  for (int i = i0; i < i1; ++i) { const int v = a[i]; a[i] = run; run += v; }
  if (threadIdx.x == 0) a[n] = total;
}

// cells of the lineage per block
__global__ void __launch_bounds__(kK0Threads) k0_valid_count(K0Args A) {
  __shared__ int wt[kK0Warps + 1];
  const int i = blockIdx.x * kK0Threads + threadIdx.x;
  const int v = (i < A.n_cells && !isnan(A.pt[i])) ? 1 : 0;
  int total;
  k0_block_scan(v, wt, &total);
  if (threadIdx.x == 0) A.blk[blockIdx.x] = total;
}

// stable compaction + rounding + integer key
__global__ void __launch_bounds__(kK0Threads) k0_compact(K0Args A) {
  __shared__ int wt[kK0Warps + 1];
  const int i = blockIdx.x * kK0Threads + threadIdx.x;
  double x = 0.0;
  int v = 0;
  if (i < A.n_cells) { x = A.pt[i]; v = isnan(x) ? 0 : 1; }
  int total;
  const int r = k0_block_scan(v, wt, &total);
  if (v) {
    const int c = A.blk[blockIdx.x] + r;
    const double X = r_round(x, 1e4);
    A.cells[c] = i;
    A.xc[c] = x;
    A.Xc[c] = X;
    A.key[0][c] = (unsigned)(llround(__dmul_rn(X, 1e4)) - A.key0);
    A.idx[0][c] = c;
  }
}

// ---- stable LSD radix sort of (key, idx), 8 bits per pass, one 1024-element tile per block -------------------------
__global__ void __launch_bounds__(kK0Threads) k0_radix_hist(const unsigned* __restrict__ key, int N, int shift, int n_tiles, int* __restrict__ hist) {
  __shared__ int h[256];
  if (threadIdx.x < 256) h[threadIdx.x] = 0;
  __syncthreads();
  const int i = blockIdx.x * kK0Threads + threadIdx.x;
  if (i < N) atomicAdd(&h[(key[i] >> shift) & 255u], 1);
  __syncthreads();
  if (threadIdx.x < 256) hist[threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(kK0Threads) k0_radix_scatter(const unsigned* __restrict__ key_in, const int* __restrict__ idx_in, unsigned* __restrict__ key_out,
                                                               int* __restrict__ idx_out, int N, int shift, int n_tiles, const int* __restrict__ base) {
  __shared__ int wcnt[kK0Warps][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = threadIdx.x; j < kK0Warps * 256; j += kK0Threads) (&wcnt[0][0])[j] = 0;
  __syncthreads();
  const int i = blockIdx.x * kK0Threads + threadIdx.x;
  const bool valid = i < N;
  unsigned k = 0;
  int id = 0;
  if (valid) { k = key_in[i]; id = idx_in[i]; }
  const unsigned d = valid ? ((k >> shift) & 255u) : 256u;
  const unsigned peers = __match_any_sync(0xffffffffu, d);
  const int rank = __popc(peers & ((1u << lane) - 1u));
  if (valid && rank == 0) wcnt[warp][d] = __popc(peers);
  __syncthreads();
  if (threadIdx.x < 256) {
    int run = 0;
#pragma unroll 4
    for (int w = 0; w < kK0Warps; ++w) { const int c = wcnt[w][threadIdx.x]; wcnt[w][threadIdx.x] = run; run += c; }
  }
  __syncthreads();
  if (valid) {
    const int pos = base[d * n_tiles + blockIdx.x] + wcnt[warp][d] + rank;
    key_out[pos] = k;
    idx_out[pos] = id;
  }
}

// sorted tables + "first cell of a distinct abscissa" flags counted per block
__global__ void __launch_bounds__(kK0Threads) k0_sorted(K0Args A) {
  __shared__ int wt[kK0Warps + 1];
  const int i = blockIdx.x * kK0Threads + threadIdx.x;
  int f = 0;
  if (i < A.N) {
    const int o = A.idx_s[i];
    A.perm[i] = A.cells[o];
    A.Xs[i] = A.Xc[o];
    A.xs[i] = A.xc[o];
    f = (i == 0 || A.key_s[i] != A.key_s[i - 1]) ? 1 : 0;
  }
  int total;
  k0_block_scan(f, wt, &total);
  if (threadIdx.x == 0) A.blk[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kK0Threads) k0_points(K0Args A, int n_blocks) {
  __shared__ int wt[kK0Warps + 1];
  const int i = blockIdx.x * kK0Threads + threadIdx.x;
  const int f = (i < A.N && (i == 0 || A.key_s[i] != A.key_s[i - 1])) ? 1 : 0;
  int total;
  const int r = k0_block_scan(f, wt, &total);
  if (f) A.pstart[A.blk[blockIdx.x] + r] = i;
  if (blockIdx.x == 0 && threadIdx.x == 0) { const int D = A.blk[n_blocks]; A.scal[0] = D; A.pstart[D] = A.N; }
}

// ---- the four order statistics of the UNROUNDED pseudotime behind quantile(x, c(0.05, 0.95)) ------------------------
// Rounding is monotone, so the k-th smallest x lies among the cells that share the rounded value of the k-th cell of the
// X-sorted order: a radix select (8 passes over a 64-bit order-preserving key) inside that group, one block per statistic.
__device__ __forceinline__ unsigned long long k0_ordered(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double k0_unordered(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(kK0Threads) k0_quantile_stats(K0Args A) {
  __shared__ int h[256];
  __shared__ unsigned long long prefix_s;
  __shared__ int r_s, g0_s, g1_s;
  const int which = blockIdx.x >> 1, second = blockIdx.x & 1;       // p = 0.05 / 0.95; floor / ceil neighbour
  if (second && !A.Q.has_hi[which]) { if (threadIdx.x == 0) A.qstat[blockIdx.x] = 0.0; return; }
  const int pos = A.Q.lo[which] - 1 + second;                        // 0-based rank in ascending x
  if (threadIdx.x == 0) {
    const unsigned kk = A.key_s[pos];
    int x = 0, y = A.N;
    while (x < y) { const int m = (x + y) >> 1; if (A.key_s[m] < kk) x = m + 1; else y = m; }
    g0_s = x;
    y = A.N;
    while (x < y) { const int m = (x + y) >> 1; if (A.key_s[m] <= kk) x = m + 1; else y = m; }
    g1_s = x;
    r_s = pos - g0_s;
    prefix_s = 0ull;
  }
  __syncthreads();
  const int g0 = g0_s, g1 = g1_s;
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    if (threadIdx.x < 256) h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long prefix = prefix_s;
    const unsigned long long mask = pass == 0 ? 0ull : (~0ull << (shift + 8));
    for (int i = g0 + (int)threadIdx.x; i < g1; i += kK0Threads) {
      const unsigned long long k = k0_ordered(A.xs[i]);
      if ((k & mask) == prefix) atomicAdd(&h[(int)((k >> shift) & 255ull)], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int r = r_s, b = 0;
      while (b < 255 && r >= h[b]) { r -= h[b]; ++b; }
      r_s = r;
      prefix_s = prefix | ((unsigned long long)b << shift);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) A.qstat[blockIdx.x] = k0_unordered(prefix_s);
}

// ---- candidate knots, bucket table ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kK0Threads) k0_knots(K0Args A) {
  __shared__ int cnt_s, first_s, last_s, n_small_s;
  __shared__ double small_s[kK0KnotCap];
  __shared__ double q_s[2];
  __shared__ Lineage L;
  const int D = A.scal[0];
  if (threadIdx.x == 0) {
    cnt_s = 0; first_s = D; last_s = -1; n_small_s = 0;
    for (int w = 0; w < 2; ++w) {                                     // quantile7 (sclane_host.hpp)
      double q = A.qstat[2 * w];
      if (A.Q.has_hi[w]) {
        const double xhi = A.qstat[2 * w + 1];
        if (A.Q.above[w] && xhi != q) q = __dadd_rn(__dmul_rn(A.Q.one_minus_h[w], q), __dmul_rn(A.Q.h[w], xhi));
      }
      q_s[w] = q;
    }
    int* z = reinterpret_cast<int*>(&L);
    for (int i = 0; i < (int)(sizeof(Lineage) / sizeof(int)); ++i) z[i] = 0;
  }
  __syncthreads();
  const double q05 = q_s[0], q95 = q_s[1];
  const int step = A.step, maxspan = A.maxspan;
  const bool r2_all = D <= 2 * maxspan;                               // max_span() returns its input when nothing is left
  auto is_red = [&](int d) -> bool {
    const int c0 = A.pstart[d], c1 = A.pstart[d + 1];
    const long long m = ((long long)c0 + step - 1) / step * step;     // first multiple of step >= c0 (min_span keeps cells 0, step, 2 step, ...)
    if (m >= c1) return false;
    if (!r2_all && (d < maxspan || d > D - maxspan - 1)) return false;
    const double X = A.Xs[c0];
    return X > q05 && X < q95;
  };
  int my_cnt = 0, my_first = D, my_last = -1;
  for (int d = threadIdx.x; d < D; d += kK0Threads)
    if (is_red(d)) { ++my_cnt; my_first = min(my_first, d); my_last = max(my_last, d); }
  if (my_cnt) { atomicAdd(&cnt_s, my_cnt); atomicMin(&first_s, my_first); atomicMax(&last_s, my_last); }
  __syncthreads();
  const int n_red = cnt_s;
  if (n_red > 0 && n_red <= A.n_knot_max && n_red <= kK0KnotCap) {
    for (int d = threadIdx.x; d < D; d += kK0Threads)
      if (is_red(d)) small_s[atomicAdd(&n_small_s, 1)] = A.Xs[A.pstart[d]];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int T = 0;
    if (n_red > A.n_knot_max) {
      const double lo = A.Xs[A.pstart[first_s]], hi = A.Xs[A.pstart[last_s]];
      const int nk = A.n_knot_max;
      const double by = (hi - lo) / (double)(nk - 1);                 // seq.default(from, to, length.out)
      for (int k = 0; k < nk; ++k) {
        const double v = (k == 0) ? lo : (k == nk - 1 ? hi : __dadd_rn(lo, __dmul_rn((double)k, by)));
        L.knots[k] = r_round(v, 1e4);
      }
      T = nk;
    } else if (n_red > 0) {
      T = n_red;
      for (int k = 0; k < T; ++k) L.knots[k] = small_s[k];
    }
    for (int a = 1; a < T; ++a) {                                     // ascending (the append order above is arbitrary)
      const double v = L.knots[a];
      int b = a - 1;
      while (b >= 0 && L.knots[b] > v) { L.knots[b + 1] = L.knots[b]; --b; }
      L.knots[b + 1] = v;
    }
    L.N = A.N; L.Npad = A.Npad; L.T = T;
    L.ref[0] = T > 0 ? L.knots[0] : A.Xs[0];
    for (int b = 1; b <= T; ++b) L.ref[b] = L.knots[b - 1];
  }
  __syncthreads();
  const int T = L.T;
  if ((int)threadIdx.x <= NBMAX) {
    const int b = threadIdx.x;
    int v = A.N;
    if (b == 0) v = 0;
    else if (b <= T) {
      const double t = L.knots[b - 1];
      int x = 0, y = A.N;
      while (x < y) { const int m = (x + y) >> 1; if (A.Xs[m] < t) x = m + 1; else y = m; }
      v = x;
    }
    L.bstart[b] = v;
  }
  __syncthreads();
  // pbstart[b] = first point whose first cell is >= bstart[b]
  if ((int)threadIdx.x <= NBMAX) {
    const int b = threadIdx.x;
    int v = D;
    if (b <= T) {
      const int c = L.bstart[b];
      int x = 0, y = D;
      while (x < y) { const int m = (x + y) >> 1; if (A.pstart[m] < c) x = m + 1; else y = m; }
      v = x;
    }
    A.pbstart[b] = v;
    A.res->pbstart[b] = v;
  }
  __syncthreads();
  {
    const int* src = reinterpret_cast<const int*>(&L);
    int* dst = reinterpret_cast<int*>(A.lin);
    for (int i = threadIdx.x; i < (int)(sizeof(Lineage) / sizeof(int)); i += kK0Threads) dst[i] = src[i];
  }
  if (threadIdx.x == 0) { A.res->D = D; A.res->n_red = n_red; A.res->q05 = q05; A.res->q95 = q95; }
}

// unweighted bucket moments, summed in cell order exactly as the host does: one block per bucket stages the bucket's
// coordinates through shared memory, thread 0 adds them up
__global__ void __launch_bounds__(256) k0_bucket_moments(K0Args A) {
  __shared__ double buf[2048];
  const int b = blockIdx.x;
  const int T = A.lin->T;
  if (b > T) return;
  const int c0 = A.lin->bstart[b], c1 = A.lin->bstart[b + 1];
  const double ref = A.lin->ref[b];
  double u0 = 0.0, u1 = 0.0, u2 = 0.0;
  for (int base = c0; base < c1; base += 2048) {
    const int n = min(2048, c1 - base);
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += 256) buf[j] = A.Xs[base + j] - ref;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll 8
      for (int j = 0; j < n; ++j) { const double u = buf[j]; u0 += 1.0; u1 = __dadd_rn(u1, u); u2 = __dadd_rn(u2, __dmul_rn(u, u)); }
    }
  }
  if (threadIdx.x == 0) { A.lin->U0[b] = u0; A.lin->U1[b] = u1; A.lin->U2[b] = u2; }
}

// per-cell coordinate, per-point tables, and the Lineage copy that goes back to the host
__global__ void __launch_bounds__(kK0Threads) k0_fill(K0Args A) {
  __shared__ int bs[NBMAX + 1];
  __shared__ double ref[NBMAX];
  __shared__ int T_s;
  if ((int)threadIdx.x <= NBMAX) bs[threadIdx.x] = A.lin->bstart[threadIdx.x];
  if ((int)threadIdx.x < NBMAX) ref[threadIdx.x] = A.lin->ref[threadIdx.x];
  if (threadIdx.x == 0) T_s = A.lin->T;
  __syncthreads();
  const int T = T_s, D = A.scal[0];
  auto bucket_of = [&](int cell) {                                    // last b <= T with bstart[b] <= cell
    int x = 0, y = T + 1;
    while (x < y) { const int m = (x + y) >> 1; if (bs[m] <= cell) x = m + 1; else y = m; }
    return x - 1;
  };
  const int i = blockIdx.x * kK0Threads + threadIdx.x;
  if (i < A.Npad) A.u[i] = i < A.N ? A.Xs[i] - ref[bucket_of(i)] : 0.0;
  if (i < D) {
    const int c0 = A.pstart[i], c1 = A.pstart[i + 1];
    const int b = bucket_of(c0);
    A.pu[i] = A.Xs[c0] - ref[b];
    A.pn[i] = (double)(c1 - c0);
    A.pb[i] = (unsigned char)b;
  }
  if (blockIdx.x == 0) {
    const int* src = reinterpret_cast<const int*>(A.lin);
    int* dst = reinterpret_cast<int*>(&A.res->L);
    for (int j = threadIdx.x; j < (int)(sizeof(Lineage) / sizeof(int)); j += kK0Threads) dst[j] = src[j];
  }
}

}  // namespace sclane
