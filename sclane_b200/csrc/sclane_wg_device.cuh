// sclane_wg_device.cuh -- warp-per-gene kernels of the X-compressed GLM path (default route of every fit without an
// offset) and of the offset quadrature (intercept-only fits with an offset).
//
// Why one gene per WARP: ncu on the CTA-per-gene kernels (profiles/r02_ncu_comp_stage_cta_per_gene.txt) showed them latency
// bound, not throughput bound -- 0.45 / 0.63 / 0.93 eligible warps per scheduler cycle, barrier stalls 9.1 / 6.4 / 2.9 per
// issued instruction (forward / backward / final): a compressed pass is only ~D/256 = 40 point evaluations per thread
// between two __syncthreads, the small algebra between passes runs on one thread while 255 wait, and only 3 CTAs fit per
// SM.  Here a gene belongs to ONE warp: no CTA barrier exists (only __syncwarp), the serial algebra of one gene overlaps
// the passes of the other genes resident on the SM, and 24+ independent genes are in flight per SM instead of 3.
//   * persistent grid: every warp pulls the next gene of the launch order from a device-wide atomic counter
//     (SURVEY.md 8e work queue) -- no waves, no tail of idle CTAs, any batch size;
//   * Work without the term tables (term_coef() on the fly): ~6 KB of shared memory per gene;
//   * flat buckets (constant linear predictor) are evaluated in parallel, lane = bucket; sloped buckets lane-strided over
//     their points with two independent points per lane in flight; shuffle-tree reductions in a fixed order, so every
//     sum is bit-reproducible whatever warp / CTA / GPU fits the gene.
#pragma once
#include "sclane_comp_device.cuh"

namespace sclane {

#ifndef SCLANE_WG_WARPS
#define SCLANE_WG_WARPS 4
#endif
#ifndef SCLANE_WG_MINB
#define SCLANE_WG_MINB 6
#endif
constexpr int kWgWarps = SCLANE_WG_WARPS;       // genes in flight per CTA
constexpr int kWgThreads = kWgWarps * 32;

// per-gene node weights of the offset quadrature (k_off_aggregate -> WgExec in offset mode)
struct OffGeneDev {
  double a0tot, b0tot;          // glm.fit start: sum w0(y_i), sum [w0 z0](y_i) - sum w0(y_i) o_i
};

struct WgArgs {
  CompArgs C;
  int* counter;                 // work queue: next position of the launch order
  // offset quadrature (STAGE_OFFNULL)
  const OffsetNodes* on;        // device copy
  const double* off_s;          // [B][kOffKMax * kOffJ] (only K * kOffJ used)
  const double* off_q;
  const OffGeneDev* offg;       // [B]
};

struct WgGeneShared {
  Work W;
  GeneState st;
  GeneStats gs;
};

struct WgExec {
  static constexpr bool kTabs = false;      // term tables attached to Work (sclane_core.cuh term_coef)
  Work* W;
  const Lineage* L;
  const CompGeneDev* cg;            // global memory (read-only here)
  const GeneStats* gs;
  const double* __restrict__ pn;
  const double* __restrict__ pu;
  const int* pbstart;               // shared
  const double* __restrict__ s1;
  const double* __restrict__ q;
  const int* __restrict__ tail;
  const int* __restrict__ big;
  const FastTabs* ft;
  // offset quadrature mode (on != nullptr): passes run over the nodes instead of the points
  const OffsetNodes* on;
  const double* __restrict__ ons;
  const double* __restrict__ onq;
  const OffGeneDev* offg;
  // bucket nodes (sclane_comp.cuh) or nullptr
  const BucketNodes* bn = nullptr;
  const double* __restrict__ bs = nullptr;
  const double* __restrict__ bq = nullptr;
  const double* bnuw = nullptr;     // [NBMAX] shared-memory copy of bn->uw (-1 = bucket without nodes)

  __device__ __forceinline__ bool leader() const { return (threadIdx.x & 31) == 0; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ PassTabs* tabs() const { return nullptr; }
  template <class F>
  __device__ __forceinline__ void par_for(int n, F f) const {
    for (int i = threadIdx.x & 31; i < n; i += 32) f(i);
  }

  template <int MODE>
  __device__ __forceinline__ void hist_terms(double* sc, double sigma, double theta, bool want_ll) const {
    const int lane = threadIdx.x & 31;
    const int ymax = cg->hdr.ymax;
    const int hmax = ymax < kHistCap ? ymax : kHistCap - 1;
    for (int k = lane; k < hmax; k += 32) hist_contrib<MODE>(k, (double)tail[k], (double)tail[k + 1], sigma, theta, want_ll, sc);
    const int n_big = cg->hdr.n_big;
    for (int e = lane; e < n_big; e += 32) big_contrib<MODE>(big[e], sigma, theta, want_ll, sc);
  }

  template <int MODE>
  __device__ __noinline__ void pass() {
    if (MODE == PASS_EMIT) return;                 // the compressed sweep recomputes mu from the eta parameters
    constexpr int NA = CompTraits<MODE>::NA;
    constexpr int NS = CompTraits<MODE>::NS;
    const int lane = threadIdx.x & 31;
    const int T = L->T;
    const double sigma = W->sigma, theta = W->theta;
    const bool want_ll = W->want_ll != 0;
    __syncwarp();
    if (MODE == PASS_IRLS_START) {
      if (on) {
        for (int idx = lane; idx < (T + 1) * 5; idx += 32) { const int b = idx / 5, k = idx - b * 5; W->acc[b][k] = 0.0; }
        __syncwarp();
        if (lane == 0) { W->acc[0][0] = offg->a0tot; W->acc[0][3] = offg->b0tot; }
      } else {
        for (int idx = lane; idx < (T + 1) * 5; idx += 32) { const int b = idx / 5, k = idx - b * 5; W->acc[b][k] = cg->ST[b][k]; }
      }
      if (lane == 0) {
        W->sc[0] = -gs->sum_ylogy + (double)cg->hdr.n_zero * theta * log1p(sigma / 6.0);
        W->sc[1] = (double)cg->hdr.n_zero;
        W->sc[2] = 0.0;
      }
      __syncwarp();
      return;
    }
    double sc[NSC];
#pragma unroll
    for (int k = 0; k < NSC; ++k) sc[k] = 0.0;
    if (on) {
      // offset quadrature: K * J nodes, lane-strided; everything lands in bucket 0 (the only column is the intercept)
      if (MODE == PASS_IRLS || MODE == PASS_THETA) {
        double acc[NACC];
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
        const double a = W->a[0];
        const int n_nodes = on->K * kOffJ;
        for (int e = lane; e < n_nodes; e += 32)
          offnode_contrib<MODE>(on->node_n[e], ons[e], onq[e], on->node_o[e], a, sigma, theta, want_ll, ft, acc, sc);
        if (NA > 0) {
          const double t0 = warp_sum(acc[0]), t3 = warp_sum(acc[3]);
          if (lane == 0) { W->acc[0][0] = t0; W->acc[0][3] = t3; }
        }
      }
    } else {
      // flat buckets: one closed-form evaluation each, lane = bucket
      for (int b = lane; b <= T; b += 32) {
        if (W->c[b] == 0.0) {
          double fa[NACC];
#pragma unroll
          for (int k = 0; k < NACC; ++k) fa[k] = 0.0;
          flat_bucket_contrib<MODE>(L->U0[b], L->U1[b], L->U2[b], cg->hdr.SB[b], cg->hdr.SBu[b], cg->hdr.SBuu[b], cg->hdr.QB[b], W->a[b], sigma,
                                    theta, want_ll, ft, fa, sc);
#pragma unroll
          for (int k = 0; k < NA; ++k) W->acc[b][k] = fa[k];
        }
      }
      // sloped buckets: on their 16 bucket nodes, two buckets per step (one per half warp), or lane-strided over the bucket's points
      const int half = lane >> 4, hl = lane & 15;
      auto node_pair = [&](int b0, int b1) {
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
      int pend = -1;
#pragma unroll 1
      for (int b = 0; b <= T; ++b) {
        const double c = W->c[b];
        if (c == 0.0) continue;                      // warp-uniform
        if (bnuw != nullptr && bnuw[b] > 0.0 && fabs(c) * bnuw[b] <= kThBinWidth) {
          if (pend < 0) { pend = b; continue; }
          node_pair(pend, b);
          pend = -1;
          continue;
        }
        const double a = W->a[b];
        double acc[NACC];
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
        const int p1 = pbstart[b + 1];
#pragma unroll 2
        for (int d = pbstart[b] + lane; d < p1; d += 32)
          point_contrib<MODE>(pn[d], s1[d], MODE == PASS_IRLS ? q[d] : 0.0, pu[d], a, c, sigma, theta, want_ll, ft, acc, sc);
#pragma unroll
        for (int k = 0; k < NA; ++k) {
          const double t = warp_sum(acc[k]);
          if (lane == 0) W->acc[b][k] = t;
        }
      }
      if (pend >= 0) node_pair(pend, -1);
    }
    if (NS > 0 && MODE != PASS_SWEEP) hist_terms<MODE>(sc, sigma, theta, want_ll);
    if (NS > 0) {
#pragma unroll
      for (int k = 0; k < NS; ++k) {
        const double t = warp_sum(sc[k]);
        if (lane == 0) W->sc[k] = t;
      }
    }
    if (MODE == PASS_SWEEP)
      for (int b = lane; b <= T; b += 32) { W->acc[b][5] = cg->hdr.SB[b]; W->acc[b][6] = cg->hdr.SBu[b]; }
    __syncwarp();
  }
};

enum WgStage : int { WG_FORWARD = 0, WG_BACKWARD = 1, WG_FINAL = 2, WG_OFFNULL = 3 };

// Persistent kernel: grid = resident CTAs of the device; every warp loops over genes pulled from the work queue.
template <int STAGE>
__global__ void __launch_bounds__(kWgThreads, SCLANE_WG_MINB) k_wg_stage(WgArgs WA) {
  __shared__ Lineage Ls;
  __shared__ FastTabs ftab;
  __shared__ int pbs[NBMAX + 1];
  __shared__ WgGeneShared wg[kWgWarps];
  __shared__ double bnuw_s[NBMAX];
  const CompArgs& A = WA.C;
  {
    for (int i = threadIdx.x; i < NBMAX; i += kWgThreads) bnuw_s[i] = (A.bn != nullptr && A.bn->elig[i]) ? A.bn->uw[i] : -1.0;
    const int* src = reinterpret_cast<const int*>(A.lineage);
    int* dst = reinterpret_cast<int*>(&Ls);
    for (int i = threadIdx.x; i < (int)(sizeof(Lineage) / sizeof(int)); i += kWgThreads) dst[i] = src[i];
    for (int i = threadIdx.x; i <= NBMAX; i += kWgThreads) pbs[i] = A.pbstart ? A.pbstart[i] : 0;
    fast_tabs_load(&ftab, threadIdx.x, kWgThreads);
  }
  __syncthreads();                                  // the only CTA-wide barrier: the warps are independent from here on
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Work& W = wg[warp].W;
  GeneState& st = wg[warp].st;
  GeneStats& gs = wg[warp].gs;
  for (;;) {
    int pos = 0;
    if (lane == 0) pos = atomicAdd(WA.counter, 1);
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if (pos >= A.n_genes) break;
    const int g = A.order ? A.order[pos] : pos;
    const CompGeneDev* cg = &A.cg[g];
    if (cg->hdr.route == 0) continue;               // this gene takes the per-cell kernels
    {
      const int* s2 = reinterpret_cast<const int*>(&A.out[g].st);
      int* d2 = reinterpret_cast<int*>(&st);
      for (int i = lane; i < (int)(sizeof(GeneState) / sizeof(int)); i += 32) d2[i] = s2[i];
      const int* s3 = reinterpret_cast<const int*>(&A.stats[g]);
      int* d3 = reinterpret_cast<int*>(&gs);
      for (int i = lane; i < (int)(sizeof(GeneStats) / sizeof(int)); i += 32) d3[i] = s3[i];
      int* w = reinterpret_cast<int*>(&W);
      for (int i = lane; i < (int)(sizeof(Work) / sizeof(int)); i += 32) w[i] = 0;      // also: no term tables (tal = tga = nullptr)
    }
    __syncwarp();
    const long long t_start = clock64();
    const bool offmode = STAGE == WG_OFFNULL;
    WgExec ex{&W, &Ls, cg, &gs, A.pn, A.pu, pbs, A.s1 ? A.s1 + (size_t)g * (size_t)A.Dpad : nullptr, A.q ? A.q + (size_t)g * (size_t)A.Dpad : nullptr,
              A.tail + (size_t)g * (size_t)kHistCap, A.big + (size_t)g * (size_t)kBigCap, &ftab,
              offmode ? WA.on : nullptr, offmode ? WA.off_s + (size_t)g * (size_t)(kOffKMax * kOffJ) : nullptr,
              offmode ? WA.off_q + (size_t)g * (size_t)(kOffKMax * kOffJ) : nullptr, offmode ? WA.offg + g : nullptr};
    if (!offmode && A.bn != nullptr) { ex.bn = A.bn; ex.bnuw = bnuw_s; ex.bs = A.bn_s + (size_t)g * (size_t)(NBMAX * kOffJ); ex.bq = A.bn_q + (size_t)g * (size_t)(NBMAX * kOffJ); }
    if (STAGE == WG_FORWARD) {
      for (int it = 0; it < A.n_iter; ++it) {
        gene_forward_fit(ex, Ls, gs, st, W);
        gene_sweep(ex, Ls, gs, st, W, A.M, A.tols_score);
      }
      __syncwarp();
      if (lane == 0 && !st.fwd_done) { st.fwd_done = 1; st.fwd_stop = 4; }
    } else if (STAGE == WG_BACKWARD) {
      gene_backward(ex, Ls, gs, st, W);
      __syncwarp();
      if (lane == 0) st.done |= 1;
    } else if (STAGE == WG_FINAL) {
      gene_final(ex, Ls, gs, st, W, false, &A.out[g].alt, &A.out[g].null, 3);
      __syncwarp();
      if (lane == 0) st.done |= 2;
    } else {
      // intercept-only fits with the offset (null model; alternative when only the intercept survived); the per-cell
      // kernel fits the alternatives with hinges afterwards (gene_final part 2) and marks the gene done
      gene_final(ex, Ls, gs, st, W, true, &A.out[g].alt, &A.out[g].null, 1);
      __syncwarp();
      if (lane == 0) st.done |= (st.fin.n <= 1) ? 6 : 4;      // bit 2: intercept-only fits done; bit 1: nothing left for the per-cell kernel
    }
    __syncwarp();
    if (lane == 0) { st.passes += W.passes; st.cycles[STAGE == WG_OFFNULL ? 2 : STAGE] += clock64() - t_start; }
    __syncwarp();
    {
      int* d2 = reinterpret_cast<int*>(&A.out[g].st);
      const int* s2 = reinterpret_cast<const int*>(&st);
      for (int i = lane; i < (int)(sizeof(GeneState) / sizeof(int)); i += 32) d2[i] = s2[i];
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// Offset quadrature, per-gene aggregation: Chebyshev moments of the gene's counts over the offset bins -> node weights.
// One CTA per gene.  The cells are visited in bin order (OffDevPlan.oord: positions in the pseudotime-sorted row, ordered by
// bin), so a bin is a contiguous range: all threads stride over it with register accumulators, one block reduction per
// bin, fixed order -> bit-reproducible.
// ---------------------------------------------------------------------------------------------
struct OffAggArgs {
  const Lineage* lineage;
  const OffsetNodes* on;
  const int* oord;              // [N] sorted-cell position of the i-th cell in bin order
  const double* oxi_ord;        // [N] Chebyshev coordinate, bin order
  const double* off_ord;        // [N] offset, bin order
  const int* ys;                // [B][Npad]
  const CompGeneDev* cg;        // [B] (glm.fit start sums from k_aggregate)
  double* off_s;                // [B][kOffKMax * kOffJ]
  double* off_q;
  OffGeneDev* offg;
  int n_genes;
};

__global__ void __launch_bounds__(kThreads) k_off_aggregate(OffAggArgs A) {
  __shared__ double red[kWarps][2 * kOffJ];
  __shared__ double cmom[2][kOffJ];
  __shared__ double wtab[64];
  __shared__ double wo_s[kWarps];
  const int g = blockIdx.x;
  if (g >= A.n_genes) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const OffsetNodes& O = *A.on;
  const int K = O.K, Npad = A.lineage->Npad, T = A.lineage->T;
  const int* __restrict__ y = A.ys + (size_t)g * (size_t)Npad;
  double* __restrict__ os = A.off_s + (size_t)g * (size_t)(kOffKMax * kOffJ);
  double* __restrict__ oq = A.off_q + (size_t)g * (size_t)(kOffKMax * kOffJ);
  if (threadIdx.x < 64) { double w, wz; start_cell(threadIdx.x, w, wz); wtab[threadIdx.x] = w; }
  __syncthreads();
  double wo = 0.0;                                   // sum w0(y_i) o_i
#pragma unroll 1
  for (int k = 0; k < K; ++k) {
    double cy[kOffJ], cq[kOffJ];
#pragma unroll
    for (int m = 0; m < kOffJ; ++m) { cy[m] = 0.0; cq[m] = 0.0; }
    const int i1 = O.kstart[k + 1];
    for (int i = O.kstart[k] + threadIdx.x; i < i1; i += kThreads) {
      const int yi = y[A.oord[i]];
      double w0;
      if (yi < 64) w0 = wtab[yi]; else { double wz; start_cell(yi, w0, wz); }
      wo += w0 * A.off_ord[i];
      if (yi != 0) {
        const double v = (double)yi, v2 = v * v, xi = A.oxi_ord[i], x2 = xi + xi;
        double t0 = 1.0, t1 = xi;
        cy[0] += v; cq[0] += v2;
        cy[1] = fma(v, xi, cy[1]); cq[1] = fma(v2, xi, cq[1]);
#pragma unroll
        for (int m = 2; m < kOffJ; ++m) {
          const double t2 = fma(x2, t1, -t0);
          cy[m] = fma(v, t2, cy[m]); cq[m] = fma(v2, t2, cq[m]);
          t0 = t1; t1 = t2;
        }
      }
    }
#pragma unroll
    for (int m = 0; m < kOffJ; ++m) {
      const double a = warp_sum(cy[m]), b = warp_sum(cq[m]);
      if (lane == 0) { red[warp][m] = a; red[warp][kOffJ + m] = b; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * kOffJ) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += red[w][threadIdx.x];
      cmom[threadIdx.x / kOffJ][threadIdx.x % kOffJ] = s;
    }
    __syncthreads();
    if (threadIdx.x < 2 * kOffJ) {
      // moments -> node weights: out_j = (cm_0 + 2 sum_{m >= 1} T_m(xi_j) cm_m) / J
      const int which = threadIdx.x / kOffJ, j = threadIdx.x % kOffJ;
      double s = 0.0;
#pragma unroll
      for (int m = kOffJ - 1; m >= 1; --m) s += O.ctab[m][j] * cmom[which][m];
      const double v = (cmom[which][0] + 2.0 * s) / (double)kOffJ;
      (which == 0 ? os : oq)[k * kOffJ + j] = v;
    }
    __syncthreads();
  }
  wo = warp_sum(wo);
  if (lane == 0) wo_s[warp] = wo;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kWarps; ++w) t += wo_s[w];
    const CompGeneDev& cg = A.cg[g];
    double a0 = 0.0, b0 = 0.0;
    for (int b = 0; b <= T; ++b) { a0 += cg.ST[b][0]; b0 += cg.ST[b][3]; }
    A.offg[g].a0tot = a0;
    A.offg[g].b0tot = b0 - t;
  }
}

}  // namespace sclane
