// sclane_gee_device.cuh -- CUDA executor and kernel of the GEE mode (sm_100a).
//
// One CTA (8 warps) per gene runs the whole per-gene GEE path (gene_gee in sclane_gee.cuh): geem fits, the GEE score
// sweep, the backward Wald pass and the final / null fits.  The gene's counts are gathered once into the lineage's
// subject-contiguous cell order (k_gather_stats); the knot bucket (uint8), bucket-local coordinate and offset of each cell
// are shared by all genes and stay L2 resident.  Every pass is a coalesced strided loop over the cells of one subject:
// per-thread register accumulators for the p x p systems, a fixed-order block reduction at the end, and (sweep only)
// warp-private shared-memory slots for the per-bucket moments, so every sum is bit-reproducible run to run.
// The n_i x n_i working covariances of the reference are never formed (closed forms, see sclane_gee.cuh).
// The AR-1 passes RECOMPUTE a cell's neighbours (2-3 cell evaluations per cell); fetching them from the neighbouring lanes by
// shuffle instead was measured slower (C3: 159 ms vs 136 ms) -- the recomputation is independent work that fills the FP64
// pipe's latency, the shuffles are not.
#pragma once
#include "sclane_device.cuh"
#include "sclane_gee.cuh"

namespace sclane {

constexpr int kGeeRed = 16;                       // widest block reduction (15 = p(p+1)/2 at p = 5)
constexpr int kGeeSlotV = 2 + 2 * GEE_PMAX;       // g, g u, h_l, h_l u

struct GeeGeneOutDev {
  GeneState st;
  GeeResult alt, null;
};

struct GeeArgs {
  const Lineage* lineage;
  const GeeLineage* glin;
  const int* ys;                 // [B][Npad] counts in lineage (subject-contiguous) order
  const unsigned char* bid;      // [Npad]
  const double* u;               // [Npad]
  const double* off;             // [Npad] or nullptr
  const GeneStats* stats;
  GeeGeneOutDev* out;
  int n_genes, M, use_offset, cor, gee_test, sandwich;
  double tols_score, mean_off;
};

struct GeeShared {
  Work W;
  TermTabs tt;
  GeeWork G;
  Lineage L;
  GeeLineage GL;
  GeneState st;
  GeneStats gs;
  double red[kWarps][kGeeRed];
  double slot[kWarps][NBMAX][kGeeSlotV];
};

struct GeeDevExec {
  static constexpr bool kTabs = true;      // term tables attached to Work (sclane_core.cuh term_coef)
  GeeShared* S;
  const int* __restrict__ y;
  const unsigned char* __restrict__ bid;
  const double* __restrict__ u;
  const double* __restrict__ off;

  __device__ __forceinline__ bool leader() const { return threadIdx.x == 0; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  __device__ __forceinline__ PassTabs* tabs() const { return nullptr; }
  template <class F>
  __device__ __forceinline__ void par_for(int n, F f) const {
    for (int i = threadIdx.x; i < n; i += kThreads) f(i);
  }

  // fixed-order block reduction of NV per-thread values into dst[0..NV)
  template <int NV>
  __device__ __forceinline__ void block_reduce(const double (&v)[NV], double* dst, bool accumulate) {
    static_assert(NV <= kGeeRed, "block_reduce width");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const double t = warp_sum(v[k]);
      if (lane == 0) S->red[warp][k] = t;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += S->red[w][threadIdx.x];
      if (accumulate) dst[threadIdx.x] += s; else dst[threadIdx.x] = s;
    }
    __syncthreads();
  }

  // add NV per-lane values into the warp's slot of their bucket; lanes of a 32-cell group usually share one bucket
  template <int NV>
  __device__ __forceinline__ void slot_add(const double (&v)[NV], int b, bool valid) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b0 = __shfl_sync(0xffffffffu, b, 0);
    if (__all_sync(0xffffffffu, !valid || b == b0)) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const double t = warp_sum(valid ? v[k] : 0.0);
        if (lane == 0) S->slot[warp][b0][k] += t;
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
          if (lane == 0) S->slot[warp][bb][k] += t;
        }
        rem &= ~same;
      }
    }
  }
  template <int NV>
  __device__ __forceinline__ void slot_zero() {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = lane; i < (S->L.T + 1) * NV; i += 32) S->slot[warp][i / NV][i % NV] = 0.0;
    __syncwarp();
  }

  __device__ __forceinline__ double offset(int j) const { return (S->W.use_offset && off) ? off[j] : 0.0; }

  __device__ __noinline__ void pass_ymom() {
    slot_zero<2>();
    const int N = S->L.N;
    for (int base = (threadIdx.x >> 5) * 32; base < N; base += kThreads) {
      const int j = base + (threadIdx.x & 31);
      const bool valid = j < N;
      double v[2] = {0.0, 0.0};
      int b = 0;
      if (valid) { b = bid[j]; const double yy = (double)y[j]; v[0] = yy; v[1] = yy * u[j]; }
      slot_add<2>(v, b, valid);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < (S->L.T + 1) * 2; idx += kThreads) {
      const int b = idx >> 1, k = idx & 1;
      double s = 0.0;
      for (int w = 0; w < kWarps; ++w) s += S->slot[w][b][k];
      S->W.acc[b][5 + k] = s;
    }
    __syncthreads();
  }

  __device__ __noinline__ void pass_resid() {
    const Work& W = S->W;
    GeeWork& G = S->G;
    const GeeLineage& GL = S->GL;
    const bool exch = G.cor == GEE_EXCH, ar1 = G.cor == GEE_AR1;
    double acc[2] = {0.0, 0.0};
    for (int i = 0; i < GL.S; ++i) {
      const int s0 = GL.sstart[i], s1 = GL.sstart[i + 1];
      double se[1] = {0.0};
      for (int j = s0 + threadIdx.x; j < s1; j += kThreads) {
        GeeCell c;
        gee_cell<false>(W, G, 0, y[j], bid[j], u[j], offset(j), c);
        acc[0] += c.r * c.r;
        se[0] += c.r;
        if (ar1 && j > s0) {
          GeeCell cp;
          gee_cell<false>(W, G, 0, y[j - 1], bid[j - 1], u[j - 1], offset(j - 1), cp);
          acc[1] += cp.r * c.r;
        }
      }
      if (exch) block_reduce<1>(se, &G.ssum[i][0], false);
    }
    block_reduce<2>(acc, G.sc, false);
  }

  template <int P>
  __device__ __noinline__ void pass_beta() {
    constexpr int NS = P * (P + 1) / 2;
    const Work& W = S->W;
    GeeWork& G = S->G;
    const GeeLineage& GL = S->GL;
    const bool exch = G.cor == GEE_EXCH, ar1 = G.cor == GEE_AR1;
    double A1[NS], A2[NS], e1[P], e2[P];
#pragma unroll
    for (int k = 0; k < NS; ++k) { A1[k] = 0.0; A2[k] = 0.0; }
#pragma unroll
    for (int k = 0; k < P; ++k) { e1[k] = 0.0; e2[k] = 0.0; }
    for (int i = 0; i < GL.S; ++i) {
      const int s0 = GL.sstart[i], s1 = GL.sstart[i + 1], n = s1 - s0;
      double sz[P + 1];
#pragma unroll
      for (int k = 0; k <= P; ++k) sz[k] = 0.0;
      for (int j = s0 + threadIdx.x; j < s1; j += kThreads) {
        GeeCell c;
        gee_cell<false>(W, G, P, y[j], bid[j], u[j], offset(j), c);
        const double cj = ar1 ? ar1_cj(j - s0, n, G.alpha) : 1.0;
        int idx = 0;
#pragma unroll
        for (int a = 0; a < P; ++a) {
          const double za = cj * c.z[a];
#pragma unroll
          for (int l = 0; l <= a; ++l) A1[idx++] += za * c.z[l];
          e1[a] += za * c.r;
        }
        if (ar1 && j > s0) {
          GeeCell cp;
          gee_cell<false>(W, G, P, y[j - 1], bid[j - 1], u[j - 1], offset(j - 1), cp);
          idx = 0;
#pragma unroll
          for (int a = 0; a < P; ++a) {
#pragma unroll
            for (int l = 0; l <= a; ++l) A2[idx++] += c.z[a] * cp.z[l] + cp.z[a] * c.z[l];
            e2[a] += c.z[a] * cp.r + cp.z[a] * c.r;
          }
        }
        if (exch) {
#pragma unroll
          for (int a = 0; a < P; ++a) sz[a] += c.z[a];
          sz[P] += c.r;
        }
      }
      if (exch) block_reduce<P + 1>(sz, G.ssum[i], false);
    }
    block_reduce<NS>(A1, G.A1, false);
    block_reduce<P>(e1, G.e1, false);
    if (ar1) {
      block_reduce<NS>(A2, G.A2, false);
      block_reduce<P>(e2, G.e2, false);
    }
  }

  // per-subject pieces of the sandwich's e_i = Z_i' R_i^-1 rr_i (combined by gee_sandwich)
  template <int P>
  __device__ __noinline__ void pass_sand() {
    const Work& W = S->W;
    GeeWork& G = S->G;
    const GeeLineage& GL = S->GL;
    const bool exch = G.cor == GEE_EXCH, ar1 = G.cor == GEE_AR1;
    for (int i = 0; i < GL.S; ++i) {
      const int s0 = GL.sstart[i], s1 = GL.sstart[i + 1], n = s1 - s0;
      double v1[P], v2[P + 1];
#pragma unroll
      for (int k = 0; k < P; ++k) v1[k] = 0.0;
#pragma unroll
      for (int k = 0; k <= P; ++k) v2[k] = 0.0;
      for (int j = s0 + threadIdx.x; j < s1; j += kThreads) {
        GeeCell c;
        gee_cell<false>(W, G, P, y[j], bid[j], u[j], offset(j), c);
        const double cj = ar1 ? ar1_cj(j - s0, n, G.alpha) : 1.0;
#pragma unroll
        for (int a = 0; a < P; ++a) v1[a] += cj * c.z[a] * c.r;
        if (ar1 && j > s0) {
          GeeCell cp;
          gee_cell<false>(W, G, P, y[j - 1], bid[j - 1], u[j - 1], offset(j - 1), cp);
#pragma unroll
          for (int a = 0; a < P; ++a) v2[a] += c.z[a] * cp.r + cp.z[a] * c.r;
        }
        if (exch) {
#pragma unroll
          for (int a = 0; a < P; ++a) v2[a] += c.z[a];
          v2[P] += c.r;
        }
      }
      block_reduce<P>(v1, &G.ssum[i][0], false);
      block_reduce<P + 1>(v2, &G.ssum[i][P], false);
    }
  }

  template <int P>
  __device__ __noinline__ void pass_prep1() {
    const Work& W = S->W;
    GeeWork& G = S->G;
    const GeeLineage& GL = S->GL;
    for (int i = 0; i < GL.S; ++i) {
      double sz[P + 1];
#pragma unroll
      for (int k = 0; k <= P; ++k) sz[k] = 0.0;
      for (int j = GL.sstart[i] + threadIdx.x; j < GL.sstart[i + 1]; j += kThreads) {
        GeeCell c;
        gee_cell<true>(W, G, P, y[j], bid[j], u[j], 0.0, c);
        sz[0] += c.r;
#pragma unroll
        for (int a = 0; a < P; ++a) sz[1 + a] += c.z[a];
      }
      block_reduce<P + 1>(sz, G.ssum[i], false);
    }
  }

  template <int P>
  __device__ __noinline__ void pass_prep2() {
    constexpr int NS = P * (P + 1) / 2;
    constexpr int NV = 2 + 2 * P;
    const Work& W = S->W;
    GeeWork& G = S->G;
    const GeeLineage& GL = S->GL;
    const int i = G.subject;
    const int s0 = GL.sstart[i], s1 = GL.sstart[i + 1], n = s1 - s0;
    const bool ar1 = G.cor == GEE_AR1;
    slot_zero<NV>();
    double j11[NS], ai[P];
#pragma unroll
    for (int k = 0; k < NS; ++k) j11[k] = 0.0;
#pragma unroll
    for (int k = 0; k < P; ++k) ai[k] = 0.0;
    for (int base = s0 + (threadIdx.x >> 5) * 32; base < s1; base += kThreads) {
      const int j = base + (threadIdx.x & 31);
      const bool valid = j < s1;
      double v[NV];
#pragma unroll
      for (int k = 0; k < NV; ++k) v[k] = 0.0;
      int b = 0;
      if (valid) {
        b = bid[j];
        const double uj = u[j];
        GeeCell c, cp, cn;
        gee_cell<true>(W, G, P, y[j], b, uj, 0.0, c);
        const bool hp = ar1 && j > s0, hn = ar1 && j < s1 - 1;
        if (hp) gee_cell<true>(W, G, P, y[j - 1], bid[j - 1], u[j - 1], 0.0, cp);
        if (hn) gee_cell<true>(W, G, P, y[j + 1], bid[j + 1], u[j + 1], 0.0, cn);
        GeePrepOut o;
        gee_prep_cell(W, G, P, i, n, j - s0, b, uj, c, hp ? &cp : nullptr, hn ? &cn : nullptr, o);
        int idx = 0;
#pragma unroll
        for (int a = 0; a < P; ++a) {
#pragma unroll
          for (int l = 0; l <= a; ++l) j11[idx++] += o.xm[a] * o.q[l];
          ai[a] += o.xm[a] * o.uvec;
          v[2 + 2 * a] = o.h[a];
          v[3 + 2 * a] = o.h[a] * uj;
        }
        v[0] = o.g;
        v[1] = o.g * uj;
      }
      slot_add<NV>(v, b, valid);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < (S->L.T + 1) * NV; idx += kThreads) {
      const int b = idx / NV, k = idx - b * NV;
      double s = 0.0;
      for (int w = 0; w < kWarps; ++w) s += S->slot[w][b][k];
      if (k < 2) G.Gm[b][k] = s; else G.Hm[b][k - 2] += s;
    }
    block_reduce<NS>(j11, G.A1, true);
    block_reduce<P>(ai, G.e1, false);
  }

  template <int GP>
  __device__ __forceinline__ void gee_pass() {
    __syncthreads();
    const int p = S->W.gee_p;
    if (GP == GPASS_YMOM) pass_ymom();
    else if (GP == GPASS_RESID) pass_resid();
    else if (GP == GPASS_BETA) {
      switch (p) {
        case 1: pass_beta<1>(); break;
        case 2: pass_beta<2>(); break;
        case 3: pass_beta<3>(); break;
        case 4: pass_beta<4>(); break;
        default: pass_beta<5>(); break;
      }
    } else if (GP == GPASS_SAND) {
      switch (p) {
        case 1: pass_sand<1>(); break;
        case 2: pass_sand<2>(); break;
        case 3: pass_sand<3>(); break;
        case 4: pass_sand<4>(); break;
        default: pass_sand<5>(); break;
      }
    } else if (GP == GPASS_PREP1) {
      switch (p) {
        case 1: pass_prep1<1>(); break;
        case 2: pass_prep1<2>(); break;
        case 3: pass_prep1<3>(); break;
        case 4: pass_prep1<4>(); break;
        default: pass_prep1<5>(); break;
      }
    } else {
      switch (p) {
        case 1: pass_prep2<1>(); break;
        case 2: pass_prep2<2>(); break;
        case 3: pass_prep2<3>(); break;
        case 4: pass_prep2<4>(); break;
        default: pass_prep2<5>(); break;
      }
    }
    __syncthreads();
  }
};

#ifndef SCLANE_GEE_MINB
#define SCLANE_GEE_MINB 2
#endif
__global__ void __launch_bounds__(kThreads, SCLANE_GEE_MINB) k_gee(GeeArgs A) {
  extern __shared__ __align__(16) unsigned char gee_smem[];
  GeeShared* S = reinterpret_cast<GeeShared*>(gee_smem);
  const int g = blockIdx.x;
  if (g >= A.n_genes) return;
  {
    const int* src = reinterpret_cast<const int*>(A.lineage);
    int* dst = reinterpret_cast<int*>(&S->L);
    for (int i = threadIdx.x; i < (int)(sizeof(Lineage) / sizeof(int)); i += kThreads) dst[i] = src[i];
    const int* s1 = reinterpret_cast<const int*>(A.glin);
    int* d1 = reinterpret_cast<int*>(&S->GL);
    for (int i = threadIdx.x; i < (int)(sizeof(GeeLineage) / sizeof(int)); i += kThreads) d1[i] = s1[i];
    const int* s3 = reinterpret_cast<const int*>(&A.stats[g]);
    int* d3 = reinterpret_cast<int*>(&S->gs);
    for (int i = threadIdx.x; i < (int)(sizeof(GeneStats) / sizeof(int)); i += kThreads) d3[i] = s3[i];
    int* w = reinterpret_cast<int*>(&S->W);
    for (int i = threadIdx.x; i < (int)(sizeof(Work) / sizeof(int)); i += kThreads) w[i] = 0;
    int* gw = reinterpret_cast<int*>(&S->G);
    for (int i = threadIdx.x; i < (int)(sizeof(GeeWork) / sizeof(int)); i += kThreads) gw[i] = 0;
    int* z = reinterpret_cast<int*>(&S->st);
    for (int i = threadIdx.x; i < (int)(sizeof(GeneState) / sizeof(int)); i += kThreads) z[i] = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    work_attach_tabs(S->W, &S->tt);
    S->st.fwd.n = 1;
    S->st.fwd.kind[0] = SCLANE_TERM_INTERCEPT;
    S->st.m_count = 1;
    if (S->gs.sum_y == 0.0) S->st.error |= SCLANE_NOTE_ALL_ZERO;
    if (S->gs.pad) S->st.error |= SCLANE_NOTE_BAD_COUNTS;
    S->G.cor = A.cor;
  }
  __syncthreads();
  const size_t row = (size_t)g * (size_t)S->L.Npad;
  GeeDevExec ex{S, A.ys + row, A.bid, A.u, A.off};
  gene_gee(ex, S->L, S->GL, S->gs, S->st, S->W, S->G, A.M, A.tols_score, A.use_offset != 0, A.mean_off, A.gee_test, A.sandwich != 0, &A.out[g].alt, &A.out[g].null);
  __syncthreads();
  if (threadIdx.x == 0) S->st.passes += S->W.passes;
  __syncthreads();
  {
    int* d2 = reinterpret_cast<int*>(&A.out[g].st);
    const int* s2 = reinterpret_cast<const int*>(&S->st);
    for (int i = threadIdx.x; i < (int)(sizeof(GeneState) / sizeof(int)); i += kThreads) d2[i] = s2[i];
  }
}

}  // namespace sclane
