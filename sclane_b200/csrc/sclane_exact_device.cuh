// sclane_exact_device.cuh -- approx.knot = FALSE (marge2.R:167-173): every abscissa kept by min_span() / max_span() is a
// candidate knot -- ~10^4 of them at 200,000 cells instead of 25.  The reference rebuilds two hinge columns and calls
// score_fun_glm() once (iteration 1) or twice per candidate; here one forward iteration of a gene costs
//   D point evaluations  (w_d = n mu/(1 + mu sigma), r_d = (s - n mu)/(1 + mu sigma), score_fun_glm.R:40,44)
// + two scans over the D abscissae that give, for EVERY abscissa p at once, the moments of w and r over the points to its
//   right and to its left in the coordinate (x - X_p):  M_k(p) = sum_{d > p} w_d (X_d - X_p)^k,  L_k(p) = sum_{d < p} w_d (X_p - X_d)^k
//   (k = 0, 1, 2; k = 0, 1 for r).  Both scans only ever ADD positive terms (the right moments are built from the right end,
//   the left moments from the left end, and moving the origin towards the data is the shift M1 += delta M0, M2 += 2 delta M1 +
//   delta^2 M0), so there is no cancellation however close a knot is to an end of the lineage;
// + O(p) algebra per candidate: tp1(x, t) = (x - t)+ lives on d > p and tp2 on d < p, and a product with a model hinge at t_j is
//   a quadratic in (x - t) on a sub-range, i.e. a combination of M(p), M(p_j) or L(p), L(p_j).
// The selection tree of marge2.R:368-587 only needs, per score vector, whether it is all NA, its maximum, and the first / last
// index of the maximum of round(score, 6): those are reduced over the candidates in index order (select_tree()).
//
// Fits: the knots a gene has selected so far define ITS buckets (<= 4 knots: <= 5 buckets), kept in a per-gene Lineage in shared
// memory; the unchanged algorithm of sclane_core.cuh runs on it with the X-compressed executor (u = X_d - ref[bucket] on the fly).
// One persistent CTA per resident slot, genes pulled from a work queue; forward pass, backward pass and (without an offset) the
// final fits of a gene run back to back in the same CTA.
#pragma once
#include "sclane_comp_device.cuh"

namespace sclane {

constexpr int kExMom = 10;      // per abscissa: M0 M1 M2 R0 R1 (right) | L0 L1 L2 Q0 Q1 (left)

struct ExactArgs {
  CompArgs C;                   // lineage (T = 0), point tables, per-gene aggregates, out
  const double* px;             // [D] abscissa per point
  const int* cand;              // [n_cand] candidate knots as point indices, ascending
  int n_cand;
  double* scratch;              // [slots][(2 + kExMom) * Dpad]: w, r and the scan results of the gene a CTA is working on
  int* counter;                 // work queue
  int do_final;                 // 1 = final + null glm.nb here (no offset); 0 = stop after the backward pass
};

// moments of a set of points in the coordinate (x - o): shift the origin by delta >= 0 AWAY from the points' side
struct Mom5 { double m0, m1, m2, r0, r1; };
__device__ __forceinline__ Mom5 mom_shift(const Mom5& a, double delta) {
  Mom5 o;
  o.m0 = a.m0;
  o.m1 = fma(delta, a.m0, a.m1);
  o.m2 = fma(delta, fma(delta, a.m0, 2.0 * a.m1), a.m2);
  o.r0 = a.r0;
  o.r1 = fma(delta, a.r0, a.r1);
  return o;
}
__device__ __forceinline__ Mom5 mom_add(const Mom5& a, const Mom5& b) { return Mom5{a.m0 + b.m0, a.m1 + b.m1, a.m2 + b.m2, a.r0 + b.r0, a.r1 + b.r1}; }
__device__ __forceinline__ Mom5 mom_point(double w, double r, double dist) { return Mom5{w, w * dist, w * dist * dist, r, r * dist}; }

// summary of one score vector for the selection tree, reduced in candidate order
struct ScoreSummary {
  int any;             // some entry is not NA
  double mx;           // max over the non-NA entries (unrounded)
  double mxr;          // max of round(score, 6)
  int first, last;     // first / last index attaining mxr
  double v_first, v_last;   // the unrounded scores there (score2 of marge2.R:665-675 is the same number)
};
__device__ __forceinline__ void ss_init(ScoreSummary& s) { s.any = 0; s.mx = -INFINITY; s.mxr = -INFINITY; s.first = -1; s.last = -1; s.v_first = nan(""); s.v_last = nan(""); }
__device__ __forceinline__ void ss_push(ScoreSummary& s, int idx, double v) {      // indices arrive in ascending order
  if (!(v == v)) return;
  const double r = r_round(v, 1e6);
  s.any = 1;
  if (v > s.mx) s.mx = v;
  if (s.first < 0 || r > s.mxr) { s.mxr = r; s.first = idx; s.last = idx; s.v_first = v; s.v_last = v; }
  else if (r == s.mxr) { s.last = idx; s.v_last = v; }
}
__device__ __forceinline__ void ss_merge(ScoreSummary& a, const ScoreSummary& b) {    // every index of b is larger than every index of a
  if (!b.any) return;
  if (!a.any) { a = b; return; }
  if (b.mx > a.mx) a.mx = b.mx;
  if (b.mxr > a.mxr) { a.mxr = b.mxr; a.first = b.first; a.last = b.last; a.v_first = b.v_first; a.v_last = b.v_last; }
  else if (b.mxr == a.mxr) { a.last = b.last; a.v_last = b.v_last; }
}

struct ExactShared {
  Work W;
  TermTabs tt;
  Lineage Lg;                   // the gene's own buckets
  GeneState st;
  GeneStats gs;
  CompGeneDev cg;               // bucket aggregates for the gene's buckets
  int pbs[NBMAX + 1];
  double scw[kCompWarps][NSC];
  FastTabs ft;
  Mom5 chunk[kCompThreads];     // per-thread chunk moments / carries of the scans
  ScoreSummary sum1[kCompThreads], sum2[kCompThreads];
  int mk_pidx[PMAX];            // point index of the knot of every model term (-1: intercept)
  double Ainv[PMAX * PMAX];     // (B'WB)^-1 of the current model
  double Ag[PMAX];              // (B'WB)^-1 B'r: the null fit's leftover score, projected out of the candidates' b
  int sel_trunc, sel_idx;       // selection broadcast
  double sel_score;
};

// (Re)build the gene's Lineage and bucket aggregates from its knot table st.ex_pidx[0 .. ex_T): bucket b = points
// [pbs[b], pbs[b + 1]), b = #{k : t_k <= x}.
__device__ void exact_rebuild(ExactShared& S, const ExactArgs& A, const Lineage* Lglob, const double* s1, const double* q, const double* a0p,
                              const double* b0p) {
  const CompArgs& C = A.C;
  __syncthreads();                                     // the knot table may just have been updated by thread 0
  const int T = S.st.ex_T;
  if (threadIdx.x == 0) {
    Lineage& L = S.Lg;
    L.N = Lglob->N; L.Npad = Lglob->Npad; L.T = T;
    for (int k = 0; k < T; ++k) L.knots[k] = S.st.ex_knots[k];
    L.ref[0] = T > 0 ? L.knots[0] : A.px[0];
    for (int b = 1; b <= T; ++b) L.ref[b] = L.knots[b - 1];
    S.pbs[0] = 0;
    for (int b = 1; b <= T; ++b) S.pbs[b] = S.st.ex_pidx[b - 1];
    for (int b = T + 1; b <= NBMAX; ++b) S.pbs[b] = C.D;
    for (int b = 0; b <= NBMAX; ++b) L.bstart[b] = C.pstart[S.pbs[b]];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b = warp; b <= T; b += kCompWarps) {
    const double ref = S.Lg.ref[b];
    double v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = 0.0;
    for (int d = S.pbs[b] + lane; d < S.pbs[b + 1]; d += 32) {
      const double u = A.px[d] - ref, n = C.pn[d], s = s1[d];
      v[0] += n; v[1] += n * u; v[2] += n * u * u;
      v[3] += s; v[4] += s * u; v[5] += s * u * u; v[6] += q[d];
      const double a0 = a0p[d], b0 = b0p[d];
      v[7] += a0; v[8] += a0 * u; v[9] += a0 * u * u; v[10] += b0; v[11] += b0 * u;
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = warp_sum(v[k]);
    if (lane == 0) {
      S.Lg.U0[b] = v[0]; S.Lg.U1[b] = v[1]; S.Lg.U2[b] = v[2];
      S.cg.hdr.SB[b] = v[3]; S.cg.hdr.SBu[b] = v[4]; S.cg.hdr.SBuu[b] = v[5]; S.cg.hdr.QB[b] = v[6];
      for (int k = 0; k < 5; ++k) S.cg.ST[b][k] = v[7 + k];
    }
  }
  __syncthreads();
}

// C[q][j], D[q], b[q] of candidate abscissa p against the current model, from the scan results (see the header comment)
__device__ __forceinline__ double exact_score(const ExactShared& S, const double* __restrict__ mom, const double* __restrict__ px, int p, bool both) {
  const Model& m = S.st.fwd;
  const int np = m.n;
  const double t = px[p];
  const double* R = mom + (size_t)p * kExMom;          // right moments at p
  const double* Lm = R + 5;                             // left moments at p
  // Rule 1: structural rank check
  bool has_pair = false;
  for (int j = 1; j < np; ++j) {
    if (S.mk_pidx[j] == p && (m.kind[j] == SCLANE_TERM_TP1 || both)) return nan("");
    if (m.kind[j] == SCLANE_TERM_TP1)
      for (int l = 1; l < np; ++l) if (m.kind[l] == SCLANE_TERM_TP2 && S.mk_pidx[l] == S.mk_pidx[j]) has_pair = true;
  }
  if (both && has_pair) return nan("");
  const int c = both ? 2 : 1;
  double Cq[2][PMAX];
  for (int j = 0; j < np; ++j) {
    double c1, c2 = 0.0;
    if (m.kind[j] == SCLANE_TERM_INTERCEPT) { c1 = R[1]; c2 = Lm[1]; }
    else {
      const int pj = S.mk_pidx[j];
      const double tj = px[pj];
      const double* Rj = mom + (size_t)pj * kExMom;
      const double* Lj = Rj + 5;
      if (m.kind[j] == SCLANE_TERM_TP1) {
        // (x - t)+ (x - t_j)+ on d > max(p, p_j);  (t - x)+ (x - t_j)+ on p_j < d < p
        c1 = pj <= p ? fma(t - tj, R[1], R[2]) : fma(tj - t, Rj[1], Rj[2]);
        if (pj < p) { const double dl = t - tj; c2 = fma(dl, Lm[1], -Lm[2]) + fma(dl, Lj[1], Lj[2]); }
      } else {
        // (x - t)+ (t_j - x)+ on p < d < p_j;  (t - x)+ (t_j - x)+ on d < min(p, p_j)
        c1 = 0.0;
        if (pj > p) { const double dl = tj - t; c1 = fma(dl, R[1], -R[2]) + fma(dl, Rj[1], Rj[2]); }
        c2 = pj >= p ? fma(tj - t, Lm[1], Lm[2]) : fma(t - tj, Lj[1], Lj[2]);
      }
    }
    Cq[0][j] = c1; Cq[1][j] = c2;
  }
  const double Dq[2] = {R[2], Lm[2]};
  double bq[2] = {R[4], Lm[4]};
  for (int j = 0; j < np; ++j) { bq[0] -= Cq[0][j] * S.Ag[j]; bq[1] -= Cq[1][j] * S.Ag[j]; }
  double Sm[2][2] = {{0, 0}, {0, 0}};
  for (int a = 0; a < c; ++a)
    for (int b = 0; b <= a; ++b) {
      double tt = 0.0;
      for (int j = 0; j < np; ++j) {
        double aj = 0.0;
        for (int l = 0; l < np; ++l) aj += S.Ainv[j * PMAX + l] * Cq[b][l];
        tt += Cq[a][j] * aj;
      }
      Sm[a][b] = (a == b ? Dq[a] : 0.0) - tt;
    }
  return llt_quadform(c, Sm[0][0], Sm[1][0], Sm[1][1], bq[0], bq[1]);
}

// One forward iteration's sweep over every candidate abscissa + selection + stat_out check + bookkeeping (marge2.R:182-758).
__device__ void exact_sweep(ExactShared& S, CompDevExec& ex, const ExactArgs& A, const Lineage* Lglob, const double* s1, const double* q,
                            const double* a0p, const double* b0p, double* scratch) {
  const CompArgs& C = A.C;
  Work& W = S.W;
  GeneState& st = S.st;
  const int D = C.D, Dpad = C.Dpad;
  __syncthreads();
  if (st.fwd_done || st.error) return;
  double* wv = scratch;
  double* rv = scratch + Dpad;
  double* mom = scratch + 2 * (size_t)Dpad;
  if (threadIdx.x == 0) {
    W.sigma = st.sigma; W.theta = st.sigma > 0.0 ? 1.0 / st.sigma : 0.0; W.use_offset = 0; W.passes++;
    for (int j = 0; j < st.fwd.n; ++j) S.mk_pidx[j] = st.fwd.kind[j] == SCLANE_TERM_INTERCEPT ? -1 : st.ex_pidx[st.fwd.knot[j]];
  }
  __syncthreads();
  // bucket moments of (w, r) on the gene's buckets -> B'WB and its inverse (stat_out_score_glm_null.R:33-40)
  coop_set_model(ex, S.Lg, st.fwd, W);
  coop_set_eta(ex, S.Lg, st.fwd, W, st.beta);
  ex.template pass<PASS_SWEEP>();
  coop_system(ex, S.Lg, st.fwd, W, 0, 3, -1);            // B'WB and B'r
  if (threadIdx.x == 0) {
    W.flag = sweep_invert(st.fwd, W) ? 1 : 0;
    for (int j = 0; j < st.fwd.n; ++j) for (int l = 0; l < st.fwd.n; ++l) S.Ainv[j * PMAX + l] = W.cov[j * NSYS + l];
    // B'r is zero at the exact MLE; what the Newton iteration left of it is projected out of every candidate's b below.  Among
    // ~10^4 candidates some are collinear with the basis to 1e-8 (a knot a few cells away from an end of the lineage): their
    // complement S is tiny and even a 1e-7 leak of B'r into b would dominate the score.
    for (int j = 0; j < st.fwd.n; ++j) {
      double v = 0.0;
      for (int l = 0; l < st.fwd.n; ++l) v += W.cov[j * NSYS + l] * W.grad[l];
      S.Ag[j] = v;                                       // A (B'r)
    }
  }
  __syncthreads();
  if (!W.flag) {
    if (threadIdx.x == 0) { st.error |= SCLANE_NOTE_NUMERIC; st.fwd_done = 1; st.fwd_stop = 5; }
    __syncthreads();
    return;
  }
  // per abscissa: w, r
  {
    const double sigma = W.sigma;
    const int T = S.Lg.T;
    for (int b = 0; b <= T; ++b) {
      const double a = W.a[b], c = W.c[b], ref = S.Lg.ref[b];
      for (int d = S.pbs[b] + threadIdx.x; d < S.pbs[b + 1]; d += kCompThreads) {
        const double mu = fast_exp(fma(c, A.px[d] - ref, a), S.ft);
        const double inv = rcp_pos(fma(sigma, mu, 1.0));
        const double n = C.pn[d];
        wv[d] = n * mu * inv;
        rv[d] = (s1[d] - n * mu) * inv;
      }
    }
  }
  __syncthreads();
  // the two scans.  Thread i owns the contiguous chunk [c0, c1) of abscissae.
  const int chunk = (D + kCompThreads - 1) / kCompThreads;
  const int c0 = min(D, (int)threadIdx.x * chunk), c1 = min(D, c0 + chunk);
  // (1) right moments: chunk totals relative to the chunk's FIRST abscissa (built from the chunk's right end)
  {
    Mom5 acc{0, 0, 0, 0, 0};
    for (int d = c1 - 1; d >= c0; --d) {
      if (d + 1 < c1) acc = mom_shift(acc, A.px[d + 1] - A.px[d]);
      acc = mom_add(acc, Mom5{wv[d], 0.0, 0.0, rv[d], 0.0});
    }
    S.chunk[threadIdx.x] = acc;                          // points of the chunk, origin px[c0]
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // carry[i] = moments of every point right of chunk i, origin = first abscissa of chunk i + 1 (serial: 256 small steps)
    Mom5 carry{0, 0, 0, 0, 0};
    int next_c0 = D;
    for (int i = kCompThreads - 1; i >= 0; --i) {
      const int i0 = min(D, i * chunk);
      const Mom5 own = S.chunk[i];
      S.chunk[i] = carry;                                // origin px[next_c0] (unused when next_c0 == D: all zero)
      if (i0 < D) {
        if (next_c0 < D) carry = mom_shift(carry, A.px[next_c0] - A.px[i0]);
        carry = mom_add(carry, own);
        next_c0 = i0;
      }
    }
  }
  __syncthreads();
  if (c0 < c1) {
    // walk the chunk right to left: cur = moments of the points right of d, origin px[d]
    Mom5 cur = S.chunk[threadIdx.x];
    if (c1 < D) cur = mom_shift(cur, A.px[c1] - A.px[c1 - 1]); else cur = Mom5{0, 0, 0, 0, 0};
    for (int d = c1 - 1; d >= c0; --d) {
      double* o = mom + (size_t)d * kExMom;
      o[0] = cur.m0; o[1] = cur.m1; o[2] = cur.m2; o[3] = cur.r0; o[4] = cur.r1;
      if (d > c0) {
        const double dl = A.px[d] - A.px[d - 1];
        cur = mom_add(mom_shift(cur, dl), mom_point(wv[d], rv[d], dl));
      }
    }
  }
  __syncthreads();
  // (2) left moments, mirrored: coordinate (X_p - x) >= 0
  {
    Mom5 acc{0, 0, 0, 0, 0};
    for (int d = c0; d < c1; ++d) {
      if (d > c0) acc = mom_shift(acc, A.px[d] - A.px[d - 1]);
      acc = mom_add(acc, Mom5{wv[d], 0.0, 0.0, rv[d], 0.0});
    }
    S.chunk[threadIdx.x] = acc;                          // points of the chunk, origin px[c1 - 1]
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Mom5 carry{0, 0, 0, 0, 0};
    int prev_last = -1;
    for (int i = 0; i < kCompThreads; ++i) {
      const int i0 = min(D, i * chunk), i1 = min(D, i0 + chunk);
      const Mom5 own = S.chunk[i];
      S.chunk[i] = carry;                                // every point left of chunk i, origin px[prev_last]
      if (i0 < i1) {
        if (prev_last >= 0) carry = mom_shift(carry, A.px[i1 - 1] - A.px[prev_last]);
        carry = mom_add(carry, own);
        prev_last = i1 - 1;
      }
    }
  }
  __syncthreads();
  if (c0 < c1) {
    Mom5 cur = S.chunk[threadIdx.x];
    if (c0 > 0) cur = mom_shift(cur, A.px[c0] - A.px[c0 - 1]); else cur = Mom5{0, 0, 0, 0, 0};
    for (int d = c0; d < c1; ++d) {
      double* o = mom + (size_t)d * kExMom + 5;
      o[0] = cur.m0; o[1] = cur.m1; o[2] = cur.m2; o[3] = cur.r0; o[4] = cur.r1;
      if (d + 1 < c1) {
        const double dl = A.px[d + 1] - A.px[d];
        cur = mom_add(mom_shift(cur, dl), mom_point(wv[d], rv[d], dl));
      }
    }
  }
  __syncthreads();
  // scores of every candidate, reduced to the summaries the selection tree needs (thread i: a contiguous run of candidates)
  const int nc = A.n_cand;
  {
    const int per = (nc + kCompThreads - 1) / kCompThreads;
    const int k0 = min(nc, (int)threadIdx.x * per), k1 = min(nc, k0 + per);
    ScoreSummary one, both;
    ss_init(one); ss_init(both);
    for (int k = k0; k < k1; ++k) {
      const int p = A.cand[k];
      ss_push(one, k, exact_score(S, mom, A.px, p, false));
      ss_push(both, k, exact_score(S, mom, A.px, p, true));
    }
    S.sum1[threadIdx.x] = one; S.sum2[threadIdx.x] = both;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ScoreSummary one = S.sum1[0], both = S.sum2[0];
    for (int i = 1; i < kCompThreads; ++i) { ss_merge(one, S.sum1[i]); ss_merge(both, S.sum2[i]); }
    // the selection tree (marge2.R:368-587) on the summaries; with q = 1 the "interaction" vectors are the constant -1e5
    // (iteration 1) or duplicates of the additive ones (later iterations)
    const int it = st.fwd_iters;
    st.fwd_iters++;
    W.flag = 0;
    const int in_set = st.fwd.n > 1 ? 1 : 0;
    VecSummary vo{!one.any, one.mx, one.first, one.last}, vb{!both.any, both.mx, both.first, both.last};
    VecSummary ci{false, -1e5, 0, nc - 1};
    const VecSummary& oi = in_set ? vo : ci;
    const VecSummary& bi = in_set ? vb : ci;
    int trunc = 0, vec = 0, last = 0;
    if (!select_tree(bi, oi, vb, vo, trunc, vec, last)) { st.fwd_done = 1; st.fwd_stop = 1; }
    else {
      // the index comes from the chosen vector (the constant interaction vector of iteration 1: first = 0, last = nc - 1);
      // score2 is the re-score of the chosen shape at that knot (marge2.R:665-675)
      const bool const_vec = (vec >= SEL_ONE_INT) && !in_set;
      const bool use_both = (vec == SEL_BOTH || vec == SEL_BOTH_INT);
      const ScoreSummary& sv = use_both ? both : one;
      const int kidx = const_vec ? (last ? nc - 1 : 0) : (sv.any ? (last ? sv.last : sv.first) : -1);
      const double score2 = kidx >= 0 ? exact_score(S, mom, A.px, A.cand[kidx], trunc == 2) : nan("");
      if (kidx < 0) { st.fwd_done = 1; st.fwd_stop = 5; st.error |= SCLANE_NOTE_NUMERIC; }
      else {
        if (it < 4) st.fwd_score[it] = score2;
        if (!(score2 == score2)) { st.fwd_done = 1; st.fwd_stop = 5; st.error |= SCLANE_NOTE_NUMERIC; }
        else { S.sel_trunc = trunc; S.sel_idx = A.cand[kidx]; S.sel_score = score2; W.flag = 1; }
      }
    }
  }
  __syncthreads();
  if (!W.flag) return;
  // insert the knot into the gene's table (ascending), remap the model's knot indices, build the candidate model
  if (threadIdx.x == 0) {
    const int p = S.sel_idx;
    int pos = 0;
    bool present = false;
    while (pos < st.ex_T && st.ex_pidx[pos] < p) ++pos;
    if (pos < st.ex_T && st.ex_pidx[pos] == p) present = true;
    if (!present) {
      for (int k = st.ex_T; k > pos; --k) { st.ex_pidx[k] = st.ex_pidx[k - 1]; st.ex_knots[k] = st.ex_knots[k - 1]; }
      st.ex_pidx[pos] = p; st.ex_knots[pos] = A.px[p];
      st.ex_T++;
      for (int j = 1; j < st.fwd.n; ++j) if (st.fwd.knot[j] >= pos) st.fwd.knot[j]++;
    }
    Model cand = st.fwd;
    cand.kind[cand.n] = SCLANE_TERM_TP1; cand.knot[cand.n] = pos; cand.n++;
    if (S.sel_trunc == 2) { cand.kind[cand.n] = SCLANE_TERM_TP2; cand.knot[cand.n] = pos; cand.n++; }
    W.bw_cur = cand;
    W.sw_score2 = S.sel_score;
  }
  exact_rebuild(S, A, Lglob, s1, q, a0p, b0p);
  // stat_out() on the candidate basis (stat_out.R:23-47) and acceptance (marge2.R:676-758)
  for (int b = threadIdx.x; b <= S.Lg.T; b += kCompThreads) { W.acc[b][5] = S.cg.hdr.SB[b]; W.acc[b][6] = S.cg.hdr.SBu[b]; }
  __syncthreads();
  StatOutFn so{&S.Lg, &W.bw_cur, &W};
  ex.par_for(W.bw_cur.n * (W.bw_cur.n + 1) / 2 + W.bw_cur.n, so);
  __syncthreads();
  if (threadIdx.x == 0) forward_accept(S.Lg, st, W, S.gs, C.M, C.tols_score);
  __syncthreads();
}

constexpr int kExactCtasPerSm = 16 / kCompWarps;       // 16 warps per SM whatever the CTA shape of the compressed kernels
__global__ void __launch_bounds__(kCompThreads, kExactCtasPerSm) k_exact(ExactArgs A) {
  extern __shared__ __align__(16) unsigned char ex_smem[];
  ExactShared& S = *reinterpret_cast<ExactShared*>(ex_smem);
  __shared__ int g_s;
  const CompArgs& C = A.C;
  fast_tabs_load(&S.ft, threadIdx.x, kCompThreads);
  double* scratch = A.scratch + (size_t)blockIdx.x * (size_t)((2 + kExMom) * C.Dpad);
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) g_s = atomicAdd(A.counter, 1);
    __syncthreads();
    const int g = g_s;
    if (g >= C.n_genes) break;
    if (C.cg[g].hdr.route == 0) continue;              // more than kBigCap counts >= kHistCap: not handled in exact-knot mode (error record)
    {
      const int* s2 = reinterpret_cast<const int*>(&C.out[g].st);
      int* d2 = reinterpret_cast<int*>(&S.st);
      for (int i = threadIdx.x; i < (int)(sizeof(GeneState) / sizeof(int)); i += kCompThreads) d2[i] = s2[i];
      const int* s3 = reinterpret_cast<const int*>(&C.stats[g]);
      int* d3 = reinterpret_cast<int*>(&S.gs);
      for (int i = threadIdx.x; i < (int)(sizeof(GeneStats) / sizeof(int)); i += kCompThreads) d3[i] = s3[i];
      const int* s4 = reinterpret_cast<const int*>(&C.cg[g]);
      int* d4 = reinterpret_cast<int*>(&S.cg);
      for (int i = threadIdx.x; i < (int)(sizeof(CompGeneDev) / sizeof(int)); i += kCompThreads) d4[i] = s4[i];
      int* w = reinterpret_cast<int*>(&S.W);
      for (int i = threadIdx.x; i < (int)(sizeof(Work) / sizeof(int)); i += kCompThreads) w[i] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) { work_attach_tabs(S.W, &S.tt); S.st.ex_T = 0; }
    __syncthreads();
    const long long t_start = clock64();
    const double* s1 = C.s1 + (size_t)g * (size_t)C.Dpad;
    const double* q = C.q + (size_t)g * (size_t)C.Dpad;
    const double* a0p = C.a0p + (size_t)g * (size_t)C.Dpad;
    const double* b0p = C.b0p + (size_t)g * (size_t)C.Dpad;
    exact_rebuild(S, A, C.lineage, s1, q, a0p, b0p);
    CompDevExec ex{&S.W, &S.Lg, &S.cg, &S.gs, C.D, C.pn, C.pu, S.pbs, s1, q, C.tail + (size_t)g * (size_t)kHistCap, C.big + (size_t)g * (size_t)kBigCap,
                   S.scw, &S.ft, A.px};
    for (int it = 0; it < C.n_iter; ++it) {
      gene_forward_fit(ex, S.Lg, S.gs, S.st, S.W);
      exact_sweep(S, ex, A, C.lineage, s1, q, a0p, b0p, scratch);
    }
    __syncthreads();
    if (threadIdx.x == 0 && !S.st.fwd_done) { S.st.fwd_done = 1; S.st.fwd_stop = 4; }
    __syncthreads();
    gene_backward(ex, S.Lg, S.gs, S.st, S.W);
    __syncthreads();
    if (threadIdx.x == 0) S.st.done |= 1;
    if (A.do_final) {
      gene_final(ex, S.Lg, S.gs, S.st, S.W, false, &C.out[g].alt, &C.out[g].null, 3);
      __syncthreads();
      if (threadIdx.x == 0) S.st.done |= 2;
    }
    __syncthreads();
    if (threadIdx.x == 0) { S.st.passes += S.W.passes; S.st.cycles[0] += clock64() - t_start; }
    __syncthreads();
    {
      int* d2 = reinterpret_cast<int*>(&C.out[g].st);
      const int* s2 = reinterpret_cast<const int*>(&S.st);
      for (int i = threadIdx.x; i < (int)(sizeof(GeneState) / sizeof(int)); i += kCompThreads) d2[i] = s2[i];
    }
  }
}

}  // namespace sclane
