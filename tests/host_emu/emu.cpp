// emu.cpp -- TEST HARNESS ONLY (never linked into libsclane_b200.so).
// Runs the shared host/device per-gene algorithm (sclane_b200/csrc/sclane_core.cuh) with a serial
// executor, so the control flow, the bucket-moment algebra and the glm.nb schedule can be checked
// against the numpy oracle on a machine without a GPU (`pytest -m "not gpu"`).  The product path has
// no CPU fallback: libsclane_b200.so only contains the CUDA executor.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../sclane_b200/csrc/sclane_comp.cuh"
#include "../../sclane_b200/csrc/sclane_gee_host.hpp"

using namespace sclane;

struct HostExec {
  static constexpr bool kTabs = true;      // term tables attached to Work (sclane_core.cuh term_coef)
  Work* W;
  const Lineage* L;
  const int32_t* y;        // sorted counts of this gene [N]
  const double* u;
  const double* off;       // may be null
  double* mu;              // [N]
  const int32_t* bucket;
  PassTabs tabs_;
  PassTabs* tabs() { return &tabs_; }
  bool leader() const { return true; }
  void sync() const {}
  template <int MODE>
  void pass() {
    const int N = L->N;
    if (MODE != PASS_EMIT) {
      for (int b = 0; b <= L->T; ++b) for (int k = 0; k < NACC; ++k) W->acc[b][k] = 0.0;
      for (int k = 0; k < NSC; ++k) W->sc[k] = 0.0;
    }
    for (int i = 0; i < N; ++i) {
      const int b = bucket[i];
      double mu_out = 0.0;
      const double o = (W->use_offset && off) ? off[i] : 0.0;
      cell_contrib<MODE>(y[i], u[i], o, W->a[b], W->c[b], W->sigma, W->theta, tabs_, nullptr, mu ? mu[i] : 0.0, W->want_ll != 0, W->acc[b], W->sc, mu_out);
      if (MODE == PASS_EMIT) mu[i] = mu_out;
    }
  }
  template <class F>
  void par_for(int n, F f) { for (int i = 0; i < n; ++i) f(i); }
};

// ---- X-compressed executor (sclane_comp.cuh), serial: defines the semantics of the compressed passes ----
struct CompHostExec {
  static constexpr bool kTabs = false;      // term tables attached to Work (sclane_core.cuh term_coef)
  Work* W;
  const Lineage* L;
  const LineagePlan* P;
  std::vector<double> s1, q, a0, b0, tail;     // per point / per count
  std::vector<int32_t> big;                    // counts >= kHistCap, ascending
  CompGeneHdr H;
  double sum_ylogy;
  int hmax_ = 0;
  // offset quadrature (intercept-only fits with an offset): node weights of this gene
  bool offnull = false;
  std::vector<double> ons, onq;
  double a0tot = 0.0, b0tot = 0.0, wo = 0.0;
  void aggregate_off(const int32_t* y) {                   // y: sorted counts; P->ON.on must be set
    const OffsetNodes& O = P->ON;
    std::vector<double> cy((size_t)O.K * kOffJ, 0.0), cq((size_t)O.K * kOffJ, 0.0);
    a0tot = b0tot = wo = 0.0;
    for (int i = 0; i < L->N; ++i) {
      double w, wz;
      start_cell(y[i], w, wz);
      a0tot += w; b0tot += wz; wo += w * P->off[(size_t)i];
      if (y[i] == 0) continue;
      const double v = (double)y[i], xi = P->oxi[(size_t)i];
      const size_t base = (size_t)P->obin[(size_t)i] * kOffJ;
      double t0 = 1.0, t1 = xi;
      cy[base] += v; cq[base] += v * v;
      cy[base + 1] += v * xi; cq[base + 1] += v * v * xi;
      for (int m = 2; m < kOffJ; ++m) { const double t2 = 2.0 * xi * t1 - t0; cy[base + m] += v * t2; cq[base + m] += v * v * t2; t0 = t1; t1 = t2; }
    }
    ons.assign((size_t)O.K * kOffJ, 0.0); onq.assign((size_t)O.K * kOffJ, 0.0);
    for (int k = 0; k < O.K; ++k) {
      off_moments_to_nodes(O, &cy[(size_t)k * kOffJ], &ons[(size_t)k * kOffJ]);
      off_moments_to_nodes(O, &cq[(size_t)k * kOffJ], &onq[(size_t)k * kOffJ]);
    }
  }
  template <int MODE>
  void pass_offnull() {
    const OffsetNodes& O = P->ON;
    for (int b = 0; b <= L->T; ++b) for (int k = 0; k < NACC; ++k) W->acc[b][k] = 0.0;
    for (int k = 0; k < NSC; ++k) W->sc[k] = 0.0;
    if (MODE == PASS_IRLS_START) {
      W->acc[0][0] = a0tot;
      W->acc[0][3] = b0tot - wo;
      W->sc[0] = -sum_ylogy + (double)H.n_zero * W->theta * log1p(W->sigma / 6.0);
      W->sc[1] = (double)H.n_zero;
      return;
    }
    const bool want_ll = W->want_ll != 0;
    for (int e = 0; e < O.K * kOffJ; ++e)
      offnode_contrib<MODE>(O.node_n[e], ons[(size_t)e], onq[(size_t)e], O.node_o[e], W->a[0], W->sigma, W->theta, want_ll, nullptr, W->acc[0], W->sc);
    for (int k = 0; k < hmax_; ++k) hist_contrib<MODE>(k, tail[(size_t)k], tail[(size_t)k + 1], W->sigma, W->theta, want_ll, W->sc);
    for (int32_t v : big) big_contrib<MODE>(v, W->sigma, W->theta, want_ll, W->sc);
  }
  PassTabs* tabs() { return nullptr; }
  bool leader() const { return true; }
  void sync() const {}
  template <class F>
  void par_for(int n, F f) { for (int i = 0; i < n; ++i) f(i); }
  // theta.ml on Chebyshev nodes: serial twin of CompDevExec::theta_prepare (sclane_comp.cuh)
  bool use_theta_nodes = true;
  ThetaNodes tn;
  long long theta_node_calls = 0, theta_node_entries = 0, theta_point_calls = 0;
  void theta_prepare() {
    tn.on = 0;
    if (!use_theta_nodes || offnull) return;
    static double ctab[kOffJ][kOffJ];
    static bool have = false;
    if (!have) { for (int m = 0; m < kOffJ; ++m) for (int j = 0; j < kOffJ; ++j) ctab[m][j] = std::cos(M_PI * (double)(m * (2 * j + 1)) / (double)(2 * kOffJ)); have = true; }
    auto u_of = [&](int, int d) { return P->pu[(size_t)d]; };
    theta_nodes_layout(tn, L->T, W->c, P->pbstart, u_of);
    if (!tn.on) return;
    for (int b = 0; b <= L->T; ++b) {
      const int kb = tn.bins[b], p0 = P->pbstart[b], p1 = P->pbstart[b + 1], e0 = tn.first[b];
      if (kb == 0) { tn.en[e0] = L->U0[b]; tn.es[e0] = H.SB[b]; tn.eeta[e0] = W->a[b]; continue; }
      if (kb < 0) { for (int d = p0; d < p1; ++d) { const int e = e0 + (d - p0); tn.en[e] = P->pn[(size_t)d]; tn.es[e] = s1[(size_t)d]; tn.eeta[e] = W->a[b] + W->c[b] * P->pu[(size_t)d]; } continue; }
      int lo = p0;
      for (int k = 0; k < kb; ++k) {
        const double u1 = tn.ulo[b] + (double)(k + 1) * tn.h[b];
        int hi = p1;
        if (k + 1 < kb) { hi = lo; while (hi < p1 && P->pu[(size_t)hi] < u1) ++hi; }
        theta_nodes_bin_serial(tn, b, k, lo, hi, P->pn.data(), s1.data(), W->a[b], W->c[b], ctab, u_of);
        lo = hi;
      }
    }
  }
  void aggregate(const int32_t* y) {                       // y: sorted counts
    const int D = P->D;
    s1.assign((size_t)D, 0.0); q.assign((size_t)D, 0.0); a0.assign((size_t)D, 0.0); b0.assign((size_t)D, 0.0);
    std::memset(&H, 0, sizeof(H));
    sum_ylogy = 0.0;
    for (int i = 0; i < L->N; ++i) { if (y[i] > H.ymax) H.ymax = y[i]; if (y[i] == 0) H.n_zero++; if (y[i] > 0) sum_ylogy += (double)y[i] * std::log((double)y[i]); }
    big.clear();
    for (int i = 0; i < L->N; ++i) if (y[i] >= kHistCap) big.push_back(y[i]);
    std::sort(big.begin(), big.end());
    H.n_big = (int)big.size();
    H.route = H.n_big <= kBigCap ? 1 : 0;
    const int hmax = H.ymax < kHistCap ? H.ymax : kHistCap - 1;     // largest count that goes through the histogram
    std::vector<double> hist((size_t)hmax + 2, 0.0);
    for (int b = 0; b <= L->T; ++b)
      for (int d = P->pbstart[b]; d < P->pbstart[b + 1]; ++d) {
        const double u = P->pu[(size_t)d];
        for (int i = P->pstart[(size_t)d]; i < P->pstart[(size_t)d + 1]; ++i) {
          const double v = (double)y[i];
          double w, wz;
          start_cell(y[i], w, wz);
          s1[(size_t)d] += v; q[(size_t)d] += v * v; a0[(size_t)d] += w; b0[(size_t)d] += wz;
          if (y[i] < kHistCap) hist[(size_t)y[i]] += 1.0;
        }
        H.SB[b] += s1[(size_t)d]; H.SBu[b] += s1[(size_t)d] * u; H.SBuu[b] += s1[(size_t)d] * u * u; H.QB[b] += q[(size_t)d];
      }
    tail.assign((size_t)hmax + 1, 0.0);                    // tail[k] = #{k < y < kHistCap}, k = 0..hmax (tail[hmax] = 0)
    double c = 0.0;
    for (int k = hmax; k >= 0; --k) { tail[(size_t)k] = c; c += hist[(size_t)k]; }
    hmax_ = hmax;
  }
  template <int MODE>
  void pass() {
    if (MODE == PASS_EMIT) return;                         // the compressed sweep recomputes mu from the eta parameters
    if (offnull && (MODE == PASS_IRLS || MODE == PASS_THETA || MODE == PASS_IRLS_START)) { pass_offnull<MODE>(); return; }
    for (int b = 0; b <= L->T; ++b) for (int k = 0; k < (MODE == PASS_SWEEP ? 5 : NACC); ++k) W->acc[b][k] = 0.0;
    for (int k = 0; k < NSC; ++k) W->sc[k] = 0.0;
    if (MODE == PASS_IRLS_START) {
      for (int b = 0; b <= L->T; ++b)
        for (int d = P->pbstart[b]; d < P->pbstart[b + 1]; ++d) {
          const double u = P->pu[(size_t)d];
          W->acc[b][0] += a0[(size_t)d]; W->acc[b][1] += a0[(size_t)d] * u; W->acc[b][2] += a0[(size_t)d] * u * u;
          W->acc[b][3] += b0[(size_t)d]; W->acc[b][4] += b0[(size_t)d] * u;
        }
      W->sc[0] = -sum_ylogy + (double)H.n_zero * W->theta * log1p(W->sigma / 6.0);
      W->sc[1] = (double)H.n_zero;
      return;
    }
    const bool want_ll = W->want_ll != 0;
    if (MODE == PASS_THETA && tn.on) {
      double dummy[NACC];
      for (int e = 0; e < tn.n; ++e) point_contrib<MODE>(tn.en[e], tn.es[e], 0.0, 0.0, tn.eeta[e], 0.0, W->sigma, W->theta, want_ll, nullptr, dummy, W->sc);
      for (int k = 0; k < hmax_; ++k) hist_contrib<MODE>(k, tail[(size_t)k], tail[(size_t)k + 1], W->sigma, W->theta, want_ll, W->sc);
      for (int32_t v : big) big_contrib<MODE>(v, W->sigma, W->theta, want_ll, W->sc);
      theta_node_calls++; theta_node_entries += tn.n;
      return;
    }
    if (MODE == PASS_THETA) { bool sloped = false; for (int b = 0; b <= L->T; ++b) sloped = sloped || W->c[b] != 0.0; if (sloped) theta_point_calls++; }
    for (int b = 0; b <= L->T; ++b) {
      if (W->c[b] == 0.0) {
        flat_bucket_contrib<MODE>(L->U0[b], L->U1[b], L->U2[b], H.SB[b], H.SBu[b], H.SBuu[b], H.QB[b], W->a[b], W->sigma, W->theta, want_ll, nullptr, W->acc[b], W->sc);
      } else {
        for (int d = P->pbstart[b]; d < P->pbstart[b + 1]; ++d)
          point_contrib<MODE>(P->pn[(size_t)d], s1[(size_t)d], q[(size_t)d], P->pu[(size_t)d], W->a[b], W->c[b], W->sigma, W->theta, want_ll, nullptr, W->acc[b], W->sc);
      }
      if (MODE == PASS_SWEEP) { W->acc[b][5] = H.SB[b]; W->acc[b][6] = H.SBu[b]; }
    }
    if (MODE != PASS_SWEEP) {
      for (int k = 0; k < hmax_; ++k) hist_contrib<MODE>(k, tail[(size_t)k], tail[(size_t)k + 1], W->sigma, W->theta, want_ll, W->sc);
      for (int32_t v : big) big_contrib<MODE>(v, W->sigma, W->theta, want_ll, W->sc);
    }
  }
};

static void gene_stats(const int32_t* y, int N, GeneStats& gs) {
  gs.sum_y = gs.sum_y2 = gs.sum_ylogy = gs.sum_lgamma_y1 = 0.0;
  gs.ymax = 0;
  for (int i = 0; i < N; ++i) {
    const double v = (double)y[i];
    gs.sum_y += v; gs.sum_y2 += v * v;
    if (y[i] > 0) gs.sum_ylogy += v * std::log(v);
    gs.sum_lgamma_y1 += std::lgamma(v + 1.0);
    if (y[i] > gs.ymax) gs.ymax = y[i];
  }
}

extern "C" int emu_test_dynamic(const int32_t* counts, int64_t n_cells, int64_t n_genes, const double* pt, int32_t n_lineages,
                                const double* size_factor, const sclane_opts* opts, sclane_record* out,
                                sclane_lineage_info* lineages) {
  for (int l = 0; l < n_lineages; ++l) {
    LineagePlan P;
    if (!build_lineage_plan(pt + (int64_t)l * n_cells, n_cells, size_factor, *opts, P)) { std::fprintf(stderr, "emu: %s\n", P.error.c_str()); return 1; }
    const Lineage& L = P.L;
    if (lineages) {
      lineages[l].n_cells = L.N; lineages[l].n_knots = L.T; lineages[l].minspan = P.minspan;
      for (int k = 0; k < L.T; ++k) lineages[l].knots[k] = L.knots[k];
      lineages[l].pt_min = P.pt_min; lineages[l].pt_max = P.pt_max;
    }
    std::vector<int32_t> y((size_t)L.N);
    std::vector<double> mu((size_t)L.N);
    for (int64_t g = 0; g < n_genes; ++g) {
      for (int i = 0; i < L.N; ++i) y[(size_t)i] = counts[g * n_cells + P.perm[(size_t)i]];
      GeneStats gs;
      gene_stats(y.data(), L.N, gs);
      GeneOut go;
      init_gene_state(go.st);
      if (gs.sum_y == 0.0) go.st.error |= SCLANE_NOTE_ALL_ZERO;
      Work W;
      std::memset(&W, 0, sizeof(W));
      TermTabs tt;
      // HostExec reads the term tables (as k_stage / k_comp_stage do), CompHostExec computes term_ab() on the fly (as the
      // warp-per-gene kernels do): both forms of term_coef() run on the CPU
      work_attach_tabs(W, &tt);
      HostExec ex{&W, &L, y.data(), P.u.data(), P.off.empty() ? nullptr : P.off.data(), mu.data(), P.bucket.data(), PassTabs()};
      // opts->force_per_cell == 1: per-cell executor everywhere; otherwise the X-compressed executor runs every fit without
      // an offset (the routing the CUDA path uses)
      const bool comp = opts->force_per_cell != 1;
      CompHostExec cx{&W, &L, &P};
      { const char* e = std::getenv("SCLANE_EMU_THETA_NODES"); cx.use_theta_nodes = !(e && e[0] == '0'); }
      if (comp) cx.aggregate(y.data());
      const bool use_comp = comp && cx.H.route;
      for (int it = 0; it < 4 && !go.st.fwd_done && !go.st.error; ++it) {
        if (use_comp) { gene_forward_fit(cx, L, gs, go.st, W); gene_sweep(cx, L, gs, go.st, W, opts->M, opts->tols_score); }
        else { gene_forward_fit(ex, L, gs, go.st, W); gene_sweep(ex, L, gs, go.st, W, opts->M, opts->tols_score); }
      }
      const int p_fwd = W.passes;
      if (use_comp) gene_backward(cx, L, gs, go.st, W); else gene_backward(ex, L, gs, go.st, W);
      const int p_bwd = W.passes;
      if (use_comp && size_factor == nullptr) gene_final(cx, L, gs, go.st, W, false, &go.alt, &go.null, 3);
      else if (use_comp && P.ON.on) {
        // offsets: the intercept-only fits (null model; alternative when nothing but the intercept survived) through the
        // offset quadrature, the alternative with hinges per cell -- the routing of the CUDA path
        cx.aggregate_off(y.data());
        cx.offnull = true;
        gene_final(cx, L, gs, go.st, W, true, &go.alt, &go.null, 1);
        cx.offnull = false;
        gene_final(ex, L, gs, go.st, W, true, &go.alt, &go.null, 2);
      }
      else gene_final(ex, L, gs, go.st, W, size_factor != nullptr, &go.alt, &go.null, 3);
      if (std::getenv("SCLANE_EMU_STAGES"))
        std::fprintf(stderr, "gene %ld passes fwd %d bwd %d final %d (alt: alt=%d irls=%d; null: alt=%d irls=%d)\n", (long)g, p_fwd, p_bwd - p_fwd,
                     W.passes - p_bwd, go.alt.alternations, go.alt.irls_iters, go.null.alternations, go.null.irls_iters);
      if (std::getenv("SCLANE_EMU_STAGES"))
        std::fprintf(stderr, "   theta passes on nodes %lld (mean entries %.1f of %d points), on points %lld\n", cx.theta_node_calls,
                     cx.theta_node_calls ? (double)cx.theta_node_entries / (double)cx.theta_node_calls : 0.0, P.D, cx.theta_point_calls);
      go.st.passes = W.passes;
      finalize_record(go, L, out[g * n_lineages + l]);
    }
  }
  return 0;
}

// ---- GEE mode: serial executor over the subject-ordered cells (defines the semantics of the GEE passes) ----
struct GeeHostExec {
  static constexpr bool kTabs = true;      // term tables attached to Work (sclane_core.cuh term_coef)
  Work* W;
  GeeWork* G;
  const Lineage* L;
  const GeeLineage* GL;
  const int32_t* y;        // [N] lineage order
  const uint8_t* bid;
  const double* u;
  const double* off;       // may be null
  PassTabs tabs_;
  PassTabs* tabs() { return &tabs_; }
  bool leader() const { return true; }
  void sync() const {}
  template <class F>
  void par_for(int n, F f) { for (int i = 0; i < n; ++i) f(i); }
  double offset(int j) const { return (W->use_offset && off) ? off[j] : 0.0; }
  template <int GP>
  void gee_pass() {
    const int pp = W->gee_p;
    if (GP == GPASS_YMOM) {
      for (int b = 0; b <= L->T; ++b) for (int k = 0; k < NACC; ++k) W->acc[b][k] = 0.0;
      for (int j = 0; j < L->N; ++j) { W->acc[bid[j]][5] += (double)y[j]; W->acc[bid[j]][6] += (double)y[j] * u[j]; }
      return;
    }
    if (GP == GPASS_RESID) {
      G->sc[0] = G->sc[1] = 0.0;
      for (int i = 0; i < GL->S; ++i) {
        G->ssum[i][0] = 0.0;
        double prev = 0.0;
        for (int j = GL->sstart[i]; j < GL->sstart[i + 1]; ++j) {
          GeeCell c;
          gee_cell<false>(*W, *G, 0, y[j], bid[j], u[j], offset(j), c);
          G->sc[0] += c.r * c.r;
          if (j > GL->sstart[i]) G->sc[1] += prev * c.r;
          G->ssum[i][0] += c.r;
          prev = c.r;
        }
      }
      return;
    }
    if (GP == GPASS_BETA) {
      for (int k = 0; k < GEE_NSYM; ++k) { G->A1[k] = 0.0; G->A2[k] = 0.0; }
      for (int k = 0; k < GEE_PMAX; ++k) { G->e1[k] = 0.0; G->e2[k] = 0.0; }
      for (int i = 0; i < GL->S; ++i) {
        for (int k = 0; k <= GEE_PMAX; ++k) G->ssum[i][k] = 0.0;
        const int s0 = GL->sstart[i], n = GL->sstart[i + 1] - s0;
        GeeCell pc;
        for (int q = 0; q < n; ++q) {
          const int j = s0 + q;
          GeeCell c;
          gee_cell<false>(*W, *G, pp, y[j], bid[j], u[j], offset(j), c);
          const double cj = G->cor == GEE_AR1 ? ar1_cj(q, n, G->alpha) : 1.0;
          for (int a = 0; a < pp; ++a) {
            for (int l = 0; l <= a; ++l) G->A1[sym_idx(a, l)] += cj * c.z[a] * c.z[l];
            G->e1[a] += cj * c.z[a] * c.r;
          }
          if (G->cor == GEE_AR1 && q > 0)
            for (int a = 0; a < pp; ++a) {
              for (int l = 0; l <= a; ++l) G->A2[sym_idx(a, l)] += c.z[a] * pc.z[l] + pc.z[a] * c.z[l];
              G->e2[a] += c.z[a] * pc.r + pc.z[a] * c.r;
            }
          if (G->cor == GEE_EXCH) { for (int a = 0; a < pp; ++a) G->ssum[i][a] += c.z[a]; G->ssum[i][pp] += c.r; }
          pc = c;
        }
      }
      return;
    }
    if (GP == GPASS_SAND) {
      for (int i = 0; i < GL->S; ++i) {
        for (int k = 0; k <= 2 * GEE_PMAX; ++k) G->ssum[i][k] = 0.0;
        const int s0 = GL->sstart[i], n = GL->sstart[i + 1] - s0;
        GeeCell pc;
        for (int q = 0; q < n; ++q) {
          const int j = s0 + q;
          GeeCell c;
          gee_cell<false>(*W, *G, pp, y[j], bid[j], u[j], offset(j), c);
          const double cj = G->cor == GEE_AR1 ? ar1_cj(q, n, G->alpha) : 1.0;
          for (int a = 0; a < pp; ++a) G->ssum[i][a] += cj * c.z[a] * c.r;
          if (G->cor == GEE_AR1 && q > 0) for (int a = 0; a < pp; ++a) G->ssum[i][pp + a] += c.z[a] * pc.r + pc.z[a] * c.r;
          if (G->cor == GEE_EXCH) { for (int a = 0; a < pp; ++a) G->ssum[i][pp + a] += c.z[a]; G->ssum[i][2 * pp] += c.r; }
          pc = c;
        }
      }
      return;
    }
    if (GP == GPASS_PREP1) {
      for (int i = 0; i < GL->S; ++i) {
        for (int k = 0; k <= GEE_PMAX; ++k) G->ssum[i][k] = 0.0;
        for (int j = GL->sstart[i]; j < GL->sstart[i + 1]; ++j) {
          GeeCell c;
          gee_cell<true>(*W, *G, pp, y[j], bid[j], u[j], 0.0, c);
          G->ssum[i][0] += c.r;
          for (int a = 0; a < pp; ++a) G->ssum[i][1 + a] += c.z[a];
        }
      }
      return;
    }
    if (GP == GPASS_PREP2) {
      const int i = G->subject;
      const int s0 = GL->sstart[i], n = GL->sstart[i + 1] - s0;
      for (int b = 0; b <= L->T; ++b) G->Gm[b][0] = G->Gm[b][1] = 0.0;
      for (int k = 0; k < GEE_PMAX; ++k) G->e1[k] = 0.0;
      for (int q = 0; q < n; ++q) {
        const int j = s0 + q;
        GeeCell c, cp, cn;
        gee_cell<true>(*W, *G, pp, y[j], bid[j], u[j], 0.0, c);
        const bool hp = q > 0, hn = q < n - 1;
        if (G->cor == GEE_AR1) {
          if (hp) gee_cell<true>(*W, *G, pp, y[j - 1], bid[j - 1], u[j - 1], 0.0, cp);
          if (hn) gee_cell<true>(*W, *G, pp, y[j + 1], bid[j + 1], u[j + 1], 0.0, cn);
        }
        GeePrepOut o;
        gee_prep_cell(*W, *G, pp, i, n, q, bid[j], u[j], c, hp ? &cp : nullptr, hn ? &cn : nullptr, o);
        const int b = bid[j];
        for (int a = 0; a < pp; ++a) {
          for (int l2 = 0; l2 <= a; ++l2) G->A1[sym_idx(a, l2)] += o.xm[a] * o.q[l2];
          G->e1[a] += o.xm[a] * o.uvec;
          G->Hm[b][2 * a] += o.h[a];
          G->Hm[b][2 * a + 1] += o.h[a] * u[j];
        }
        G->Gm[b][0] += o.g;
        G->Gm[b][1] += o.g * u[j];
      }
      return;
    }
  }
};

extern "C" int emu_test_dynamic_gee(const int32_t* counts, int64_t n_cells, int64_t n_genes, const double* pt, int32_t n_lineages,
                                    const int32_t* subject_id, const double* size_factor, const sclane_opts* opts, sclane_record* out,
                                    sclane_lineage_info* lineages) {
  for (int l = 0; l < n_lineages; ++l) {
    LineagePlan P;
    if (!build_lineage_plan(pt + (int64_t)l * n_cells, n_cells, size_factor, *opts, P)) { std::fprintf(stderr, "emu: %s\n", P.error.c_str()); return 1; }
    GeePlan GP;
    if (!build_gee_plan(P, pt + (int64_t)l * n_cells, n_cells, subject_id, size_factor, GP)) { std::fprintf(stderr, "emu: %s\n", GP.error.c_str()); return 1; }
    const Lineage& L = P.L;
    if (lineages) {
      lineages[l].n_cells = L.N; lineages[l].n_knots = L.T; lineages[l].minspan = P.minspan;
      for (int k = 0; k < L.T; ++k) lineages[l].knots[k] = L.knots[k];
      lineages[l].pt_min = P.pt_min; lineages[l].pt_max = P.pt_max;
    }
    std::vector<int32_t> y((size_t)L.N);
    for (int64_t g = 0; g < n_genes; ++g) {
      for (int i = 0; i < L.N; ++i) y[(size_t)i] = counts[g * n_cells + GP.cells[(size_t)i]];
      GeneStats gs;
      gene_stats(y.data(), L.N, gs);
      GeeGeneOut go;
      init_gene_state(go.st);
      if (gs.sum_y == 0.0) go.st.error |= SCLANE_NOTE_ALL_ZERO;
      Work W;
      std::memset(&W, 0, sizeof(W));
      TermTabs tt;
      work_attach_tabs(W, &tt);
      GeeWork G;
      std::memset(&G, 0, sizeof(G));
      G.cor = opts->gee_cor;
      GeeHostExec ex{&W, &G, &L, &GP.GL, y.data(), GP.bid.data(), GP.u.data(), GP.off.empty() ? nullptr : GP.off.data(), PassTabs()};
      gene_gee(ex, L, GP.GL, gs, go.st, W, G, opts->M, opts->tols_score, size_factor != nullptr, GP.mean_off, opts->gee_test, opts->gee_bias_correction != 0, &go.alt, &go.null);
      go.st.passes = W.passes;
      finalize_record_gee(go, L, opts->gee_test, opts->gee_bias_correction, GP.GL.S, out[g * n_lineages + l]);
    }
  }
  return 0;
}

extern "C" int emu_candidate_knots(const double* x, int64_t n, const sclane_opts* opts, double* knots_out, int32_t* minspan_out) {
  std::vector<double> xv(x, x + n), X;
  int ms = 0;
  std::vector<double> k = candidate_knots(xv, opts->approx_knot != 0, opts->n_knot_max, opts->minspan, &X, &ms);
  if (minspan_out) *minspan_out = ms;
  for (size_t i = 0; i < k.size() && i < (size_t)SCLANE_TMAX; ++i) knots_out[i] = k[i];
  return (int)k.size();
}

extern "C" double emu_pchisq_upper(double q, double df) { return pchisq_upper(q, df); }
extern "C" double emu_r_round(double x, int digits) { return r_round_digits(x, digits); }
extern "C" double emu_digamma(double x) { return digamma_pos(x); }
extern "C" double emu_trigamma(double x) { return trigamma_pos(x); }
extern "C" void emu_psi_tri(int y, double th, double* out) { psi_tri_diff(y, th, out[0], out[1]); }
extern "C" double emu_nb_lambda(int y, double sigma) { return nb_lambda(y, sigma, 1.0 / sigma); }
// table-driven exp / log / log1p of the fit kernels (sclane_fastmath.cuh), vectorised over n arguments
extern "C" void emu_fastmath(int which, const double* x, double* out, long n) {
  static FastTabs ft;
  static bool have = false;
  if (!have) { fast_tabs_host(&ft); have = true; }
  for (long i = 0; i < n; ++i) {
    if (which == 0) out[i] = fast_exp(x[i], ft);
    else if (which == 1) out[i] = fast_log_ge1(x[i], ft);
    else { const double t = 1.0 + x[i]; out[i] = fast_log1p_pos(x[i], t, 1.0 / t, ft); }
  }
}
extern "C" void emu_default_opts(sclane_opts* o) {
  std::memset(o, 0, sizeof(*o));
  o->abi_version = SCLANE_ABI_VERSION; o->mode = 0; o->M = 5; o->approx_knot = 1; o->n_knot_max = 25; o->minspan = -1; o->tols_score = 1e-5; o->gee_cor = SCLANE_COR_AR1;
}

// finalize_record_gee() on a hand-made device result: alternative fitted (3 terms), null geem fit failed
extern "C" void emu_finalize_gee_null_failed(int gee_test, sclane_record* out) {
  Lineage L;
  std::memset(&L, 0, sizeof(L));
  L.N = 300; L.T = 3; L.knots[0] = 0.25; L.knots[1] = 0.5; L.knots[2] = 0.75;
  GeeGeneOut go;
  std::memset(&go, 0, sizeof(go));
  init_gene_state(go.st);
  go.st.fin.n = 3;
  go.st.fin.kind[0] = SCLANE_TERM_INTERCEPT; go.st.fin.kind[1] = SCLANE_TERM_TP1; go.st.fin.kind[2] = SCLANE_TERM_TP2;
  go.st.fin.knot[1] = 1; go.st.fin.knot[2] = 1;
  go.alt.ok = 1; go.alt.converged = 1; go.alt.niter = 4; go.alt.alpha = 0.3; go.alt.phi = 1.2; go.alt.score_stat = 7.0;
  go.alt.coef[0] = 1.0; go.alt.coef[1] = 0.5; go.alt.coef[2] = -0.25;
  for (int j = 0; j < 3; ++j) { go.alt.cov[j * GEE_PMAX + j] = 0.01; go.alt.cov_robust[j * GEE_PMAX + j] = 0.02; }
  go.null.ok = 0;
  finalize_record_gee(go, L, gee_test, 0, 3, *out);
}
