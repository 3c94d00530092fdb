// sclane_host.hpp -- host-side pieces of the path (plain C++17, no CUDA):
//   K0: candidate knots + pseudotime sort + bucket tables per lineage (marge2.R:150-173,
//       min_span.R:15-36, max_span.R:14-26), gene independent, so done once per lineage;
//   record finalisation: z / p-values, LRT, DF rule and pchisq (modelLRT.R:43-53,
//       pullMARGESummary.R:70-98, pullNullSummary.R:66-80).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <future>
#include <limits>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "sclane_comp.cuh"

namespace sclane {

// ---- host parallelism for the per-lineage plan (K0) ---------------------------------------------------
// fn(chunk, begin, end) over kPlanChunks FIXED chunks of [0, n): any per-chunk partial result is combined in chunk order
// afterwards, so every plan is bit-identical whatever the number of host threads (<= kPlanChunks) that ran it.
constexpr int kPlanChunks = 16;
// Host threads the library may use (plan, record finalisation, narrowing of the counts for the transfer): the machine's, unless
// SCLANE_HOST_THREADS says otherwise -- several processes sharing one host (one rank per GPU) should each take their share.
inline unsigned host_threads() {
  static const unsigned n = [] {
    const char* e = std::getenv("SCLANE_HOST_THREADS");
    const long v = e ? std::atol(e) : 0;
    if (v > 0) return (unsigned)v;
    const unsigned hw = std::thread::hardware_concurrency();
    return hw == 0 ? 4u : hw;
  }();
  return n;
}
template <class F>
inline void parallel_chunks(size_t n, F fn, size_t min_parallel = 32768) {
  unsigned hw = host_threads();
  int nt = (int)(hw == 0 ? 4 : (hw > (unsigned)kPlanChunks ? (unsigned)kPlanChunks : hw));
  if (n < min_parallel) nt = 1;
  auto run = [&](int t) {
    for (int c = t; c < kPlanChunks; c += nt) fn(c, n * (size_t)c / kPlanChunks, n * (size_t)(c + 1) / kPlanChunks);
  };
  if (nt == 1) { run(0); return; }
  std::vector<std::thread> th;
  th.reserve((size_t)nt);
  try {
    for (int t = 1; t < nt; ++t) th.emplace_back(run, t);
  } catch (...) { for (auto& x : th) x.join(); throw; }
  run(0);
  for (auto& x : th) x.join();
}

// ---- R scalar semantics ------------------------------------------------------------------------
inline double r_round_digits(double x, int digits) { return r_round(x, std::pow(10.0, digits)); }

inline double r_signif(double x, int digits) {           // nmath/fprec.c for ordinary magnitudes
  if (x == 0.0 || !std::isfinite(x)) return x;
  const int e10 = digits - 1 - (int)std::floor(std::log10(std::fabs(x)));
  if (e10 >= 0) { const double p = std::pow(10.0, e10); return std::nearbyint(x * p) / p; }
  const double p = std::pow(10.0, -e10);
  return std::nearbyint(x / p) * p;
}

// stats::quantile(type = 7) (marge2.R:158-159): the two neighbouring order statistics by selection (O(n), no full sort)
inline double quantile7(std::vector<double> xs, double p) {
  const size_t n = xs.size();
  const double index = 1.0 + (double)(n > 0 ? n - 1 : 0) * p;
  const size_t lo = (size_t)std::floor(index), hi = (size_t)std::ceil(index);
  std::nth_element(xs.begin(), xs.begin() + (lo - 1), xs.end());
  double q = xs[lo - 1];
  const double h = index - (double)lo;
  if (hi > lo) {
    const double xhi = *std::min_element(xs.begin() + lo, xs.end());      // the next order statistic
    if (index > (double)lo && xhi != q) q = (1.0 - h) * q + h * xhi;
  }
  return q;
}

// two quantiles (p_lo < p_hi) from ONE working copy: the second selection only looks at the part right of the first
inline void quantile7_pair(const std::vector<double>& x, double p_lo, double p_hi, double& q_lo, double& q_hi, std::vector<double>* ws = nullptr) {
  std::vector<double> local;
  std::vector<double>& xs = ws ? *ws : local;
  xs.assign(x.begin(), x.end());
  const size_t n = xs.size();
  auto one = [&](double p, size_t from) -> double {
    const double index = 1.0 + (double)(n > 0 ? n - 1 : 0) * p;
    const size_t lo = (size_t)std::floor(index), hi = (size_t)std::ceil(index);
    std::nth_element(xs.begin() + from, xs.begin() + (lo - 1), xs.end());
    double q = xs[lo - 1];
    const double h = index - (double)lo;
    if (hi > lo) {
      const double xhi = *std::min_element(xs.begin() + lo, xs.end());
      if (index > (double)lo && xhi != q) q = (1.0 - h) * q + h * xhi;
    }
    return q;
  };
  q_lo = one(p_lo, 0);
  const size_t lo_pos = (size_t)std::floor(1.0 + (double)(n > 0 ? n - 1 : 0) * p_lo) - 1;   // everything left of it is <= q_lo
  q_hi = one(p_hi, lo_pos);
}

// min_span.R:15-36 on an ASCENDING copy of round(x, 4)
inline std::vector<double> min_span_sorted(const std::vector<double>& x, int q, int minspan_in, int* minspan_out) {
  const size_t n = x.size();
  int minspan = minspan_in;
  if (minspan < 0) minspan = (int)std::nearbyint(-std::log2(-(1.0 / ((double)q * (double)n)) * std::log(1.0 - 0.05)) / 2.5);
  if (minspan_out) *minspan_out = minspan;
  std::vector<double> out;
  out.push_back(x[0]);
  size_t cc = 1;
  while (cc + (size_t)minspan <= n) {
    const size_t idx = cc + (size_t)minspan + 1;   // 1-based
    if (idx <= n) out.push_back(x[idx - 1]);
    cc += (size_t)minspan + 1;
  }
  std::vector<double> uniq;
  for (double v : out) if (uniq.empty() || uniq.back() != v) uniq.push_back(v);   // ascending input
  return uniq;
}

// max_span.R:14-26 on the ASCENDING UNIQUE values of round(x, 4)
inline std::vector<double> max_span_sorted_unique(const std::vector<double>& x, int q) {
  const long n = (long)x.size();
  const long maxspan = (long)std::nearbyint(3.0 - std::log2(0.05 / (double)q));
  const long tail_from = (long)std::floor((double)(n - maxspan + 1));   // 1-based
  std::vector<double> out;
  for (long i = 1; i <= n; ++i) {
    if (i <= maxspan) continue;
    if (i >= tail_from) continue;
    out.push_back(x[(size_t)i - 1]);
  }
  if (out.empty()) return x;
  return out;
}

inline std::vector<double> min_span(const std::vector<double>& X, int q, int minspan_in, int* minspan_out) {
  std::vector<double> x(X);
  std::sort(x.begin(), x.end());
  return min_span_sorted(x, q, minspan_in, minspan_out);
}

// max_span.R:14-26
inline std::vector<double> max_span(const std::vector<double>& X, int q) {
  std::vector<double> x(X);
  std::sort(x.begin(), x.end());
  x.erase(std::unique(x.begin(), x.end()), x.end());
  return max_span_sorted_unique(x, q);
}

// marge2.R:152-166 given xs = ASCENDING round(x, 4): returns the candidate knots (ascending)
// xs is consumed (reduced to its unique values in place)
inline std::vector<double> candidate_knots_sorted(const std::vector<double>& x_pred, std::vector<double>& xs, bool approx_knot,
                                                  int n_knot_max, int minspan_in, int* minspan_out,
                                                  std::future<std::pair<double, double>>* quantiles = nullptr) {
  std::vector<double> r1 = min_span_sorted(xs, 1, minspan_in, minspan_out);
  xs.erase(std::unique(xs.begin(), xs.end()), xs.end());
  std::vector<double> r2 = max_span_sorted_unique(xs, 1);
  std::vector<double> red;
  for (double v : r1) if (std::binary_search(r2.begin(), r2.end(), v)) red.push_back(v);
  if (approx_knot) {
    double q05, q95;
    if (quantiles) { const std::pair<double, double> qq = quantiles->get(); q05 = qq.first; q95 = qq.second; }     // computed concurrently by the caller
    else quantile7_pair(x_pred, 0.05, 0.95, q05, q95);
    std::vector<double> f;
    for (double v : red) if (v > q05 && v < q95) f.push_back(v);
    red.swap(f);
    if ((int)red.size() > n_knot_max) {
      const double lo = *std::min_element(red.begin(), red.end()), hi = *std::max_element(red.begin(), red.end());
      std::vector<double> s((size_t)n_knot_max);
      const double by = (hi - lo) / (double)(n_knot_max - 1);     // seq.default(from, to, length.out)
      for (int k = 0; k < n_knot_max; ++k) s[(size_t)k] = (k == 0) ? lo : (k == n_knot_max - 1 ? hi : lo + (double)k * by);
      for (double& v : s) v = r_round(v, 1e4);
      red.swap(s);
    }
  }
  return red;
}

// marge2.R:152-166: returns candidate knots (ascending); X_out = round(x, 4)
inline std::vector<double> candidate_knots(const std::vector<double>& x_pred, bool approx_knot, int n_knot_max,
                                           int minspan_in, std::vector<double>* X_out, int* minspan_out) {
  std::vector<double> X(x_pred.size());
  for (size_t i = 0; i < X.size(); ++i) X[i] = r_round(x_pred[i], 1e4);
  std::vector<double> xs(X);
  std::sort(xs.begin(), xs.end());
  std::vector<double> red = candidate_knots_sorted(x_pred, xs, approx_knot, n_knot_max, minspan_in, minspan_out);
  if (X_out) X_out->swap(X);
  return red;
}

// Stable order of the cells by X = round(x, 4).  X is a multiple of 1e-4, so llround(1e4 X) is an order-preserving integer
// key and a counting sort does it in O(N) when the key range is small (pseudotime in [0, 1]: 10,001 keys); otherwise a
// comparison sort of (X, index) pairs -- both give exactly the order of a stable sort by X.
inline void stable_order_by_x(const std::vector<double>& X, std::vector<int32_t>& order, std::vector<int32_t>& start, std::vector<int32_t>& key) {
  const size_t N = X.size();
  order.resize(N);
  double lo = X[0], hi = X[0];
  for (double v : X) { if (v < lo) lo = v; if (v > hi) hi = v; }
  const double span = (hi - lo) * 1e4;
  if (std::isfinite(span) && span <= 4.0 * (double)N + 1024.0 && std::fabs(lo) < 1e8 && std::fabs(hi) < 1e8) {
    const long long k0 = std::llround(lo * 1e4);
    const size_t R = (size_t)(std::llround(hi * 1e4) - k0) + 1;
    start.assign(R + 1, 0);
    key.resize(N);
    parallel_chunks(N, [&](int, size_t a, size_t b) { for (size_t i = a; i < b; ++i) key[i] = (int32_t)(std::llround(X[i] * 1e4) - k0); });
    for (size_t i = 0; i < N; ++i) start[(size_t)key[i] + 1]++;
    for (size_t k = 0; k < R; ++k) start[k + 1] += start[k];
    for (size_t i = 0; i < N; ++i) order[(size_t)start[(size_t)key[i]]++] = (int32_t)i;
    return;
  }
  std::vector<std::pair<double, int32_t>> keyed(N);
  for (size_t i = 0; i < N; ++i) keyed[i] = {X[i], (int32_t)i};
  std::sort(keyed.begin(), keyed.end());
  for (size_t i = 0; i < N; ++i) order[i] = keyed[i].second;
}
inline std::vector<int32_t> stable_order_by_x(const std::vector<double>& X) {
  std::vector<int32_t> order, start, key;
  stable_order_by_x(X, order, start, key);
  return order;
}

// ---- lineage plan ---------------------------------------------------------------------------------
struct LineagePlan {
  Lineage L;
  std::vector<int32_t> perm;      // sorted position -> original cell index (within the full cell axis)
  std::vector<double> u;          // [Npad] bucket-local coordinate
  std::vector<double> off;        // [Npad] log(1/size_factor) or empty
  std::vector<int32_t> bucket;    // [N] bucket of each sorted cell (host harness only)
  std::vector<double> Xs;         // [N] sorted rounded pseudotime
  // X-compressed layout (sclane_comp.cuh): distinct rounded pseudotime values of the sorted cells
  int D = 0;
  std::vector<int32_t> pstart;    // [D + 1] range of sorted cells of point d
  std::vector<double> pn, pu;     // [D] cells per point, bucket-local coordinate
  std::vector<uint8_t> pb;        // [D] bucket of the point
  int pbstart[NBMAX + 1] = {0};   // point index range per bucket
  double pt_min = 0, pt_max = 0;
  int minspan = 0;
  std::string error;
  // Offset quadrature (sclane_comp.cuh, "offset nodes"): the lineage's offsets o_i = log(1/sf_i) binned into K intervals of
  // width <= kOffBinWidth, each carrying kOffJ Chebyshev nodes.  Filled by build_offset_nodes() when a size factor is given.
  OffsetNodes ON;
  std::vector<double> oxi;        // [Npad] position of o_i inside its bin, in [-1, 1] (sorted-cell order)
  std::vector<uint8_t> obin;      // [Npad] bin of o_i
  std::vector<int32_t> oord;      // [N] sorted-cell positions ordered by (bin, position)
  std::vector<int32_t> ogrp;      // [K * (T + 2)] first entry of oord inside bin k whose position is >= bstart[b], b = 0 .. T + 1: the
                                  // cells of (offset bin k, bucket b) are oord[ogrp[k (T + 2) + b] .. ogrp[k (T + 2) + b + 1])
  // approx.knot = FALSE (marge2.R:167-173): every candidate knot is a data abscissa; the lineage carries ONE bucket (T = 0)
  // and every gene builds its own buckets from the knots it selects (sclane_exact_device.cuh)
  // scratch of build_lineage_plan(), kept with the plan so that a reused plan object (the library keeps a pool across calls) does
  // not re-allocate: at 200,000 cells the ~15 fresh 1.6 MB vectors cost more in page faults than the arithmetic
  std::vector<int32_t> ws_cells, ws_order, ws_start, ws_key;
  std::vector<double> ws_x, ws_X, ws_xs, ws_q;
  bool exact = false;
  std::vector<double> px;         // [D] abscissa (rounded pseudotime) of point d
  std::vector<int32_t> cand;      // [n_cand] point index of the candidate knots, ascending
};

// distinct abscissae of the sorted cells (a point never straddles a bucket: buckets are defined by the abscissa)
inline void build_points(LineagePlan& P) {
  const Lineage& L = P.L;
  P.pstart.clear(); P.pn.clear(); P.pu.clear();
  for (int b = 0; b <= NBMAX; ++b) P.pbstart[b] = 0;
  int b_cur = 0;
  for (int i = 0; i < L.N; ++i) {
    if (i == 0 || P.Xs[(size_t)i] != P.Xs[(size_t)i - 1]) {
      const int d = (int)P.pstart.size();
      while (b_cur <= L.T && i >= L.bstart[b_cur + 1]) { ++b_cur; P.pbstart[b_cur] = d; }
      P.pstart.push_back(i);
      P.pu.push_back(P.u[(size_t)i]);
    }
  }
  P.D = (int)P.pstart.size();
  P.pstart.push_back(L.N);
  for (int b = b_cur + 1; b <= NBMAX; ++b) P.pbstart[b] = P.D;
  P.pn.resize((size_t)P.D);
  P.pb.resize((size_t)P.D);
  for (int d = 0; d < P.D; ++d) P.pn[(size_t)d] = (double)(P.pstart[(size_t)d + 1] - P.pstart[(size_t)d]);
  for (int b = 0; b <= L.T; ++b) for (int d = P.pbstart[b]; d < P.pbstart[b + 1]; ++d) P.pb[(size_t)d] = (uint8_t)b;
}

// Offset quadrature plan (see sclane_comp.cuh): bins, per-cell Chebyshev coordinate, bin-ordered cell list, node abscissae
// and the gene-independent node weights n_kj = sum_i l_j(xi_i) over the cells of bin k.
inline void build_offset_nodes(LineagePlan& P) {
  OffsetNodes& O = P.ON;
  std::memset(&O, 0, sizeof(O));
  const int N = P.L.N;
  if (P.off.empty() || N < 1) return;
  double lo = P.off[0], hi = P.off[0], so = 0.0;
  for (int i = 0; i < N; ++i) { const double o = P.off[(size_t)i]; if (!(o == o) || std::isinf(o)) return; if (o < lo) lo = o; if (o > hi) hi = o; so += o; }
  const double R = hi - lo;
  int K = R > 0.0 ? (int)std::ceil(R / kOffBinWidth) : 1;
  if (K < 1) K = 1;
  if (K > kOffKMax) return;                                 // offsets spanning more than e^32: keep the per-cell fits
  const double w = R > 0.0 ? R / (double)K : 1.0;
  O.on = 1; O.K = K; O.omin = lo; O.w = w; O.sum_o = so;
  const double PI = 3.14159265358979323846;
  for (int m = 0; m < kOffJ; ++m)
    for (int j = 0; j < kOffJ; ++j) O.ctab[m][j] = std::cos((double)m * (2.0 * j + 1.0) * PI / (2.0 * kOffJ));
  for (int k = 0; k < K; ++k)
    for (int j = 0; j < kOffJ; ++j) O.node_o[k * kOffJ + j] = lo + ((double)k + 0.5) * w + 0.5 * w * O.ctab[1][j];
  P.oxi.assign((size_t)P.L.Npad, 0.0);
  P.obin.assign((size_t)P.L.Npad, 0);
  std::vector<int32_t> cnt((size_t)K + 1, 0);
  std::vector<double> cm((size_t)K * kOffJ, 0.0);            // Chebyshev moments sum_i T_m(xi_i) per bin
  for (int i = 0; i < N; ++i) {
    const double o = P.off[(size_t)i];
    int k = (int)std::floor((o - lo) / w);
    if (k < 0) k = 0;
    if (k > K - 1) k = K - 1;
    double xi = 2.0 * (o - (lo + ((double)k + 0.5) * w)) / w;
    if (xi < -1.0) xi = -1.0;
    if (xi > 1.0) xi = 1.0;
    P.oxi[(size_t)i] = xi;
    P.obin[(size_t)i] = (uint8_t)k;
    cnt[(size_t)k + 1]++;
    double t0 = 1.0, t1 = xi;
    cm[(size_t)k * kOffJ] += 1.0;
    cm[(size_t)k * kOffJ + 1] += xi;
    for (int m = 2; m < kOffJ; ++m) { const double t2 = 2.0 * xi * t1 - t0; cm[(size_t)k * kOffJ + m] += t2; t0 = t1; t1 = t2; }
  }
  for (int k = 0; k < K; ++k) cnt[(size_t)k + 1] += cnt[(size_t)k];
  for (int k = 0; k <= K; ++k) O.kstart[k] = cnt[(size_t)k];
  for (int k = K + 1; k <= kOffKMax; ++k) O.kstart[k] = N;
  P.oord.assign((size_t)N, 0);
  {
    std::vector<int32_t> pos(cnt.begin(), cnt.end() - 1);
    for (int i = 0; i < N; ++i) P.oord[(size_t)pos[P.obin[(size_t)i]]++] = i;
  }
  for (int k = 0; k < K; ++k) off_moments_to_nodes(O, &cm[(size_t)k * kOffJ], &O.node_n[k * kOffJ]);
  // (bin, bucket) groups: inside a bin the positions ascend, so a bucket is a contiguous range
  const int T = P.L.T;
  P.ogrp.assign((size_t)K * (size_t)(T + 2), 0);
  for (int k = 0; k < K; ++k) {
    const int32_t* lo_p = P.oord.data() + O.kstart[k];
    const int32_t* hi_p = P.oord.data() + O.kstart[k + 1];
    for (int b = 0; b <= T + 1; ++b)
      P.ogrp[(size_t)k * (size_t)(T + 2) + (size_t)b] = (int32_t)(std::lower_bound(lo_p, hi_p, (int32_t)P.L.bstart[b]) - P.oord.data());
  }
}

// pt_col: n_cells doubles with NaN = not in lineage.  size_factor may be null.
inline bool build_lineage_plan(const double* pt_col, int64_t n_cells, const double* size_factor,
                               const sclane_opts& o, LineagePlan& P) {
  std::vector<int32_t>& cells = P.ws_cells;
  std::vector<double>& x = P.ws_x;
  cells.clear(); x.clear();
  cells.reserve((size_t)n_cells); x.reserve((size_t)n_cells);
  P.error.clear();
  for (int64_t i = 0; i < n_cells; ++i)
    if (!std::isnan(pt_col[i])) { cells.push_back((int32_t)i); x.push_back(pt_col[i]); }
  const size_t N = x.size();
  if (N < 30) { P.error = "lineage has fewer than 30 cells"; return false; }
  // the 5 % / 95 % quantiles of the UNROUNDED pseudotime (marge2.R:158-159) only need x: selected on another thread while
  // this one rounds and sorts
  std::future<std::pair<double, double>> qfut;
  if (o.approx_knot != 0)
    qfut = std::async(N >= 32768 ? std::launch::async : std::launch::deferred,
                      [&x, &P]() { double a, b; quantile7_pair(x, 0.05, 0.95, a, b, &P.ws_q); return std::make_pair(a, b); });
  std::vector<double>& X = P.ws_X;
  X.resize(N);
  for (size_t i = 0; i < N; ++i) X[i] = r_round(x[i], 1e4);
  std::vector<int32_t>& order = P.ws_order;
  stable_order_by_x(X, order, P.ws_start, P.ws_key);
  P.Xs.resize(N);
  for (size_t i = 0; i < N; ++i) P.Xs[i] = X[(size_t)order[i]];
  P.ws_xs.assign(P.Xs.begin(), P.Xs.end());
  std::vector<double> knots = candidate_knots_sorted(x, P.ws_xs, o.approx_knot != 0, o.n_knot_max, o.minspan, &P.minspan,
                                                     qfut.valid() ? &qfut : nullptr);
  if (knots.empty()) { P.error = "no candidate knots (pseudotime has too few distinct values)"; return false; }
  std::sort(knots.begin(), knots.end());
  P.exact = o.approx_knot == 0;
  std::vector<double> cand_values;
  if (P.exact) { cand_values.swap(knots); knots.clear(); }        // the lineage itself has no knot: one bucket
  if ((int)knots.size() > TMAX) { P.error = "more than SCLANE_TMAX candidate knots"; return false; }
  Lineage& L = P.L;
  std::memset(&L, 0, sizeof(L));
  L.N = (int)N;
  L.Npad = (int)((N + 127) / 128 * 128);
  L.T = (int)knots.size();
  for (int k = 0; k < L.T; ++k) L.knots[k] = knots[(size_t)k];
  L.ref[0] = L.T > 0 ? L.knots[0] : P.Xs[0];                      // exact mode: reference = smallest abscissa
  for (int b = 1; b <= L.T; ++b) L.ref[b] = L.knots[b - 1];
  // the cells are sorted by X, so bucket b = #{k: t_k <= x} is the contiguous range starting at the first cell with X >= t_{b-1}
  L.bstart[0] = 0;
  for (int b = 1; b <= L.T; ++b) L.bstart[b] = (int)(std::lower_bound(P.Xs.begin(), P.Xs.end(), knots[(size_t)b - 1]) - P.Xs.begin());
  for (int b = L.T + 1; b <= NBMAX; ++b) L.bstart[b] = (int)N;
  P.perm.resize(N);
  P.bucket.resize(N);
  P.u.assign((size_t)L.Npad, 0.0);
  if (size_factor) P.off.assign((size_t)L.Npad, 0.0); else P.off.clear();
  parallel_chunks(N, [&](int, size_t a, size_t b2) {
    int b = (int)(std::upper_bound(L.bstart, L.bstart + L.T + 1, (int)a) - L.bstart) - 1;
    for (size_t i = a; i < b2; ++i) {
      while (b < L.T && (int)i >= L.bstart[b + 1]) ++b;
      const int32_t src = order[i];
      P.perm[i] = cells[(size_t)src];
      P.bucket[i] = b;
      P.u[i] = P.Xs[i] - L.ref[b];
      if (size_factor) P.off[i] = std::log(1.0 / size_factor[cells[(size_t)src]]);
    }
  });
  // unweighted bucket moments: one bucket per task, summed in cell order (bit-identical to a serial pass)
  {
    const int nb = L.T + 1;
    unsigned hw = host_threads();
    int nt = (int)(hw == 0 ? 4 : (hw > 8 ? 8 : hw));
    if (N < 32768) nt = 1;
    auto run = [&](int t) {
      for (int b = t; b < nb; b += nt) {
        double u0 = 0.0, u1 = 0.0, u2 = 0.0;
        for (int i = L.bstart[b]; i < L.bstart[b + 1]; ++i) { const double u = P.u[(size_t)i]; u0 += 1.0; u1 += u; u2 += u * u; }
        L.U0[b] = u0; L.U1[b] = u1; L.U2[b] = u2;
      }
    };
    if (nt == 1) run(0);
    else {
      std::vector<std::thread> th;
      for (int t = 1; t < nt; ++t) th.emplace_back(run, t);
      run(0);
      for (auto& x2 : th) x2.join();
    }
  }
  build_points(P);
  if (P.exact) {
    P.px.resize((size_t)P.D);
    for (int d = 0; d < P.D; ++d) P.px[(size_t)d] = P.Xs[(size_t)P.pstart[(size_t)d]];
    P.cand.clear();
    for (double v : cand_values) {
      const auto it = std::lower_bound(P.px.begin(), P.px.end(), v);
      if (it != P.px.end() && *it == v) P.cand.push_back((int32_t)(it - P.px.begin()));      // candidates are data values
    }
    if (P.cand.empty()) { P.error = "no candidate knots (pseudotime has too few distinct values)"; return false; }
  }
  build_offset_nodes(P);
  P.pt_min = *std::min_element(x.begin(), x.end());
  P.pt_max = *std::max_element(x.begin(), x.end());
  return true;
}

// ---- distribution functions ----------------------------------------------------------------------
// regularised upper incomplete gamma Q(a, x) (series / Lentz continued fraction)
inline double gamma_q(double a, double x) {
  if (std::isnan(a) || std::isnan(x)) return std::numeric_limits<double>::quiet_NaN();
  if (x <= 0.0) return 1.0;
  const double gln = std::lgamma(a);
  if (x < a + 1.0) {
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 10000; ++n) {
      ap += 1.0; del *= x / ap; sum += del;
      if (std::fabs(del) < std::fabs(sum) * 1e-17) break;
    }
    return 1.0 - sum * std::exp(-x + a * std::log(x) - gln);
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 10000; ++i) {
    const double an = -(double)i * ((double)i - a);
    b += 2.0;
    d = an * d + b; if (std::fabs(d) < tiny) d = tiny;
    c = b + an / c; if (std::fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < 1e-16) break;
  }
  return std::exp(-x + a * std::log(x) - gln) * h;
}
// stats::pchisq(q, df, lower.tail = FALSE)
inline double pchisq_upper(double q, double df) {
  if (std::isnan(q) || std::isnan(df)) return std::numeric_limits<double>::quiet_NaN();
  if (q <= 0.0) return 1.0;
  return gamma_q(0.5 * df, 0.5 * q);
}
// 2 * pnorm(-|z|)
inline double pnorm_two_sided(double z) {
  if (std::isnan(z)) return z;
  return std::erfc(std::fabs(z) / std::sqrt(2.0));
}

// ---- record assembly -----------------------------------------------------------------------------
// Raw per-gene device output
struct GeneOut {
  GeneState st;
  NbResult alt, null;
};

inline void finalize_record(const GeneOut& g, const Lineage& L, sclane_record& r, bool basis_only = false) {
  const double NaN = std::numeric_limits<double>::quiet_NaN();
  std::memset(&r, 0, sizeof(r));
  const GeneState& st = g.st;
  // basis_only (marge2(is.glmm = TRUE)): the record is OK when the backward pass selected a basis; nothing was fitted
  const bool basis_ok = (st.fin.n >= 1) && !(st.error & (SCLANE_NOTE_FWD_EMPTY | SCLANE_NOTE_NUMERIC | SCLANE_NOTE_ALL_ZERO | SCLANE_NOTE_BAD_COUNTS));
  const bool marge_ok = basis_only ? basis_ok : (basis_ok && g.alt.ok);
  const bool null_ok = !basis_only && g.null.ok && !(st.error & (SCLANE_NOTE_ALL_ZERO | SCLANE_NOTE_BAD_COUNTS));
  r.status = (marge_ok ? SCLANE_MARGE_OK : 0) | (null_ok ? SCLANE_NULL_OK : 0);
  r.marge_notes = st.error | ((marge_ok && !basis_only) ? g.alt.notes : 0);
  r.null_notes = null_ok ? g.null.notes : (st.error & (SCLANE_NOTE_ALL_ZERO | SCLANE_NOTE_BAD_COUNTS));
  for (int j = 0; j < SCLANE_PMAX; ++j) {
    r.term_kind[j] = 0; r.term_knot_idx[j] = -1;
    r.term_knot[j] = r.term_label[j] = r.coef[j] = r.se[j] = r.z[j] = r.p[j] = NaN;
  }
  for (int j = 0; j < SCLANE_PMAX * SCLANE_PMAX; ++j) r.cov[j] = NaN;
  r.theta = r.se_theta = r.loglik = r.deviance = NaN;
  r.null_coef = r.null_se = r.null_z = r.null_p = r.null_theta = r.null_se_theta = r.null_loglik = r.null_deviance = NaN;
  r.test_stat = r.df = r.p_val = NaN;
  r.gee_alpha = r.gee_phi = r.null_gee_alpha = r.null_gee_phi = NaN;
  if (marge_ok) {
    const int p = st.fin.n;
    r.n_terms = p;
    for (int j = 0; j < p; ++j) {
      r.term_kind[j] = st.fin.kind[j];
      r.term_knot_idx[j] = st.fin.kind[j] == SCLANE_TERM_INTERCEPT ? -1 : st.fin.knot[j];
      if (st.fin.kind[j] != SCLANE_TERM_INTERCEPT) {
        r.term_knot[j] = st.ex_T > 0 ? st.ex_knots[st.fin.knot[j]] : L.knots[st.fin.knot[j]];      // exact-knot mode: the gene's own knot table
        r.term_label[j] = r_signif(r.term_knot[j], 4);
      }
      if (basis_only) continue;
      r.coef[j] = g.alt.coef[j];
      r.se[j] = std::sqrt(g.alt.cov[j * PMAX + j]);
      r.z[j] = r.coef[j] / r.se[j];
      r.p[j] = pnorm_two_sided(r.z[j]);
      for (int l = 0; l < p; ++l) r.cov[j * p + l] = g.alt.cov[j * PMAX + l];
    }
    if (!basis_only) {
    r.theta = g.alt.theta; r.se_theta = g.alt.se_theta;
    r.loglik = g.alt.loglik; r.deviance = g.alt.deviance;
    r.alternations = g.alt.alternations; r.irls_iters = g.alt.irls_iters;
    }
  }
  if (null_ok) {
    r.null_coef = g.null.coef[0];
    r.null_se = std::sqrt(g.null.cov[0]);
    r.null_z = r.null_coef / r.null_se;
    r.null_p = pnorm_two_sided(r.null_z);
    r.null_theta = g.null.theta; r.null_se_theta = g.null.se_theta;
    r.null_loglik = g.null.loglik; r.null_deviance = g.null.deviance;
    r.null_alternations = g.null.alternations; r.null_irls_iters = g.null.irls_iters;
  }
  if (marge_ok && null_ok) {                      // modelLRT.R:43-53
    r.test_stat = 2.0 * (r.loglik - r.null_loglik);
    double df = (double)((r.n_terms + 1) - 2);    // logLik df = rank + 1 (theta) for both models
    if (df == 0.0) df = 1.0;
    r.df = df;
    r.p_val = pchisq_upper(r.test_stat, df);
  }
  r.fwd_n_terms = st.fwd.n;
  for (int j = 0; j < st.fwd.n && j < SCLANE_PMAX; ++j) { r.fwd_kind[j] = st.fwd.kind[j]; r.fwd_knot_idx[j] = st.fwd.kind[j] ? st.fwd.knot[j] : -1; }
  for (int j = 0; j < 4; ++j) { r.fwd_score[j] = j < st.fwd_iters ? st.fwd_score[j] : NaN; r.fwd_sigma[j] = j < st.fwd_iters ? st.fwd_sigma[j] : NaN; }
  r.n_wic = st.n_wic;
  for (int j = 0; j < SCLANE_PMAX; ++j) r.wic[j] = j < st.n_wic ? st.wic[j] : NaN;
  r.fit_passes = st.passes;
  for (int j = 0; j < 3; ++j) r.stage_cycles[j] = (double)st.cycles[j];
}

inline void init_gene_state(GeneState& st) {
  std::memset(&st, 0, sizeof(st));
  st.fwd.n = 1;
  st.fwd.kind[0] = SCLANE_TERM_INTERCEPT;
  st.fwd.knot[0] = 0;
  st.m_count = 1;
}

}  // namespace sclane
