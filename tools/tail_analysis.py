"""Which genes make the launch tails of the final stage?  Prints fit passes per gene against the forward fit's sigma."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import sclane_b200 as sb
from sclane_b200 import api
d_counts, pt = bench.make_counts_gpu(0, 3000, 30000, torch.device("cuda", 0))
counts = d_counts.cpu().numpy()
recs, _ = api.run_records(counts, pt, None, api._opts(gpu_ids=[0]))
passes = np.array([r.fit_passes for r in recs]); alts = np.array([r.alternations + r.null_alternations for r in recs])
sig = np.array([[v for v in r.fwd_sigma] for r in recs]); last = np.array([s[~np.isnan(s)][-1] if (~np.isnan(s)).any() else np.nan for s in sig])
theta = np.array([r.theta for r in recs]); nth = np.array([r.null_theta for r in recs])
o = np.argsort(-passes)
print("top 15 by passes: passes, alternations, last fwd sigma, theta, null theta")
for i in o[:15]: print(passes[i], alts[i], last[i], theta[i], nth[i])
print("quantiles of passes", np.quantile(passes, [0.5, 0.9, 0.99, 1.0]))
for thr in (1e-3, 1e-2, 3e-2):
    m = last < thr
    print("sigma <", thr, ": n", m.sum(), "mean passes", passes[m].mean() if m.any() else None, "| others mean", passes[~m].mean(), "max", passes[~m].max())
from scipy.stats import spearmanr
nt = np.array([r.n_terms for r in recs]); mean_y = counts.mean(axis=1); fs = np.array([np.nanmax(np.array(list(r.fwd_score))) if np.isfinite(np.array(list(r.fwd_score))).any() else 0 for r in recs])
final_passes = np.array([r.alternations + r.null_alternations for r in recs]) * 1.0 + np.array([r.irls_iters + r.null_irls_iters for r in recs])
for name, key in (("last sigma", last), ("mean y", mean_y), ("n_terms", nt), ("fwd score", fs), ("sigma*mean", last * mean_y), ("log mean y", np.log(mean_y + 1e-9))):
    ok = np.isfinite(key)
    print(name, "spearman with passes", round(spearmanr(key[ok], passes[ok]).correlation, 3), "with alternations+irls", round(spearmanr(key[ok], final_passes[ok]).correlation, 3))
print("--- with offsets (per-cell final fits)")
sf = api.createCellOffset(counts)
recs2, _ = api.run_records(counts, pt, sf, api._opts(gpu_ids=[0]))
fin = np.array([r.alternations + r.null_alternations + r.irls_iters + r.null_irls_iters for r in recs2]) * 1.0
tot = np.array([r.fit_passes for r in recs2]) * 1.0
sig2 = np.array([[v for v in r.fwd_sigma] for r in recs2]); last2 = np.array([s[~np.isnan(s)][-1] if (~np.isnan(s)).any() else np.nan for s in sig2])
fs2 = np.array([np.nanmax(np.array(list(r.fwd_score))) if np.isfinite(np.array(list(r.fwd_score))).any() else 0 for r in recs2])
edges = [0, 1e-6, 1e-3, 1e-2, 5e-2, 0.2, 1.0, 1e9]
for a, b in zip(edges[:-1], edges[1:]):
    m = (last2 >= a) & (last2 < b)
    if m.any(): print(f"sigma [{a}, {b}): n {m.sum()} passes mean {tot[m].mean():.0f} p90 {np.quantile(tot[m], 0.9):.0f} max {tot[m].max():.0f}")
print("spearman passes~score", round(spearmanr(fs2, tot).correlation, 3), "passes~sigma", round(spearmanr(last2, tot).correlation, 3))
o2 = np.argsort(-tot)
for i in o2[:12]: print(int(tot[i]), recs2[i].alternations, recs2[i].null_alternations, last2[i], fs2[i], recs2[i].theta, recs2[i].null_theta)
