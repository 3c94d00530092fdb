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
