"""C4-shaped run (2 lineages, offsets): per-gene cycles of the final stage (per-cell alternative fits) vs the kernel time."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import sclane_b200 as sb
from sclane_b200 import api
ap = argparse.ArgumentParser()
ap.add_argument("--genes", type=int, default=4000)
ap.add_argument("--cells", type=int, default=50000)
a = ap.parse_args()
dev = torch.device("cuda", 0)
d_counts, pt = bench.make_counts_gpu(0, a.genes, a.cells, dev)
counts = d_counts.cpu().numpy(); del d_counts
n = a.cells
rng = np.random.default_rng(312); v = rng.random(n); in_a, in_b = v < 0.6, v > 0.4
def rank01(x, mask):
    out = np.full(n, np.nan); idx = np.flatnonzero(mask); order = np.argsort(x[idx], kind="stable")
    r = np.empty(idx.size); r[order] = (np.arange(idx.size) + 0.5) / idx.size; out[idx] = r; return out
pt2 = np.stack([rank01(pt, in_a), rank01(1.0 - pt, in_b)], 1)
sf = api.createCellOffset(counts)
clk = 1.965e9
o = api._opts(gpu_ids=[0])
for _ in range(2):
    recs, _ = api.run_records(counts, pt2, sf, o)
p = api.get_profile()
print({k: round(v, 2) for k, v in p["ms"].items() if v > 0.01})
cyc = np.array([r.stage_cycles[2] for r in recs]); nt = np.array([r.n_terms for r in recs]); passes = np.array([r.fit_passes for r in recs])
alt = np.array([r.alternations for r in recs]); irls = np.array([r.irls_iters for r in recs])
h = nt > 1
print("records", len(recs), "with hinges", h.sum(), "final-stage kernel ms (2 lineages)", round(p["ms"]["final"], 2))
print("sum(cycles)/444 units: %.2f ms | slowest %.2f ms | median(hinge genes) %.3f ms | p90 %.3f | p99 %.3f" % (cyc.sum() / 444 / clk * 1e3, cyc.max() / clk * 1e3, np.median(cyc[h]) / clk * 1e3, np.quantile(cyc[h], 0.9) / clk * 1e3, np.quantile(cyc[h], 0.99) / clk * 1e3))
print("alt fits: alternations median %d p90 %d max %d; irls iters median %d p90 %d max %d" % (np.median(alt[h]), np.quantile(alt[h], 0.9), alt[h].max(), np.median(irls[h]), np.quantile(irls[h], 0.9), irls[h].max()))
o2 = np.argsort(-cyc)
for i in o2[:8]: print("  gene-lineage", i, "ms %.2f" % (cyc[i] / clk * 1e3), "alternations", alt[i], "irls", irls[i], "theta %.4g" % recs[i].theta, "n_terms", nt[i], "fwd sigma", [round(x, 4) for x in recs[i].fwd_sigma[:2]])
ms = cyc / clk * 1e3
for thr in (2, 5, 10, 20, 40):
    sel = ms > thr
    print("  genes slower than %2d ms: %5d  (share of all CTA time %.3f)" % (thr, sel.sum(), ms[sel].sum() / ms.sum()))
for lo, hi in ((1, 2), (3, 5), (6, 10), (11, 24), (25, 25)):
    sel = h & (alt >= lo) & (alt <= hi)
    if sel.any():
        print("  alternations %2d-%2d: %5d genes, median %.2f ms, max %.2f ms, share of CTA time %.3f" % (lo, hi, sel.sum(), np.median(ms[sel]), ms[sel].max(), ms[sel].sum() / ms.sum()))
