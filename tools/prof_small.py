"""Small profiling driver: one testDynamic() call on synthetic data through the C ABI (for ncu / timing)."""
import argparse
import sys
import os
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sclane_b200 as sb
from tools import synth

ap = argparse.ArgumentParser()
ap.add_argument("--genes", type=int, default=296)
ap.add_argument("--cells", type=int, default=10000)
ap.add_argument("--reps", type=int, default=1)
args = ap.parse_args()
counts, pt, _ = synth.make_counts(args.genes, args.cells, seed=312)
opts = sb.api._opts(gpu_ids=[0])
for r in range(args.reps):
    t0 = time.time()
    recs, infos = sb.run_records(counts, pt, None, opts)
    dt = time.time() - t0
p = sb.get_profile()
passes = np.array([recs[i].fit_passes for i in range(args.genes)])
alts = np.array([recs[i].alternations for i in range(args.genes)])
irls = np.array([recs[i].irls_iters for i in range(args.genes)])
print(f"genes {args.genes} cells {args.cells} wall {dt*1e3:.1f} ms; stage ms {p['ms']}")
print(f"passes/gene mean {passes.mean():.1f} median {np.median(passes):.0f} max {passes.max()}; alternations mean {alts.mean():.1f} max {alts.max()}; irls mean {irls.mean():.1f}")
print("status counts", np.bincount([recs[i].status for i in range(args.genes)], minlength=4), "fwd terms", np.bincount([recs[i].fwd_n_terms for i in range(args.genes)], minlength=6))
