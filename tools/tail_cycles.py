"""Where does a stage's time go -- throughput or tail?  Runs the default workload shape on a gene subset in both kernel
organisations and prints, per stage: kernel time, sum of per-gene cycles / resident units (the no-tail bound), the slowest gene."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import sclane_b200 as sb
from sclane_b200 import api
ap = argparse.ArgumentParser()
ap.add_argument("--genes", type=int, default=8000)
ap.add_argument("--cells", type=int, default=200000)
a = ap.parse_args()
dev = torch.device("cuda", 0)
d_counts, pt = bench.make_counts_gpu(0, a.genes, a.cells, dev)
clk = 1.965e9
for name, kw, units in (("warp per gene", {"comp_warp_per_gene": True}, 148 * 24), ("CTA per gene", {}, 148 * 3)):
    o = api._opts(gpu_ids=[0], **kw)
    for _ in range(2):
        recs, _ = api.run_records(a.genes, pt, None, o, device_ptr=d_counts.data_ptr(), device=0)
    p = api.get_profile()
    cyc = np.array([[r.stage_cycles[k] for k in range(3)] for r in recs])
    passes = np.array([r.fit_passes for r in recs])
    print(f"== {name}: genes {a.genes}, cells {a.cells}; resident units {units}")
    for k, st in enumerate(("comp_forward", "comp_backward", "comp_final")):
        c = cyc[:, k]
        print(f"  {st:14s} kernel {p['ms'][st]:8.2f} ms | sum(gene time)/units {c.sum() / units / clk * 1e3:8.2f} ms | slowest gene {c.max() / clk * 1e3:8.2f} ms | median {np.median(c) / clk * 1e3:7.3f} ms | p99 {np.quantile(c, 0.99) / clk * 1e3:7.3f} ms")
    print("  passes per gene: median", np.median(passes), "p99", np.quantile(passes, 0.99), "max", passes.max())
