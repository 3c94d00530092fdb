"""BASELINE configs[3]: synthetic 10,000 genes x 50,000 cells, 2 lineages (60 % of the cells each, 20 % shared), GLM mode
with cell-size offsets (createCellOffset).  Prints one JSON line: end-to-end genes/s through the C ABI (host buffers) and the
per-stage CUDA-event times.  With offsets the forward / backward fits run on the X-compressed kernels and the final + null
glm.nb on the per-cell kernels."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes", type=int, default=10000)
    ap.add_argument("--cells", type=int, default=50000)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--h2d-mode", type=int, default=0)
    ap.add_argument("--genes-per-batch", type=int, default=0)
    args = ap.parse_args()
    import torch
    import bench
    import sclane_b200 as sb
    from sclane_b200 import api

    dev = torch.device("cuda", 0)
    d_counts, pt = bench.make_counts_gpu(0, args.genes, args.cells, dev)
    counts = d_counts.cpu().numpy()
    del d_counts
    n = args.cells
    rng = np.random.default_rng(312)
    v = rng.random(n)
    in_a, in_b = v < 0.6, v > 0.4

    def rank01(x, mask):
        out = np.full(n, np.nan)
        idx = np.flatnonzero(mask)
        order = np.argsort(x[idx], kind="stable")
        r = np.empty(idx.size)
        r[order] = (np.arange(idx.size) + 0.5) / idx.size
        out[idx] = r
        return out

    pt2 = np.stack([rank01(pt, in_a), rank01(1.0 - pt, in_b)], 1)
    sf = api.createCellOffset(counts)
    o = api._opts(gpu_ids=[0], h2d_mode=args.h2d_mode, genes_per_batch=args.genes_per_batch)
    for _ in range(args.warmup):
        recs, _ = api.run_records(counts, pt2, sf, o)
    sb.reset_launch_count()
    ms = {}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        recs, _ = api.run_records(counts, pt2, sf, o)
        for k, val in sb.get_profile()["ms"].items():
            ms[k] = ms.get(k, 0.0) + val
    wall = (time.perf_counter() - t0) / args.steps
    ok = sum(1 for r in recs if r.status == 3)
    line = {"metric": "genes/sec for testDynamic GLM mode", "value": args.genes / wall, "unit": "genes/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall * 1e3, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic NB {args.genes} genes x {args.cells} cells, 2 lineages (60 % of the cells each), GLM with "
                                   "size-factor offsets (BASELINE configs[3])"},
            "timing": "host wall clock around sclane_test_dynamic (host buffers: H2D + kernels + D2H); a gene counts once both lineages are done",
            "stage_ms_per_step": {k: val / args.steps for k, val in ms.items() if val},
            "gpu_launches": int(sb.launch_count()), "gene_lineage_records_ok": ok, "records": len(recs)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
