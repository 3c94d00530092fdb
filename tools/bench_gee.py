"""GEE-mode measurement (BASELINE configs[2]: 5,000 genes x 30,000 cells, 3 subjects, AR-1 working correlation).

Not the driver's bench (bench.py measures the headline GLM metric); this prints one JSON line with the whole-call
throughput through the C ABI (host buffers, H2D inside the timed region) and the k_gee kernel time from CUDA events.
--order sorted: cells ordered by pseudotime inside each subject (the layout AR-1 is meant for; bundled data has it);
--order random: arbitrary cell order inside each subject (every 32-cell group then spans many knot buckets).
"""
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
    ap.add_argument("--genes", type=int, default=5000)
    ap.add_argument("--cells", type=int, default=30000)
    ap.add_argument("--subjects", type=int, default=3)
    ap.add_argument("--cor", default="ar1")
    ap.add_argument("--test", default="wald")
    ap.add_argument("--order", default="sorted", choices=["sorted", "random"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    args = ap.parse_args()
    import torch
    import bench
    import sclane_b200 as sb
    from sclane_b200 import _abi, api

    dev = torch.device("cuda", 0)
    d_counts, pt = bench.make_counts_gpu(0, args.genes, args.cells, dev)
    n = args.cells
    subject = (np.arange(n) * args.subjects // n).astype(np.int32)
    if args.order == "sorted":
        order = np.lexsort((pt, subject))
        d_counts = d_counts[:, torch.from_numpy(order).to(dev)].contiguous()
        pt = pt[order]
    counts = d_counts.cpu().numpy()
    del d_counts
    sf = api.createCellOffset(counts)
    o = api._opts(gpu_ids=[0])
    o.mode = 1
    o.gee_cor = _abi.COR_CODES[args.cor]
    o.gee_test = 1 if args.test == "score" else 0
    for _ in range(args.warmup):
        recs, _ = api.run_records(counts, pt, sf, o, subject_id=subject)
    sb.reset_launch_count()
    t0 = time.perf_counter()
    ms = {}
    for _ in range(args.steps):
        recs, _ = api.run_records(counts, pt, sf, o, subject_id=subject)
        p = sb.get_profile()
        for k, v in p["ms"].items():
            ms[k] = ms.get(k, 0.0) + v
    wall = (time.perf_counter() - t0) / args.steps
    ok = sum(1 for r in recs if r.status == 3)
    passes = float(np.mean([r.fit_passes for r in recs]))
    nterms = np.bincount([r.n_terms for r in recs], minlength=6).tolist()
    dyn = sum(1 for r in recs if r.status == 3 and r.p_val < 0.01)
    line = {"metric": "genes/sec for testDynamic GEE mode", "value": args.genes / wall, "unit": "genes/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall * 1e3, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic NB {args.genes} genes x {args.cells} cells, {args.subjects} subjects, GEE {args.cor}, "
                                   f"gee.test = {args.test}, cells {args.order} by pseudotime inside each subject (BASELINE configs[2])"},
            "timing": "host wall clock around sclane_test_dynamic_gee (host buffers: H2D + kernels + D2H)",
            "stage_ms_per_step": {k: v / args.steps for k, v in ms.items() if v},
            "gpu_launches": int(sb.launch_count()), "records_ok": ok, "mean_passes_per_gene": passes,
            "final_model_size_histogram": nterms, "dynamic_at_0.01_unadjusted": dyn}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
