"""How far do the engine's two deterministic rules sit from the reference's third-party behaviour (oracle/flavours.py)?

1. Bundled golden (100 genes, the reference's own R run): term-set agreement of every (rank rule, fitter) combination.
2. Synthetic genes: disagreement rate between the exact-MLE oracle (what the engine implements) and the gamlss-RS emulation.
Prints a small report (profiles/r02_rule_flavours.txt)."""
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
EXC = {"ACP2", "ACP5", "ACTG1", "ACTR10", "ACVR1", "ACVR1B", "ADAMTSL2", "ADM", "ADRBK1", "AGO2", "AKAP13"}


def run(a):
    from threadpoolctl import threadpool_limits
    from oracle.flavours import flavour
    from oracle.testdynamic import test_dynamic
    counts, pt, genes, sf, rr, ft = a
    with threadpool_limits(limits=1), flavour(rank_rule=rr, fitter=ft):
        r = test_dynamic(counts, pt, genes, sf, keep_preds=False)
    return {g: (r[g]["Lineage_A"]["MARGE_Summary"]["term"] if r[g]["Lineage_A"]["MARGE_Summary"] else None, r[g]["Lineage_A"]["P_Val"]) for g in genes}


def run_parallel(counts, pt, genes, sf, rr, ft, cores):
    chunks = [(counts[i::cores], pt, genes[i::cores], sf, rr, ft) for i in range(cores)]
    out = {}
    with ProcessPoolExecutor(max_workers=cores) as ex:
        for part in ex.map(run, chunks):
            out.update(part)
    return out


def main():
    cores = min(os.cpu_count() or 1, 16)
    d = np.load(os.path.join(ROOT, "tests", "golden", "sim_inputs.npz"))
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "sclane_models.json")))
    counts = np.ascontiguousarray(d["counts"], dtype=np.int32)
    pt, genes = d["pt"], [str(g) for g in d["genes"]]
    sf = 1e4 / counts.sum(axis=0)
    print("== bundled golden: genes whose final term set equals the reference's (of 100; of the 11 stated exceptions)")
    for rr in ("structural", "colpiv_qr", "pair_unchecked"):
        for ft in ("exact_mle", "emulate_gamlss"):
            t = time.time()
            r = run_parallel(counts, pt, genes, sf, rr, ft, cores)
            ok = [g for g in genes if r[g][0] == gold[g]["MARGE_Summary"]["term"]]
            print(f"  rank_rule={rr:15s} fitter={ft:15s}: {len(ok):3d} / 100   exceptions recovered {len(EXC & set(ok)):2d} / 11   "
                  f"non-exception genes lost {len([g for g in genes if g not in EXC and g not in ok]):2d}   ({time.time() - t:.0f} s)")
    from tools import synth
    n_genes, n_cells = 1000, 1500
    counts, pt, _ = synth.make_counts(n_genes, n_cells, seed=77)
    genes = [f"g{i}" for i in range(n_genes)]
    a = run_parallel(counts, pt.reshape(-1), genes, None, "structural", "exact_mle", cores)
    b = run_parallel(counts, pt.reshape(-1), genes, None, "structural", "emulate_gamlss", cores)
    diff = [g for g in genes if a[g][0] != b[g][0]]
    pa = np.array([a[g][1] for g in genes], dtype=float)
    pb = np.array([b[g][1] for g in genes], dtype=float)
    from oracle.rmath import p_adjust_bh
    ca, cb = p_adjust_bh(pa) < 0.01, p_adjust_bh(pb) < 0.01
    print(f"== synthetic {n_genes} genes x {n_cells} cells (bench generator): exact MLE (engine) vs gamlss-RS emulation")
    print(f"  different final term set: {len(diff)} / {n_genes} ({100.0 * len(diff) / n_genes:.1f} %)")
    print(f"  different dynamic / static call at FDR 0.01: {int(np.sum(ca != cb))} / {n_genes}")


if __name__ == "__main__":
    main()
