"""approx.knot = FALSE: GPU vs oracle on the parity-test inputs; prints the genes whose final term sets differ, with the traces."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import sclane_b200 as sb
from sclane_b200 import api
from tools import synth
from test_parity_r02 import oracle_parallel_opts
n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
n_genes = int(sys.argv[2]) if len(sys.argv) > 2 else 56
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 41
use_off = len(sys.argv) > 4 and sys.argv[4] == "off"
counts, pt, _ = synth.make_counts(n_genes, n_cells, seed=seed)
genes = [f"g{i}" for i in range(n_genes)]
sf = sb.createCellOffset(counts) if use_off else None
res = sb.testDynamic(counts, pt, genes=genes, size_factor_offset=sf, approx_knot=False, n_gpus=1, preds="none")
ref = oracle_parallel_opts(counts, pt.reshape(-1), genes, sf, approx_knot=False)
bad = 0
for i, g in enumerate(genes):
    r, o = res[g]["Lineage_A"], ref[g]["Lineage_A"]
    ta = r["MARGE_Summary"]["term"] if r["MARGE_Summary"] else None
    tb = o["MARGE_Summary"]["term"] if o["MARGE_Summary"] else None
    if ta == tb and r["Model_Status"] == o["Model_Status"]:
        continue
    bad += 1
    rec = res.records[i]
    tr = o["_trace"]
    print(g, "GPU", ta, "| oracle", tb)
    print("   GPU fwd", [(rec.fwd_kind[j], rec.fwd_knot_idx[j]) for j in range(rec.fwd_n_terms)], "scores", [round(v, 6) for v in rec.fwd_score[:2]], "wic", [round(v, 4) for v in rec.wic[:rec.n_wic]],
          "LRT", round(rec.test_stat, 4), "ll", round(rec.loglik, 4))
    print("   ORA fwd", [(t.kind, t.knot_idx) for t in tr.forward_terms], "scores", [round(float(s), 6) for s in tr.forward_scores], "wic", [round(float(v), 4) for v in tr.WIC_vec[1:]],
          "wald", [np.round(w, 4).tolist() for w in tr.wald_steps], "LRT", round(float(o["Test_Stat"]), 4), "ll", round(float(o["LogLik_MARGE"]), 4))
print("mismatches", bad, "of", len(genes), "| candidate knots", res.lineage_info[0]["n_knots"])
