"""Small run of every route (compressed GLM, per-cell GLM with offsets, routed gene, GEE ar1 / exchangeable + score test)
for compute-sanitizer."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sclane_b200 as sb
from tools import synth

counts, pt, _ = synth.make_counts(24, 4000, seed=3)
counts = np.ascontiguousarray(counts, dtype=np.int32)
counts[2, :] = 0
counts[3, 5:25] += 5000
counts[4, :] = 4200 + (np.arange(4000) % 500)
pt = np.asarray(pt, dtype=np.float64).reshape(-1)
sf = sb.createCellOffset(counts)
a = sb.testDynamic(counts, pt, n_gpus=1, preds="none")
b = sb.testDynamic(counts, pt, size_factor_offset=sf, n_gpus=1, preds="none")
c = sb.testDynamic(counts, pt, size_factor_offset=sf, n_gpus=1, preds="none", force_per_cell=True)
pt2 = np.stack([np.where(np.arange(4000) % 3 == 0, np.nan, pt), np.where(np.arange(4000) % 5 == 0, np.nan, pt)], 1)
d = sb.testDynamic(counts, pt2, size_factor_offset=sf, n_gpus=1, preds="eager")
subject = np.arange(4000) * 3 // 4000
for cor, test in (("ar1", "wald"), ("exchangeable", "score"), ("independence", "wald")):
    g = sb.testDynamic(counts[:8], pt, size_factor_offset=sf, is_gee=True, id_vec=subject, cor_structure=cor, gee_test=test, n_gpus=1, preds="none")
print("ok", sum(r.status == 3 for r in a.records), sum(r.status == 3 for r in b.records), sum(r.status == 3 for r in d.records), sb.launch_count())
