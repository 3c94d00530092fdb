"""Debug helper: GEE mode on the GPU vs the host harness on the bundled data; prints every differing record."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sclane_b200 as sb  # noqa: E402
from sclane_b200 import _abi, api  # noqa: E402

d = np.load(os.path.join(ROOT, "tests/golden/sim_inputs.npz"), allow_pickle=True)
counts = np.ascontiguousarray(d["counts"], dtype=np.int32)
pt = np.ascontiguousarray(d["pt"], dtype=np.float64).reshape(-1)
genes = [str(g) for g in d["genes"]]
sf = 1e4 / counts.sum(axis=0).astype(np.float64)
sid = np.ascontiguousarray(np.unique(d["subject"], return_inverse=True)[1], dtype=np.int32)
lib = C.CDLL(os.path.join(ROOT, "tests/host_emu/libsclane_emu.so"))


def show(r):
    return (f"status {r.status} notes {r.marge_notes} terms {[(r.term_kind[j], round(r.term_knot[j], 4)) for j in range(r.n_terms)]} "
            f"coef {[r.coef[j] for j in range(r.n_terms)]} W {r.test_stat} alpha {r.gee_alpha} phi {r.gee_phi} it {r.gee_iters} "
            f"fwd {r.fwd_n_terms} {list(r.fwd_score)} sig {list(r.fwd_sigma)} wic {list(r.wic)[:r.n_wic]} passes {r.fit_passes}")


for cor in ("ar1", "exchangeable", "independence"):
    o = api._opts(gpu_ids=[0])
    o.mode = 1
    o.gee_cor = _abi.COR_CODES[cor]
    recs, _ = api.run_records(counts, pt[:, None], sf, o, subject_id=sid)
    eo = _abi.default_opts()
    eo.mode = 1
    eo.gee_cor = _abi.COR_CODES[cor]
    er = (_abi.SclaneRecord * len(genes))()
    ei = (_abi.SclaneLineageInfo * 1)()
    lib.emu_test_dynamic_gee(C.c_void_p(counts.ctypes.data), C.c_int64(counts.shape[1]), C.c_int64(len(genes)), C.c_void_p(pt.ctypes.data), 1,
                             C.c_void_p(sid.ctypes.data), C.c_void_p(sf.ctypes.data), C.byref(eo), er, ei)
    nbad = 0
    worst = 0.0
    for g, name in enumerate(genes):
        a, b = recs[g], er[g]
        same = a.status == b.status and a.n_terms == b.n_terms and all(a.term_kind[j] == b.term_kind[j] and a.term_knot_idx[j] == b.term_knot_idx[j] for j in range(a.n_terms))
        if same and a.status & 1:
            for j in range(a.n_terms):
                worst = max(worst, abs(a.coef[j] - b.coef[j]) / (abs(b.coef[j]) + 1e-6))
            worst = max(worst, abs(a.test_stat - b.test_stat) / (abs(b.test_stat) + 1e-6))
        if not same:
            nbad += 1
            print(cor, name, "\n  gpu:", show(a), "\n  emu:", show(b))
    print(cor, "mismatching records:", nbad, "worst rel diff:", worst, "profile ms:", {k: round(v, 2) for k, v in sb.get_profile()["ms"].items() if v})
