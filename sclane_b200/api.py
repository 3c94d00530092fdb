"""Host-side mirror of the reference's R interface for the per-gene fitting path.

Same names, argument meaning and error behaviour as scLANE's R functions (R is not available in
this image, so the host side above the C ABI is Python; the R shim a maintainer would use is in
r/ and INTEGRATION.md):

    testDynamic()       R/testDynamic.R:60-448      -> sclane_test_dynamic()
    marge2()            R/marge2.R:56-889           -> sclane_test_dynamic() on one gene
    modelLRT()          R/modelLRT.R:12-63
    getResultsDE()      R/getResultsDE.R:23-45
    testSlope()         R/testSlope.R:21-39
    createCellOffset()  R/createCellOffset.R:19-38
    eigenMapMatMult / eigenMapMatrixInvert / eigenMapPseudoInverse   R/RcppExports.R:4-14

All numerics run in libsclane_b200.so on the GPU; this module only marshals buffers and rebuilds
the 21-field result list of testDynamic.R:320-342 from the fixed-size records.
"""
from __future__ import annotations

import ctypes as C
import math
import time
from typing import Any

import numpy as np

from . import _abi
from ._lib import SclaneError, check, lib

LETTERS = "ABCDEFGHIJKLMNOPQRSTUVWXYZ"
_ERRLEN = 1024


# ------------------------------------------------------------------------------------------------
# small R semantics needed on the host side
# ------------------------------------------------------------------------------------------------
def _fmt_num(x: float) -> str:
    """as.character(<double>): 15 significant digits (paste() at marge2.R:593-594)."""
    return f"{x:.15g}"


def _p_adjust(p: np.ndarray, method: str = "fdr") -> np.ndarray:
    """stats::p.adjust(p, method) for every method of stats::p.adjust.methods (getResultsDE.R:25,39; testSlope.R:24,31);
    NA stays NA and is not counted in n."""
    if method not in ("holm", "hochberg", "hommel", "bonferroni", "BH", "BY", "fdr", "none"):
        raise ValueError("Please choose a valid p-value adjustment method.")
    p = np.asarray(p, dtype=np.float64)
    out = np.full(p.shape, np.nan)
    ok = ~np.isnan(p)
    pv = p[ok]
    n = pv.size
    if n == 0 or method == "none":
        out[ok] = pv
        return out
    if method == "bonferroni":
        adj = np.minimum(1.0, n * pv)
    elif method == "holm":
        o = np.argsort(pv, kind="stable")
        ro = np.argsort(o, kind="stable")
        adj = np.minimum(1.0, np.maximum.accumulate((n - np.arange(n)) * pv[o]))[ro]
    elif method == "hochberg":
        o = np.argsort(-pv, kind="stable")
        ro = np.argsort(o, kind="stable")
        i = np.arange(n, 0, -1, dtype=np.float64)
        adj = np.minimum(1.0, np.minimum.accumulate((n - i + 1.0) * pv[o]))[ro]
    elif method in ("BH", "fdr", "BY"):
        o = np.argsort(-pv, kind="stable")
        ro = np.argsort(o, kind="stable")
        i = np.arange(n, 0, -1, dtype=np.float64)
        q = float(np.sum(1.0 / np.arange(1, n + 1))) if method == "BY" else 1.0
        adj = np.minimum(1.0, np.minimum.accumulate(q * n / i * pv[o]))[ro]
    else:   # hommel (stats::p.adjust source, restated)
        if n > 1:
            o = np.argsort(pv, kind="stable")
            ps = pv[o]
            ro = np.argsort(o, kind="stable")
            i = np.arange(1, n + 1, dtype=np.float64)
            q = pa = np.full(n, np.min(n * ps / i))
            for m in range(n - 1, 1, -1):
                i1 = np.arange(0, n - m + 1)
                i2 = np.arange(n - m + 1, n)
                q1 = np.min(m * ps[i2] / np.arange(2, m + 1))
                q = q.copy()
                q[i1] = np.minimum(m * ps[i1], q1)
                q[i2] = q[n - m]
                pa = np.maximum(pa, q)
            adj = np.maximum(pa, ps)[ro]
        else:
            adj = pv.copy()
    out[ok] = adj
    return out


def _p_adjust_bh(p: np.ndarray) -> np.ndarray:
    return _p_adjust(p, "fdr")


def _na(x):
    return None if (isinstance(x, float) and math.isnan(x)) else x


# ------------------------------------------------------------------------------------------------
# results container
# ------------------------------------------------------------------------------------------------
class ScLANE(dict):
    """Result of testDynamic(): {gene: {"Lineage_A": {...21 fields...}, ...}} (class "scLANE", testDynamic.R:446)."""

    lineage_info: list
    profile: dict
    records: Any


class LazyPreds:
    """MARGE_Preds / Null_Preds (testDynamic.R:339-340) computed on first use by sclane_link_preds():
    the per-cell link fit and its standard error are 2 x 2 doubles per cell per gene (128 GB at
    20k genes x 200k cells), so they are materialised on demand instead of eagerly."""

    def __init__(self, record, x, size_factor, which, device=0):
        self._rec, self._x, self._sf, self._which, self._dev = record, x, size_factor, which, device
        self._val = None

    def compute(self) -> dict:
        if self._val is None:
            f1, s1, f0, s0 = link_preds(self._rec, self._x, self._sf, self._dev)
            self._val = ({"marge_link_fit": f1, "marge_link_se": s1} if self._which == "marge"
                         else {"null_link_fit": f0, "null_link_se": s0})
        return self._val

    def __getitem__(self, k):
        return self.compute()[k]

    def keys(self):
        return self.compute().keys()


# ------------------------------------------------------------------------------------------------
# low-level calls
# ------------------------------------------------------------------------------------------------
def device_count() -> int:
    return int(lib().sclane_device_count())


def launch_count() -> int:
    return int(lib().sclane_launch_count())


def reset_launch_count() -> None:
    lib().sclane_reset_launch_count()


def release() -> None:
    """Free the per-device work buffers the library keeps between calls."""
    lib().sclane_release()


class _Profile(C.Structure):
    _fields_ = [("ms", C.c_double * 15), ("launches", C.c_int64 * 15), ("sweep_bytes", C.c_double),
                ("sweep_gene_iters", C.c_int64), ("h2d_bytes", C.c_double), ("d2h_bytes", C.c_double),
                ("wall_ms", C.c_double)]


STAGE_NAMES = ["gather", "fwd_fit", "sweep", "backward", "final", "copy", "select", "gee", "aggregate", "comp_forward",
               "comp_backward", "comp_final", "off_aggregate", "off_null", "plan"]


def get_profile() -> dict:
    p = _Profile()
    lib().sclane_get_profile(C.byref(p))
    return {"ms": {n: p.ms[i] for i, n in enumerate(STAGE_NAMES)},
            "launches": {n: int(p.launches[i]) for i, n in enumerate(STAGE_NAMES)},
            "sweep_bytes": p.sweep_bytes, "sweep_gene_iters": int(p.sweep_gene_iters),
            "h2d_bytes": p.h2d_bytes, "d2h_bytes": p.d2h_bytes, "wall_ms": p.wall_ms}


def _opts(n_potential_basis_fns=5, approx_knot=True, n_knot_max=25, minspan=None, tols_score=1e-5,
          n_gpus=0, gpu_ids=None, genes_per_batch=0, verbose=False, force_per_cell=False, h2d_mode=0,
          comp_warp_per_gene=False, offset_per_cell=False, host_plan=False, batch_streams=0) -> _abi.SclaneOpts:
    o = _abi.default_opts()
    o.M = int(n_potential_basis_fns)
    o.approx_knot = 1 if approx_knot else 0
    o.n_knot_max = int(n_knot_max)
    o.minspan = -1 if minspan is None else int(minspan)
    o.tols_score = float(tols_score)
    if gpu_ids is not None:
        o.n_gpus = len(gpu_ids)
        for i, d in enumerate(gpu_ids):
            o.gpu_ids[i] = int(d)
    elif n_gpus:
        o.n_gpus = int(n_gpus)
        for i in range(int(n_gpus)):
            o.gpu_ids[i] = i
    o.genes_per_batch = int(genes_per_batch)
    o.verbose = 1 if verbose else 0
    o.force_per_cell = 1 if force_per_cell else 0
    o.h2d_mode = int(h2d_mode)
    o.comp_warp_per_gene = 1 if comp_warp_per_gene else 0
    o.offset_per_cell = 1 if offset_per_cell else 0
    o.host_plan = 1 if host_plan else 0
    o.batch_streams = int(batch_streams)
    return o


def lineage_plan(pt, opts: _abi.SclaneOpts | None = None, on_device: bool = False, device: int = 0) -> dict:
    """sclane_lineage_plan: K0 (sort by round(x, 4), candidate knots of marge2.R:150-166, bucket and abscissa tables) for one
    pseudotime column (NaN = not in the lineage), built on the host or by the device kernels of the batch path."""
    x = np.ascontiguousarray(pt, dtype=np.float64).reshape(-1)
    n = x.shape[0]
    if opts is None:
        opts = _opts()
    T2 = _abi.SCLANE_TMAX + 2
    perm = np.zeros(n + 1, dtype=np.int32); u = np.zeros(n + 1); knots = np.zeros(T2); bstart = np.zeros(T2, dtype=np.int32)
    U = np.zeros(3 * (T2 - 1)); pstart = np.zeros(n + 1, dtype=np.int32); pn = np.zeros(n + 1); pu = np.zeros(n + 1)
    pb = np.zeros(n + 1, dtype=np.uint8); pbstart = np.zeros(T2, dtype=np.int32); dims = np.zeros(4, dtype=np.int32)
    err = C.create_string_buffer(_ERRLEN)
    rc = lib().sclane_lineage_plan(x.ctypes.data, n, C.byref(opts), 1 if on_device else 0, int(device), perm.ctypes.data, u.ctypes.data,
                                   knots.ctypes.data, bstart.ctypes.data, U.ctypes.data, pstart.ctypes.data, pn.ctypes.data, pu.ctypes.data,
                                   pb.ctypes.data, pbstart.ctypes.data, dims.ctypes.data, err, _ERRLEN)
    if rc != 0:
        raise SclaneError(rc, err.value.decode(errors="replace"))
    N, T, D, minspan = (int(v) for v in dims)
    return {"N": N, "T": T, "D": D, "minspan": minspan, "perm": perm[:N], "u": u[:N], "knots": knots[:T], "bstart": bstart[:T + 2],
            "U": U[:3 * (T + 1)].reshape(3, T + 1), "pstart": pstart[:D + 1], "pn": pn[:D], "pu": pu[:D], "pb": pb[:D], "pbstart": pbstart[:T + 2]}


def _alloc_records(n: int):
    """Record array for the library to fill: uninitialised memory (the library writes every byte of every record; a zeroed
    ctypes array of 20,000 records costs ~8 ms of page faults and memset on the calling thread before the call even starts)."""
    buf = np.empty(max(1, n) * C.sizeof(_abi.SclaneRecord), dtype=np.uint8)
    return (_abi.SclaneRecord * n).from_buffer(buf)


def _is_sparse(m) -> bool:
    return hasattr(m, "tocsc") and hasattr(m, "nnz")


def run_records(counts, pt, size_factor=None, opts: _abi.SclaneOpts | None = None, device_ptr: int | None = None,
                device: int = 0, subject_id=None, basis_only: bool = False):
    """Thin call of sclane_test_dynamic[_device|_gee]: counts int32 [genes, cells] C-contiguous (gene-major), pt float64
    [cells, lineages]; subject_id (GEE mode, opts.mode = 1): int32 per cell, cells ordered by subject.
    Returns (records ctypes array [genes * lineages], lineage infos)."""
    L = lib()
    pt = np.asarray(pt, dtype=np.float64)
    if pt.ndim == 1:
        pt = pt[:, None]
    n_cells, n_lin = pt.shape
    pt_cm = np.ascontiguousarray(pt.T)          # column-major cells x lineages
    if opts is None:
        opts = _opts()
    sf = None if size_factor is None else np.ascontiguousarray(size_factor, dtype=np.float64)
    if sf is not None and sf.shape != (n_cells,):
        raise ValueError("size.factor.offset must have one entry per cell")
    err = C.create_string_buffer(_ERRLEN)
    if device_ptr is None and _is_sparse(counts):
        # genes x cells sparse matrix (dgCMatrix layout = scipy CSC): sclane_test_dynamic_csc
        csc = counts.tocsc()
        if csc.shape[1] != n_cells:
            raise ValueError("expr.mat must be genes x cells with one column per row of pt")
        if subject_id is not None:
            raise SclaneError(4, "sparse input is built for GLM mode only")
        csc.sort_indices()
        n_genes = csc.shape[0]
        if int(csc.indptr[-1]) > np.iinfo(np.int32).max:
            raise ValueError("sparse expr.mat has more than 2^31 - 1 non-zeros: split the genes into several calls")
        indptr = np.ascontiguousarray(csc.indptr, dtype=np.int32)
        indices = np.ascontiguousarray(csc.indices, dtype=np.int32)
        data = np.ascontiguousarray(csc.data, dtype=np.float64)
        recs = _alloc_records(n_genes * n_lin)
        infos = (_abi.SclaneLineageInfo * n_lin)()
        rc = L.sclane_test_dynamic_csc(indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, n_cells, n_genes, pt_cm.ctypes.data,
                                       n_lin, None if sf is None else sf.ctypes.data, C.byref(opts), recs, infos, err, _ERRLEN)
    elif device_ptr is None:
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        if counts.ndim != 2 or counts.shape[1] != n_cells:
            raise ValueError("expr.mat must be genes x cells with one column per row of pt")
        n_genes = counts.shape[0]
        recs = _alloc_records(n_genes * n_lin)
        infos = (_abi.SclaneLineageInfo * n_lin)()
        if subject_id is not None:
            sid = np.ascontiguousarray(subject_id, dtype=np.int32)
            if sid.shape != (n_cells,):
                raise ValueError("id.vec must have one entry per cell")
            rc = L.sclane_test_dynamic_gee(counts.ctypes.data, n_cells, n_genes, pt_cm.ctypes.data, n_lin, sid.ctypes.data,
                                           None if sf is None else sf.ctypes.data, C.byref(opts), recs, infos, err, _ERRLEN)
        elif basis_only:
            rc = L.sclane_marge_basis(counts.ctypes.data, n_cells, n_genes, pt_cm.ctypes.data, n_lin, C.byref(opts), recs, infos, err, _ERRLEN)
        else:
            rc = L.sclane_test_dynamic(counts.ctypes.data, n_cells, n_genes, pt_cm.ctypes.data, n_lin,
                                       None if sf is None else sf.ctypes.data, C.byref(opts), recs, infos, err, _ERRLEN)
    else:
        n_genes = int(counts)  # when a device pointer is given, `counts` is the number of genes
        recs = _alloc_records(n_genes * n_lin)
        infos = (_abi.SclaneLineageInfo * n_lin)()
        rc = L.sclane_test_dynamic_device(C.c_void_p(device_ptr), device, n_cells, n_genes, pt_cm.ctypes.data, n_lin,
                                          None if sf is None else sf.ctypes.data, C.byref(opts), recs, infos, err, _ERRLEN)
    check(rc, err)
    return recs, infos


def link_preds(record, x, size_factor=None, device: int = 0):
    """predict(type = "link", se.fit = TRUE) for both models of one record (pullMARGESummary.R:66-69)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.size
    sf = None if size_factor is None else np.ascontiguousarray(size_factor, dtype=np.float64)
    outs = [np.empty(n) for _ in range(4)]
    err = C.create_string_buffer(_ERRLEN)
    rc = lib().sclane_link_preds(C.byref(record), x.ctypes.data, n, None if sf is None else sf.ctypes.data, device,
                                 *[o.ctypes.data for o in outs], err, _ERRLEN)
    check(rc, err)
    return tuple(outs)


def candidate_knots(x, approx_knot=True, n_knot_max=25, minspan=None):
    """marge2.R:150-173 (host-side; no GPU needed)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    o = _opts(approx_knot=approx_knot, n_knot_max=n_knot_max, minspan=minspan)
    out = np.empty(_abi.SCLANE_TMAX)
    ms = C.c_int32(0)
    n = lib().sclane_candidate_knots(x.ctypes.data, x.size, C.byref(o), out.ctypes.data, C.byref(ms))
    if n < 0:
        raise SclaneError(-n, "sclane_candidate_knots failed")
    return out[:n].copy(), int(ms.value)


def stat_out_score_glm_null(counts, x, knots, term_kind, term_knot_idx, device: int = 0, want_mu: bool = True):
    """stat_out_score_glm_null.R:18-48 for a batch of genes sharing one basis.  Returns (beta [G, p], sigma [G], mu [G, N])."""
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    G, N = counts.shape
    x = np.ascontiguousarray(x, dtype=np.float64)
    knots = np.ascontiguousarray(knots, dtype=np.float64)
    tk = np.ascontiguousarray(term_kind, dtype=np.int32)
    ti = np.ascontiguousarray(term_knot_idx, dtype=np.int32)
    p = tk.size
    beta = np.empty((G, p))
    sigma = np.empty(G)
    mu = np.empty((G, N)) if want_mu else None
    err = C.create_string_buffer(_ERRLEN)
    rc = lib().sclane_null_stats_glm(counts.ctypes.data, N, G, x.ctypes.data, knots.ctypes.data, knots.size, tk.ctypes.data,
                                     ti.ctypes.data, p, device, beta.ctypes.data, sigma.ctypes.data,
                                     None if mu is None else mu.ctypes.data, err, _ERRLEN)
    check(rc, err)
    return beta, sigma, mu


def score_fun_glm_sweep(counts, mu, sigma, x, knots, term_kind, term_knot_idx, device: int = 0):
    """score_fun_glm.R:21-54 for every candidate knot and both candidate shapes.  Returns scores [G, 2, T]
    ([:, 0] = tp1 only, [:, 1] = hinge pair); NaN = NA."""
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    G, N = counts.shape
    x = np.ascontiguousarray(x, dtype=np.float64)
    knots = np.ascontiguousarray(knots, dtype=np.float64)
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    tk = np.ascontiguousarray(term_kind, dtype=np.int32)
    ti = np.ascontiguousarray(term_knot_idx, dtype=np.int32)
    out = np.empty((G, 2, knots.size))
    err = C.create_string_buffer(_ERRLEN)
    rc = lib().sclane_score_glm(counts.ctypes.data, mu.ctypes.data, N, G, x.ctypes.data, knots.ctypes.data, knots.size,
                                tk.ctypes.data, ti.ctypes.data, tk.size, sigma.ctypes.data, device, out.ctypes.data, err, _ERRLEN)
    check(rc, err)
    return out


# ---- drop-ins for the RcppEigen helpers (R/RcppExports.R:4-14) -------------------------------------
def eigenMapMatMult(A, B, n_cores: int = 1, device: int = 0):
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    if A.ndim != 2 or B.ndim != 2 or A.shape[1] != B.shape[0]:
        raise ValueError("non-conformable arguments")
    m, k = A.shape
    n = B.shape[1]
    Af, Bf = np.asfortranarray(A), np.asfortranarray(B)
    Cm = np.empty((m, n), order="F")
    err = C.create_string_buffer(_ERRLEN)
    check(lib().sclane_matmul(Af.ctypes.data, Bf.ctypes.data, Cm.ctypes.data, m, k, n, device, err, _ERRLEN), err)
    return Cm


def eigenMapMatrixInvert(A, n_cores: int = 1, device: int = 0):
    A = np.asfortranarray(np.asarray(A, dtype=np.float64))
    n = A.shape[0]
    out = np.empty((n, n), order="F")
    err = C.create_string_buffer(_ERRLEN)
    check(lib().sclane_llt_inverse(A.ctypes.data, out.ctypes.data, n, device, err, _ERRLEN), err)
    return out


def eigenMapPseudoInverse(A, tolerance: float = 1e-6, n_cores: int = 1, device: int = 0):
    A = np.asfortranarray(np.asarray(A, dtype=np.float64))
    m, n = A.shape
    out = np.empty((n, m), order="F")
    err = C.create_string_buffer(_ERRLEN)
    check(lib().sclane_pinv(A.ctypes.data, out.ctypes.data, m, n, float(tolerance), device, err, _ERRLEN), err)
    return out


# ------------------------------------------------------------------------------------------------
# record -> the reference's list fields
# ------------------------------------------------------------------------------------------------
def _term_names(rec, lineage_letter: str) -> list[str]:
    """marge2.R:593-594 + :817-821: Intercept, h_<Lineage>_<knot> (tp1) / h_<knot>_<Lineage> (tp2)."""
    var = f"Lineage_{lineage_letter}"
    names = ["Intercept"]
    for j in range(1, rec.n_terms):
        lab = _fmt_num(rec.term_label[j])
        names.append(f"h_{var}_{lab}" if rec.term_kind[j] == _abi.TERM_TP1 else f"h_{lab}_{var}")
    return names


def _status_string(rec) -> str:
    """testDynamic.R:266-278."""
    m_ok = bool(rec.status & _abi.SCLANE_MARGE_OK)
    n_ok = bool(rec.status & _abi.SCLANE_NULL_OK)
    return f"MARGE model {'OK' if m_ok else 'error'}, null model {'OK' if n_ok else 'error'}"


def _notes(bits: int, ok: bool) -> str | None:
    if ok:
        return None
    msgs = []
    if bits & _abi.NOTE_FWD_EMPTY:
        msgs.append("Error in wic_mat_2[...] : subscript out of bounds (forward pass added fewer than two basis functions)")
    if bits & _abi.NOTE_ALL_ZERO:
        msgs.append("all counts are zero")
    if bits & _abi.NOTE_BAD_COUNTS:
        msgs.append("negative values not allowed for the 'Negative Binomial' family")
    if bits & _abi.NOTE_NUMERIC:
        msgs.append("missing value where TRUE/FALSE needed")
    return "; ".join(msgs) if msgs else "model fit failed"


def _slope_data(rec, names, pt_lineage, gene, letter):
    """createSlopeTestData.R:15-74 + testDynamic.R:295-298."""
    m_ok = bool(rec.status & _abi.SCLANE_MARGE_OK)
    if not m_ok or rec.n_terms == 1:
        note = "MARGE model error" if not m_ok else "No non-intercept coefficients"
        return {"Gene": [gene], "Lineage": [letter], "Breakpoint": [None], "Rounded_Breakpoint": [None], "Direction": [None],
                "Beta": [None], "Test_Stat": [None], "P_Val": [None], "Notes": [note]}
    k = rec.n_terms - 1
    rounded = [float(rec.term_label[j]) for j in range(1, rec.n_terms)]
    dirs = ["Right" if rec.term_kind[j] == _abi.TERM_TP1 else "Left" for j in range(1, rec.n_terms)]
    brk = [float(pt_lineage[int(np.argmin(np.abs(pt_lineage - x)))]) for x in rounded]
    return {"Gene": [gene] * k, "Lineage": [letter] * k, "Breakpoint": brk, "Rounded_Breakpoint": rounded, "Direction": dirs,
            "Beta": [float(rec.coef[j]) for j in range(1, rec.n_terms)], "Test_Stat": [float(rec.z[j]) for j in range(1, rec.n_terms)],
            "P_Val": [float(rec.p[j]) for j in range(1, rec.n_terms)], "Notes": [None] * k}


def _summarize_model(rec, pt_min, pt_max):
    """summarizeModel.R:28-140 (GLM branch), including its duplicate-breakpoint indexing."""
    if not (rec.status & _abi.SCLANE_MARGE_OK):
        return {"Breakpoint": [None], "Slope.Segment": [None], "Trend.Segment": [None]}
    if rec.n_terms == 1:
        return {"Breakpoint": [None], "Slope.Segment": [0.0], "Trend.Segment": [0.0]}
    rows = [(float(rec.term_label[j]), "Right" if rec.term_kind[j] == _abi.TERM_TP1 else "Left", float(rec.coef[j]))
            for j in range(1, rec.n_terms)]
    rows.sort(key=lambda r: (r[0], r[1]))
    ranges = [(b, pt_max) if d == "Right" else (pt_min, b) for b, d, _ in rows]
    bps = [r[0] for r in rows]
    uniq = []
    for b in bps:
        if b not in uniq:
            uniq.append(b)
    nseg = len(uniq) + 1

    def bp(i):
        return bps[i - 1] if 1 <= i <= len(bps) else float("nan")

    segs = [(pt_min, bp(1)) if i == 1 else ((bp(i - 1), pt_max) if i == nseg else (bp(i - 1), bp(i))) for i in range(1, nseg + 1)]
    slopes = []
    for lo, hi in segs:
        s = 0.0
        for j, (a, b) in enumerate(ranges):
            if lo < b and hi > a:
                s += rows[j][2] if rows[j][1] == "Right" else -rows[j][2]
        slopes.append(s)
    trends = [1.0 if s > 0 else (-1.0 if s < 0 else 0.0) for s in slopes]
    return {"Breakpoint": uniq, "Slope.Segment": slopes, "Trend.Segment": trends}


def _record_to_list(rec, gene, lin_idx, info, pt_lineage, x_lineage, sf_lineage, preds, device, is_gee=False, gee_test="wald"):
    letter = LETTERS[lin_idx]
    m_ok = bool(rec.status & _abi.SCLANE_MARGE_OK)
    n_ok = bool(rec.status & _abi.SCLANE_NULL_OK)
    names = _term_names(rec, letter) if m_ok else []
    out = {
        "Gene": gene, "Lineage": letter,
        "Test_Stat": _na(rec.test_stat), "Test_Stat_Type": ("Score" if gee_test == "score" else "Wald") if is_gee else "LRT",
        "Test_Stat_Note": None if (m_ok and n_ok) else "No test performed due to model failure.",
        "Degrees_Freedom": _na(rec.df), "P_Val": _na(rec.p_val),
        "LogLik_MARGE": _na(rec.loglik), "LogLik_Null": _na(rec.null_loglik),
        "Dev_MARGE": _na(rec.deviance), "Dev_Null": _na(rec.null_deviance),
        "Model_Status": _status_string(rec),
        "MARGE_Fit_Notes": _notes(rec.marge_notes, m_ok), "Null_Fit_Notes": _notes(rec.null_notes, n_ok),
        "Gene_Time": rec.gene_time,
    }
    out["MARGE_Summary"] = ({"term": names, "estimate": [rec.coef[j] for j in range(rec.n_terms)],
                             "std.error": [rec.se[j] for j in range(rec.n_terms)],
                             "statistic": [rec.z[j] for j in range(rec.n_terms)],
                             "p.value": [rec.p[j] for j in range(rec.n_terms)]} if m_ok else None)
    out["Null_Summary"] = ({"term": ["Intercept"], "estimate": [rec.null_coef], "std.error": [rec.null_se],
                            "statistic": [rec.null_z], "p.value": [rec.null_p]} if n_ok else None)
    if preds == "none":
        out["MARGE_Preds"] = None
        out["Null_Preds"] = None
    else:
        # GEE mode: predict.geem() returns X %*% beta without the offset (pullMARGESummary.R:38-41); se from naiv.var
        sf_p = None if is_gee else sf_lineage
        out["MARGE_Preds"] = LazyPreds(rec, x_lineage, sf_p, "marge", device) if m_ok else None
        out["Null_Preds"] = LazyPreds(rec, x_lineage, sf_p, "null", device) if n_ok else None
        if preds == "eager":
            for k in ("MARGE_Preds", "Null_Preds"):
                if out[k] is not None:
                    out[k] = out[k].compute()
    out["MARGE_Slope_Data"] = _slope_data(rec, names, pt_lineage, gene, letter)
    gd = _summarize_model(rec, info.pt_min, info.pt_max)
    flat = {"Gene": [gene], "Lineage": [letter]}
    for key in ("Breakpoint", "Slope.Segment", "Trend.Segment"):
        vals = gd[key]
        if len(vals) == 1:
            flat[key] = [vals[0]]
        else:
            for q, v in enumerate(vals, 1):
                flat[f"{key}{q}"] = [v]
    out["Gene_Dynamics"] = flat
    # extras (not in the reference's list): fitted dispersion and the record itself
    out["Theta"] = (_na(rec.theta), _na(rec.null_theta))
    if is_gee:
        out["GEE_Working_Correlation"] = (_na(rec.gee_alpha), _na(rec.null_gee_alpha))
        out["GEE_Scale"] = (_na(rec.gee_phi), _na(rec.null_gee_phi))
    out["record"] = rec
    return out


# ------------------------------------------------------------------------------------------------
# the reference's public functions
# ------------------------------------------------------------------------------------------------
def createCellOffset(expr_mat, scale_factor: float = 1e4) -> np.ndarray:
    """createCellOffset.R:19-38: scale.factor / colSums(counts); expr.mat is genes x cells."""
    if expr_mat is None:
        raise ValueError("Please provide expr.mat to createCellOffset().")
    if _is_sparse(expr_mat):
        return float(scale_factor) / np.asarray(expr_mat.sum(axis=0), dtype=np.float64).reshape(-1)
    m = np.asarray(expr_mat)
    if m.ndim != 2:
        raise ValueError("Input expr.mat must be coerceable to a matrix of integer counts.")
    return float(scale_factor) / m.sum(axis=0).astype(np.float64)


def testDynamic(expr_mat=None, pt=None, genes=None, size_factor_offset=None, is_gee: bool = False,
                cor_structure: str = "ar1", gee_bias_correction_method=None, gee_test: str = "wald",
                is_glmm: bool = False, glmm_adaptive: bool = True, id_vec=None, n_potential_basis_fns: int = 5,
                n_cores: int = 4, approx_knot: bool = True, verbose: bool = False, random_seed: int = 312,
                n_gpus: int = 0, gpu_ids=None, genes_per_batch: int = 0, preds: str = "lazy",
                force_per_cell: bool = False) -> ScLANE:
    """testDynamic.R:60-448, GLM mode and GEE mode (is_gee = True with id_vec, cor_structure in {"ar1", "exchangeable",
    "independence"}, gee_test = "wald").  expr_mat: genes x cells integer counts (rows named by `genes`); pt: cells x
    lineages pseudotime (NaN = cell not in the lineage).  `n_cores` is kept for signature compatibility; the work is
    sharded over `n_gpus` B200s instead (0 = all visible).  Returns a ScLANE dict of per-gene, per-lineage lists."""
    if expr_mat is None or pt is None:
        raise ValueError("You forgot some inputs to testDynamic().")
    subject_id = None
    if is_gee or is_glmm:
        if id_vec is None:
            raise ValueError("You must provide a vector of IDs if you're using GEE / GLMM modes.")
        if is_gee and is_glmm:
            raise ValueError("is.gee and is.glmm cannot both be specified.")
        if is_glmm:
            raise SclaneError(_abi_unsupported(), "GLMM mode is not built into sclane_b200; GLM and GEE modes run on the GPU")
        cor_structure = str(cor_structure).lower()
        if cor_structure not in _abi.COR_CODES:
            raise ValueError("GEE models require a specified correlation structure.")
        gee_test = str(gee_test).lower()
        if gee_test not in ("wald", "score"):
            raise ValueError('gee.test must be one of "wald" or "score".')
        if gee_bias_correction_method is not None:
            gee_bias_correction_method = str(gee_bias_correction_method).lower()
            if gee_bias_correction_method not in ("kc", "df"):
                raise ValueError("Unsupported bias correction method in waldTestGEE().")
        ids = np.asarray(id_vec)
        _, first = np.unique(ids, return_index=True)
        subject_id = np.unique(ids, return_inverse=True)[1].astype(np.int32)
        # testDynamic.R:107-118: the data must be ordered by subject
        if np.count_nonzero(np.diff(subject_id) != 0) != len(first) - 1:
            raise ValueError("Data must be ordered by subject, please do so before running testDynamic() with is.gee = TRUE.")
    counts = expr_mat if _is_sparse(expr_mat) else np.ascontiguousarray(expr_mat, dtype=np.int32)
    if counts.ndim != 2:
        raise ValueError("Input expr.mat must be coerceable to a matrix of integer counts.")
    pt_arr = np.asarray(pt, dtype=np.float64)
    if pt_arr.ndim == 1:
        pt_arr = pt_arr[:, None]
    n_genes = int(counts.shape[0])
    if genes is None:
        genes = [f"Gene_{i + 1}" for i in range(n_genes)]
    genes = [str(g) for g in genes]
    if len(genes) != n_genes:
        raise ValueError("genes must name every row of expr.mat")
    t0 = time.time()
    o = _opts(n_potential_basis_fns, approx_knot, 25, None, 1e-5, n_gpus, gpu_ids, genes_per_batch, verbose, force_per_cell)
    if is_gee:
        o.mode = 1
        o.gee_cor = _abi.COR_CODES[cor_structure]
        o.gee_test = 1 if gee_test == "score" else 0
        o.gee_bias_correction = {None: 0, "kc": 1, "df": 2}[gee_bias_correction_method]
    recs, infos = run_records(counts, pt_arr, size_factor_offset, o, subject_id=subject_id)
    n_lin = pt_arr.shape[1]
    res = ScLANE()
    dev0 = int(gpu_ids[0]) if gpu_ids else 0
    lin_cache = []
    for l in range(n_lin):
        cells = np.flatnonzero(~np.isnan(pt_arr[:, l]))
        ptl = pt_arr[cells, l]
        sfl = None if size_factor_offset is None else np.asarray(size_factor_offset, dtype=np.float64)[cells]
        lin_cache.append((ptl, sfl))
    for g in range(n_genes):
        per_l = {}
        for l in range(n_lin):
            ptl, sfl = lin_cache[l]
            per_l[f"Lineage_{LETTERS[l]}"] = _record_to_list(recs[g * n_lin + l], genes[g], l, infos[l], ptl, ptl, sfl, preds, dev0, is_gee, gee_test if is_gee else "wald")
        res[genes[g]] = per_l
    res.lineage_info = [{"n_cells": int(i.n_cells), "n_knots": int(i.n_knots), "knots": [i.knots[k] for k in range(min(i.n_knots, _abi.SCLANE_TMAX))], "minspan": int(i.minspan),
                         "pt_min": i.pt_min, "pt_max": i.pt_max} for i in infos]
    res.profile = get_profile()
    res.records = recs
    res.is_gee = bool(is_gee)
    res.genes = genes
    if verbose:
        dt = time.time() - t0
        print(f"scLANE testing in {'GEE' if is_gee else 'GLM'} mode completed for {n_genes} genes across {n_lin} "
              f"{'lineage' if n_lin == 1 else 'lineages'} in {dt:.3f} secs")
    return res


def _abi_unsupported() -> int:
    return 4


def _raw_term_names(rec, var: str) -> list[str]:
    """colnames(B_final) (marge2.R:593-594): "Intercept", "(<var>-<knot>)" (tp1), "(<knot>-<var>)" (tp2)."""
    names = ["Intercept"]
    for j in range(1, rec.n_terms):
        lab = _fmt_num(rec.term_label[j])
        names.append(f"({var}-{lab})" if rec.term_kind[j] == _abi.TERM_TP1 else f"({lab}-{var})")
    return names


def marge2(X_pred=None, Y=None, Y_offset=None, M: int = 5, is_gee: bool = False, id_vec=None, cor_structure: str = "ar1",
           sandwich_var: bool = False, approx_knot: bool = True, n_knot_max: int = 25, tols_score: float = 1e-5, minspan=None,
           device: int = 0, is_glmm: bool = False) -> dict:
    """marge2.R:56-889 for one gene: X_pred = pseudotime of the lineage's cells, Y = counts, Y_offset = size factors; GEE mode
    with is_gee / id_vec / cor_structure / sandwich_var as in the reference.  Returns final_mod-like fields (coefficients,
    SEs, theta, logLik, deviance -- or the geem fields alpha / phi in GEE mode), coef_names, marge_coef_names, model_type."""
    if X_pred is None or Y is None:
        raise ValueError("Some required inputs to marge2() are missing.")
    y = np.ascontiguousarray(np.asarray(Y).reshape(1, -1), dtype=np.int32)
    x = np.asarray(X_pred, dtype=np.float64).reshape(-1)
    o = _opts(M, approx_knot, n_knot_max, minspan, tols_score, gpu_ids=[device])
    sid = None
    if is_glmm:
        # marge2.R:871-879: forward + backward pass only; the caller (fitGLMM) harvests the knots
        if is_gee:
            raise ValueError("is.gee and is.glmm cannot both be specified.")
        recs, infos = run_records(y, x[:, None], None, o, basis_only=True)
        rec = recs[0]
        if not (rec.status & _abi.SCLANE_MARGE_OK):
            raise SclaneError(0, _notes(rec.marge_notes, False))
        return {"final_mod": None, "basis_mtx": None, "WIC_mtx": None, "GCV": None, "model_type": "GLMM",
                "coef_names": _term_names(rec, "A"), "marge_coef_names": _raw_term_names(rec, "Lineage_A"), "record": rec,
                "knots": [infos[0].knots[k] for k in range(min(infos[0].n_knots, _abi.SCLANE_TMAX))]}
    if is_gee:
        if id_vec is None:
            raise ValueError("id.vec in marge2() must be non-null if is.gee = TRUE.")
        cor_structure = str(cor_structure).lower()
        if cor_structure not in _abi.COR_CODES:
            raise ValueError("cor.structure in marge2() must be a known type if is.gee = TRUE.")
        o.mode = 1
        o.gee_cor = _abi.COR_CODES[cor_structure]
        o.gee_bias_correction = 2 if sandwich_var else 0      # any non-zero value switches the fits to the sandwich variance
        sid = np.unique(np.asarray(id_vec), return_inverse=True)[1].astype(np.int32)
    recs, infos = run_records(y, x[:, None], Y_offset, o, subject_id=sid)
    rec = recs[0]
    if not (rec.status & _abi.SCLANE_MARGE_OK):
        raise SclaneError(0, _notes(rec.marge_notes, False))
    names = _term_names(rec, "A")
    p = rec.n_terms
    if is_gee:
        return {"final_mod": {"coefficients": dict(zip(names, [rec.coef[j] for j in range(p)])),
                              "std.error": [rec.se[j] for j in range(p)], "alpha": rec.gee_alpha, "phi": rec.gee_phi,
                              "converged": bool(rec.gee_converged), "niter": int(rec.gee_iters),
                              ("var" if sandwich_var else "naiv.var"): np.array([rec.cov[j] for j in range(p * p)]).reshape(p, p)},
                "basis_mtx": None, "WIC_mtx": None, "GCV": None, "model_type": "GEE",
                "coef_names": names, "marge_coef_names": names, "record": rec,
                "knots": [infos[0].knots[k] for k in range(min(infos[0].n_knots, _abi.SCLANE_TMAX))]}
    return {"final_mod": {"coefficients": dict(zip(names, [rec.coef[j] for j in range(p)])),
                          "std.error": [rec.se[j] for j in range(p)], "theta": rec.theta, "SE.theta": rec.se_theta,
                          "twologlik": 2.0 * rec.loglik, "logLik": rec.loglik, "df": p + 1, "deviance": rec.deviance,
                          "converged": not bool(rec.marge_notes & _abi.NOTE_NOT_CONVERGED),
                          "th.warn": _th_warn(rec.marge_notes),
                          "vcov": np.array([rec.cov[j] for j in range(p * p)]).reshape(p, p)},
            "basis_mtx": None, "WIC_mtx": None, "GCV": None, "model_type": "GLM",
            "coef_names": names, "marge_coef_names": names, "record": rec,
            "knots": [infos[0].knots[k] for k in range(min(infos[0].n_knots, _abi.SCLANE_TMAX))]}


def glmmKnots(X_pred=None, Y=None, id_vec=None, M_glm: int = 3, approx_knot: bool = True, device: int = 0) -> dict:
    """The adaptive knot harvesting of fitGLMM() (fitGLMM.R:54-75): marge2(is.glmm = TRUE, M = M.glm) on every subject's cells --
    ONE library call, a pt column per subject -- then the table of (knot, coef, old_coef, tp_fun) over the subjects whose model kept
    a hinge.  The GLMM fit itself (glmmTMB) is out of scope."""
    if X_pred is None or Y is None or id_vec is None:
        raise ValueError("You forgot some inputs to fitGLMM().")
    ids = np.asarray(id_vec)
    x = np.asarray(X_pred, dtype=np.float64).reshape(-1)
    if np.any(ids[1:] != ids[:-1]) and len(np.unique(ids)) != 1 + np.count_nonzero(ids[1:] != ids[:-1]):
        raise ValueError("Your data must be ordered by subject, please do so before running fitGLMM().")
    subjects = list(dict.fromkeys(ids.tolist()))
    pt = np.full((x.size, len(subjects)), np.nan)
    for k, sname in enumerate(subjects):
        pt[ids == sname, k] = x[ids == sname]
    y = np.ascontiguousarray(np.asarray(Y).reshape(1, -1), dtype=np.int32)
    o = _opts(M_glm, approx_knot, gpu_ids=[device])
    recs, _ = run_records(y, pt, None, o, basis_only=True)
    out = {"knot": [], "coef": [], "old_coef": [], "tp_fun": [], "subject": []}
    for k, sname in enumerate(subjects):
        rec = recs[k]
        if not (rec.status & _abi.SCLANE_MARGE_OK) or rec.n_terms <= 1:
            continue                                      # keepmods: only models with more than the intercept (fitGLMM.R:65-66)
        clean, raw = _term_names(rec, "A"), _raw_term_names(rec, "Lineage_A")
        for j in range(1, rec.n_terms):
            out["knot"].append(float(rec.term_label[j]))
            out["coef"].append(clean[j])
            out["old_coef"].append("B_final" + raw[j])
            out["tp_fun"].append("tp1" if rec.term_kind[j] == _abi.TERM_TP1 else "tp2")
            out["subject"].append(sname)
    return out


def _th_warn(bits: int):
    if bits & _abi.NOTE_ALT_LIMIT:
        return "alternation limit reached"
    if bits & _abi.NOTE_TH_ITER_LIMIT:
        return "iteration limit reached"
    if bits & _abi.NOTE_TH_TRUNC:
        return "estimate truncated at zero"
    return None


def modelLRT(mod_1=None, mod_0=None, is_glmm: bool = False) -> dict:
    """modelLRT.R:12-63.  mod_1: result of marge2() (or None/exception for a failed fit); mod_0: dict with logLik and df."""
    if mod_1 is None or mod_0 is None or isinstance(mod_1, Exception) or isinstance(mod_0, Exception):
        return {"Alt_Mod_LL": None, "Null_Mod_LL": None, "LRT_Stat": None, "DF": None, "P_Val": None, "Model_Type": None,
                "Notes": "No test performed due to model failure."}
    from scipy.stats import chi2  # host-side convenience only
    m1 = mod_1["final_mod"]
    ll1, ll0 = m1["logLik"], mod_0["logLik"]
    lrt = 2.0 * (ll1 - ll0)
    df = m1["df"] - mod_0["df"]
    if df == 0:
        df = 1
    return {"Alt_Mod_LL": ll1, "Null_Mod_LL": ll0, "LRT_Stat": lrt, "DF": df, "P_Val": float(chi2.sf(lrt, df)),
            "Model_Type": "GLM", "Notes": None}


def smoothedCountsMatrix(test_dyn_res=None, size_factor_offset=None, pt=None, genes=None, log1p_norm: bool = False,
                         n_cores: int = 2, device: int = 0) -> dict:
    """smoothedCountsMatrix.R:27-74: per lineage a cells x genes matrix of fitted means exp(marge_link_fit) (times the size
    factor when one is given), for every gene whose MARGE model fitted; one kernel launch per lineage
    (sclane_smoothed_counts).  Returns {"Lineage_A": {"matrix": ndarray [cells of the lineage, genes kept], "genes": [...]}}.
    `n_cores` is kept for signature compatibility."""
    if test_dyn_res is None or pt is None:
        raise ValueError("Please provide the scLANE output from testDynamic().")
    pt_arr = np.asarray(pt, dtype=np.float64)
    if pt_arr.ndim == 1:
        pt_arr = pt_arr[:, None]
    all_genes = list(test_dyn_res.keys())
    genes = all_genes if genes is None else [str(g) for g in genes]
    n_lin = pt_arr.shape[1]
    recs = (_abi.SclaneRecord * (len(genes) * n_lin))()
    for gi, g in enumerate(genes):
        for l in range(n_lin):
            recs[gi * n_lin + l] = test_dyn_res[g][f"Lineage_{LETTERS[l]}"]["record"]
    sf_all = None if size_factor_offset is None else np.asarray(size_factor_offset, dtype=np.float64)
    is_gee = bool(getattr(test_dyn_res, "is_gee", False))
    out = {}
    err = C.create_string_buffer(_ERRLEN)
    for l in range(n_lin):
        cells = np.flatnonzero(~np.isnan(pt_arr[:, l]))
        x = np.ascontiguousarray(pt_arr[cells, l])
        sfl = None if sf_all is None else np.ascontiguousarray(sf_all[cells])
        mat = np.empty((len(genes), x.size), dtype=np.float64)            # gene-major == column-major cells x genes
        rc = lib().sclane_smoothed_counts(recs, len(genes), n_lin, l, x.ctypes.data, x.size, None if sfl is None else sfl.ctypes.data,
                                          1 if is_gee else 0, 1 if log1p_norm else 0, device, mat.ctypes.data, err, _ERRLEN)
        check(rc, err)
        keep = [gi for gi in range(len(genes)) if recs[gi * n_lin + l].status & _abi.SCLANE_MARGE_OK]
        out[f"Lineage_{LETTERS[l]}"] = {"matrix": np.ascontiguousarray(mat[keep].T), "genes": [genes[gi] for gi in keep]}
    return out


def getFittedValues(test_dyn_res=None, genes=None, pt=None, expr_mat=None, size_factor_offset=None, log1p_norm: bool = False,
                    cell_meta_data=None, id_vec=None, ci_alpha: float = 0.05, filter_lineage=None, device: int = 0):
    """getFittedValues.R:45-150: long table (one row per cell, lineage and gene) of counts, link-scale fit / SE / CI and the
    response-scale fit / CI (multiplied by the size factor in GLM mode, :112-115).  expr_mat: genes x cells counts."""
    import pandas as pd
    from scipy.stats import norm
    if test_dyn_res is None or genes is None or pt is None or expr_mat is None:
        raise ValueError("You forgot one or more of the arguments to getFittedValues().")
    pt_arr = np.asarray(pt, dtype=np.float64)
    if pt_arr.ndim == 1:
        pt_arr = pt_arr[:, None]
    counts = np.asarray(expr_mat)
    all_genes = list(test_dyn_res.keys())
    z = float(norm.isf(ci_alpha / 2.0))
    is_gee = bool(getattr(test_dyn_res, "is_gee", False))
    sf_all = None if size_factor_offset is None else np.asarray(size_factor_offset, dtype=np.float64)
    frames = []
    for l in range(pt_arr.shape[1]):
        letter = LETTERS[l]
        if filter_lineage is not None and letter in filter_lineage:
            continue
        cells = np.flatnonzero(~np.isnan(pt_arr[:, l]))
        for g in [str(v) for v in genes]:
            rec = test_dyn_res[g][f"Lineage_{letter}"]
            df = {"cell": cells, "lineage": letter, "pt": pt_arr[cells, l], "gene": g,
                  "rna": counts[all_genes.index(g), cells].astype(np.float64)}
            if id_vec is not None:
                df["subj_id"] = np.asarray(id_vec)[cells]
            if sf_all is not None:
                df["size_factor"] = sf_all[cells]
                df["model_offset"] = np.log(1.0 / sf_all[cells])
            preds = rec["MARGE_Preds"]
            if preds is None:
                fit = se = np.full(cells.size, np.nan)
            else:
                fit, se = np.asarray(preds["marge_link_fit"]), np.asarray(preds["marge_link_se"])
            df.update({"scLANE_pred_link": fit, "scLANE_se_link": se, "scLANE_ci_ll_link": fit - z * se, "scLANE_ci_ul_link": fit + z * se,
                       "scLANE_pred": np.exp(fit), "scLANE_ci_ll": np.exp(fit - z * se), "scLANE_ci_ul": np.exp(fit + z * se)})
            if sf_all is not None and not is_gee:
                for k in ("rna", "scLANE_pred", "scLANE_ci_ll", "scLANE_ci_ul"):
                    df[k] = df[k] * sf_all[cells]
            if log1p_norm:
                for k in ("rna", "scLANE_pred", "scLANE_ci_ll", "scLANE_ci_ul"):
                    df[k + "_log1p"] = np.log1p(df[k])
            frame = pd.DataFrame(df)
            if cell_meta_data is not None:
                frame = pd.concat([frame, pd.DataFrame(cell_meta_data).iloc[cells].reset_index(drop=True)], axis=1)
            frames.append(frame)
    return pd.concat(frames, ignore_index=True)


def getResultsDE(test_dyn_res=None, p_adj_method: str = "fdr", fdr_cutoff: float = 0.01):
    """getResultsDE.R:23-45: first 15 fields -> table sorted by decreasing Test_Stat, BH adjustment, dynamic if
    P_Val_Adj < cutoff (NA -> 0), per-gene max over lineages.  Returns a pandas DataFrame."""
    if test_dyn_res is None:
        raise ValueError("Please provide a result list from testDynamic().")
    import pandas as pd
    keys = ["Gene", "Lineage", "Test_Stat", "Test_Stat_Type", "Test_Stat_Note", "Degrees_Freedom", "P_Val", "LogLik_MARGE",
            "LogLik_Null", "Dev_MARGE", "Dev_Null", "Model_Status", "MARGE_Fit_Notes", "Null_Fit_Notes", "Gene_Time"]
    rows = [{k: r[k] for k in keys} for lins in test_dyn_res.values() for r in lins.values()]
    df = pd.DataFrame(rows, columns=keys)
    df = df.sort_values("Test_Stat", ascending=False, kind="stable", na_position="last").reset_index(drop=True)
    padj = _p_adjust(df["P_Val"].to_numpy(dtype=np.float64, na_value=np.nan), p_adj_method)
    df["P_Val_Adj"] = padj
    df["Gene_Dynamic_Lineage"] = np.where(np.isnan(padj), 0, (padj < fdr_cutoff).astype(int))
    df["Gene_Dynamic_Overall"] = df.groupby("Gene")["Gene_Dynamic_Lineage"].transform("max")
    front = ["Gene", "Lineage", "Test_Stat", "P_Val"]
    return df[front + [c for c in df.columns if c not in front]]


def testSlope(test_dyn_res=None, p_adj_method: str = "fdr", fdr_cutoff: float = 0.01):
    """testSlope.R:21-39."""
    if test_dyn_res is None:
        raise ValueError("You forgot to provide results from testDynamic() to testSlope().")
    import pandas as pd
    frames = [pd.DataFrame(r["MARGE_Slope_Data"]) for lins in test_dyn_res.values() for r in lins.values()]
    df = pd.concat(frames, ignore_index=True)
    df["_abs"] = df["Test_Stat"].astype(float).abs()
    df = df.sort_values("_abs", ascending=False, kind="stable", na_position="last").drop(columns="_abs")
    df["P_Val_Adj"] = _p_adjust(df["P_Val"].to_numpy(dtype=np.float64, na_value=np.nan), p_adj_method)
    df = df.sort_values(["Gene", "Breakpoint"], kind="stable", na_position="last").reset_index(drop=True)
    padj = df["P_Val_Adj"].to_numpy(dtype=np.float64)
    df["Gene_Dynamic_Lineage_Slope"] = np.where(np.isnan(padj), 0, (padj < fdr_cutoff).astype(int))
    df["Gene_Dynamic_Lineage"] = df.groupby(["Gene", "Lineage"])["Gene_Dynamic_Lineage_Slope"].transform("max")
    df["Gene_Dynamic_Overall"] = df.groupby("Gene")["Gene_Dynamic_Lineage"].transform("max")
    return df
