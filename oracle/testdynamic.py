"""testDynamic() per-gene loop body, GLM mode (R/testDynamic.R:180-375), with the record
builders it calls (pullMARGESummary.R, pullNullSummary.R, createSlopeTestData.R,
summarizeModel.R, extractBreakpoints.R), modelLRT.R, createCellOffset.R and getResultsDE.R."""
from __future__ import annotations

import time

import numpy as np

from .knots import candidate_knots
from .marge import TP1, MargeResult, build_design, marge2
from .nbfit import glm_nb
from .rmath import p_adjust_bh, pchisq_upper, pnorm_two_sided, r_signif

LETTERS = "ABCDEFGHIJKLMNOPQRSTUVWXYZ"


def create_cell_offset(counts_gene_major: np.ndarray, scale_factor: float = 1e4) -> np.ndarray:
    """createCellOffset.R:32-36: scale.factor / colSums(counts) (cells are columns in R;
    here counts are genes x cells, so the sum runs over axis 0)."""
    return scale_factor / counts_gene_major.sum(axis=0).astype(np.float64)


def model_lrt(alt_fit, null_fit, p_alt: int):
    """modelLRT.R:12-63.  logLik df = rank + 1 for both models, so DF = p_alt - 1, with 0 -> 1."""
    if alt_fit is None or null_fit is None:
        return {"Alt_Mod_LL": np.nan, "Null_Mod_LL": np.nan, "LRT_Stat": np.nan, "DF": np.nan, "P_Val": np.nan,
                "Notes": "No test performed due to model failure."}
    lrt = 2.0 * (alt_fit.loglik - null_fit.loglik)
    df = (p_alt + 1) - (1 + 1)
    if df == 0:
        df = 1
    return {"Alt_Mod_LL": alt_fit.loglik, "Null_Mod_LL": null_fit.loglik, "LRT_Stat": lrt, "DF": df,
            "P_Val": pchisq_upper(lrt, df), "Notes": None}


def coef_table(names, fit):
    """broom.mixed::tidy(glm) columns (pullMARGESummary.R:70-74): summary.negbin uses dispersion 1."""
    z = fit.coef / fit.se
    return {"term": list(names), "estimate": [float(v) for v in fit.coef], "std.error": [float(v) for v in fit.se],
            "statistic": [float(v) for v in z], "p.value": [pnorm_two_sided(float(v)) for v in z]}


def link_preds(Xd, fit, offset):
    """predict(type = "link", se.fit = TRUE): fit = X beta + offset, se = sqrt(diag(X (X'WX)^-1 X'))."""
    eta = Xd @ fit.coef + (0.0 if offset is None else offset)
    se = np.sqrt(np.einsum("ij,jk,ik->i", Xd, fit.cov_unscaled, Xd))
    return eta, se


def slope_data(res: MargeResult, pt: np.ndarray, gene: str, lineage: str):
    """createSlopeTestData.R:15-74 (+ Gene/Lineage columns, testDynamic.R:295-298)."""
    if not res.ok:
        row = {"Breakpoint": [None], "Rounded_Breakpoint": [None], "Direction": [None], "Beta": [None],
               "Test_Stat": [None], "P_Val": [None], "Notes": ["MARGE model error"]}
    elif len(res.final_terms) == 1:
        row = {"Breakpoint": [None], "Rounded_Breakpoint": [None], "Direction": [None], "Beta": [None],
               "Test_Stat": [None], "P_Val": [None], "Notes": ["No non-intercept coefficients"]}
    else:
        hinges = res.final_terms[1:]
        rounded = [r_signif(float(res.knots[t.knot_idx]), 4) for t in hinges]     # parsed back from the names
        dirs = ["Right" if t.kind == TP1 else "Left" for t in hinges]
        brk = [float(pt[int(np.argmin(np.abs(pt - x)))]) for x in rounded]          # createSlopeTestData.R:46
        z = res.fit.coef / res.fit.se
        row = {"Breakpoint": brk, "Rounded_Breakpoint": rounded, "Direction": dirs,
               "Beta": [float(v) for v in res.fit.coef[1:]], "Test_Stat": [float(v) for v in z[1:]],
               "P_Val": [pnorm_two_sided(float(v)) for v in z[1:]], "Notes": [None] * len(hinges)}
    n = len(row["Notes"])
    return {"Gene": [gene] * n, "Lineage": [lineage] * n, **row}


def summarize_model(res: MargeResult, pt: np.ndarray):
    """summarizeModel.R:28-140 (GLM branch): per-segment slopes and trends."""
    if not res.ok:
        return {"Breakpoint": [None], "Slope.Segment": [None], "Trend.Segment": [None]}
    if len(res.final_terms) == 1:
        return {"Breakpoint": [None], "Slope.Segment": [0.0], "Trend.Segment": [0.0]}
    hinges = res.final_terms[1:]
    rows = [(r_signif(float(res.knots[t.knot_idx]), 4), "Right" if t.kind == TP1 else "Left", float(b))
            for t, b in zip(hinges, res.fit.coef[1:])]
    rows.sort(key=lambda r: (r[0], r[1]))                 # order(Breakpoint, Direction): "Left" < "Right"
    MIN, MAX = float(np.min(pt)), float(np.max(pt))
    ranges = [(b, MAX) if d == "Right" else (MIN, b) for b, d, _ in rows]
    bps = [r[0] for r in rows]
    uniq = []
    for b in bps:
        if b not in uniq:
            uniq.append(b)
    nseg = len(uniq) + 1
    # NB: the reference indexes z$Breakpoint (with duplicates) by segment number (summarizeModel.R:84-92)
    def bp(i):  # 1-based, NA beyond the end
        return bps[i - 1] if 1 <= i <= len(bps) else float("nan")
    segs = []
    for i in range(1, nseg + 1):
        if i == 1:
            segs.append((MIN, bp(1)))
        elif i == nseg:
            segs.append((bp(i - 1), MAX))
        else:
            segs.append((bp(i - 1), bp(i)))
    slopes = []
    for lo, hi in segs:
        idx = [j for j, (a, b) in enumerate(ranges) if lo < b and hi > a]
        s = 0.0
        for j in idx:
            s += rows[j][2] if rows[j][1] == "Right" else -rows[j][2]
        slopes.append(s)
    trends = [1.0 if s > 0 else (-1.0 if s < 0 else 0.0) for s in slopes]
    return {"Breakpoint": uniq, "Slope.Segment": slopes, "Trend.Segment": trends}


def test_dynamic(counts: np.ndarray, pt: np.ndarray, genes=None, size_factor_offset=None,
                 n_potential_basis_fns: int = 5, approx_knot: bool = True, keep_preds: bool = True):
    """GLM-mode testDynamic().  counts: genes x cells (int); pt: cells x lineages (NaN = not in
    lineage).  Returns {gene: {"Lineage_A": record, ...}} with the 21 fields of testDynamic.R:320-342."""
    counts = np.asarray(counts)
    pt = np.asarray(pt, dtype=np.float64)
    if pt.ndim == 1:
        pt = pt[:, None]
    G = counts.shape[0]
    if genes is None:
        genes = [f"G{i}" for i in range(G)]
    n_lin = pt.shape[1]
    out = {}
    caches = []
    for j in range(n_lin):
        cells = np.flatnonzero(~np.isnan(pt[:, j]))
        caches.append((cells, candidate_knots(pt[cells, j], 1, approx_knot)))
    for i in range(G):
        rec_l = {}
        for j in range(n_lin):
            cells, kc = caches[j]
            lin = LETTERS[j]
            y = counts[i, cells].astype(np.float64)
            sf = None if size_factor_offset is None else np.asarray(size_factor_offset, dtype=np.float64)[cells]
            t0 = time.time()
            off = None if sf is None else np.log(1.0 / sf)
            if y.sum() == 0.0:
                # Documented rule: a lineage with no counts for the gene is an error record for both
                # models (the reference stops in stat_out() -- "GCV.null argument to stat_out() cannot be 0",
                # stat_out.R:31 -- and its glm.nb degenerates to mu -> 0).
                from .marge import MargeResult, MargeTrace
                res = MargeResult(False, "all counts are zero", [], kc[1], kc[0], None, [], MargeTrace())
                null_fit, null_err = None, "all counts are zero"
            else:
                res = marge2(pt[cells, j], y, sf, M=n_potential_basis_fns, approx_knot=approx_knot,
                             var_name=f"Lineage_{lin}", knots_cache=kc)
                try:
                    null_fit = glm_nb(np.ones((y.size, 1)), y, off)
                    null_err = None
                except (FloatingPointError, np.linalg.LinAlgError, ZeroDivisionError) as e:
                    null_fit, null_err = None, str(e)
            gene_time = time.time() - t0
            status = ("MARGE model " + ("OK" if res.ok else "error") + ", null model " + ("OK" if null_fit is not None else "error"))
            rec = {"Gene": genes[i], "Lineage": lin, "Test_Stat": np.nan, "Test_Stat_Type": "LRT", "Test_Stat_Note": None,
                   "Degrees_Freedom": np.nan, "P_Val": np.nan,
                   "LogLik_MARGE": res.fit.loglik if res.ok else np.nan,
                   "LogLik_Null": null_fit.loglik if null_fit is not None else np.nan,
                   "Dev_MARGE": res.fit.deviance if res.ok else np.nan,
                   "Dev_Null": null_fit.deviance if null_fit is not None else np.nan,
                   "Model_Status": status, "MARGE_Fit_Notes": None if res.ok else res.error,
                   "Null_Fit_Notes": null_err, "Gene_Time": gene_time}
            if res.ok:
                Xd = build_design(res.final_terms, res.X, res.knots)
                rec["MARGE_Summary"] = coef_table(res.coef_names, res.fit)
                if keep_preds:
                    f, s = link_preds(Xd, res.fit, off)
                    rec["MARGE_Preds"] = {"marge_link_fit": f, "marge_link_se": s}
            else:
                rec["MARGE_Summary"] = None
                rec["MARGE_Preds"] = None
            if null_fit is not None:
                rec["Null_Summary"] = coef_table(["Intercept"], null_fit)
                if keep_preds:
                    f, s = link_preds(np.ones((y.size, 1)), null_fit, off)
                    rec["Null_Preds"] = {"null_link_fit": f, "null_link_se": s}
            else:
                rec["Null_Summary"] = None
                rec["Null_Preds"] = None
            rec["MARGE_Slope_Data"] = slope_data(res, pt[cells, j], genes[i], lin)
            gd = summarize_model(res, pt[cells, j])
            flat = {"Gene": [genes[i]], "Lineage": [lin]}
            for key in ("Breakpoint", "Slope.Segment", "Trend.Segment"):
                vals = gd[key]
                if len(vals) == 1:
                    flat[key] = [vals[0]]
                else:
                    for q, v in enumerate(vals, 1):
                        flat[f"{key}{q}"] = [v]
            rec["Gene_Dynamics"] = flat
            lrt = model_lrt(res.fit if res.ok else None, null_fit, len(res.final_terms))
            rec["Test_Stat"] = lrt["LRT_Stat"]
            rec["Test_Stat_Note"] = lrt["Notes"]
            rec["Degrees_Freedom"] = lrt["DF"]
            rec["P_Val"] = lrt["P_Val"]
            rec["_trace"] = res.trace
            rec["_theta"] = (res.fit.theta if res.ok else np.nan, null_fit.theta if null_fit is not None else np.nan)
            rec_l[f"Lineage_{lin}"] = rec
        out[genes[i]] = rec_l
    return out


def get_results_de(res: dict, fdr_cutoff: float = 0.01):
    """getResultsDE.R:23-45: rows sorted by decreasing Test_Stat (NA last), BH adjustment over all
    rows, Gene_Dynamic_Lineage = P_Val_Adj < cutoff (NA -> 0), Gene_Dynamic_Overall = per-gene max."""
    rows = []
    for g, lins in res.items():
        for ln, r in lins.items():
            rows.append({k: r[k] for k in ("Gene", "Lineage", "Test_Stat", "P_Val", "Degrees_Freedom", "Model_Status",
                                           "LogLik_MARGE", "LogLik_Null", "Dev_MARGE", "Dev_Null")})
    stat = np.array([np.nan if r["Test_Stat"] is None else r["Test_Stat"] for r in rows], dtype=np.float64)
    order = sorted(range(len(rows)), key=lambda i: (np.isnan(stat[i]), -stat[i] if not np.isnan(stat[i]) else 0.0))
    rows = [rows[i] for i in order]
    p = np.array([np.nan if r["P_Val"] is None else r["P_Val"] for r in rows], dtype=np.float64)
    padj = p_adjust_bh(p)
    for r, a in zip(rows, padj):
        r["P_Val_Adj"] = float(a)
        r["Gene_Dynamic_Lineage"] = 1 if (not np.isnan(a) and a < fdr_cutoff) else 0
    overall = {}
    for r in rows:
        overall[r["Gene"]] = max(overall.get(r["Gene"], 0), r["Gene_Dynamic_Lineage"])
    for r in rows:
        r["Gene_Dynamic_Overall"] = overall[r["Gene"]]
    return rows
