"""marge2() for one gene, GLM mode, q = 1 predictor (the only case testDynamic() produces,
testDynamic.R:189) -- dense restatement of R/marge2.R:56-889 with R/score_fun_glm.R,
R/stat_out_score_glm_null.R, R/stat_out.R and R/backward_sel_WIC.R.

Deterministic rules that replace platform-dependent behaviour of the reference:
  rule 1  RcppEigen::fastLmPure's numerical rank check (score_fun_glm.R:32-34) is replaced by
          the exact structural one: a candidate is NA iff [B XA] is exactly collinear, i.e.
          XA repeats a column already in B (same knot, same direction), or XA is a hinge PAIR
          at t2 while B already holds the full pair at some t1 (h1(t)-h2(t) = x - t for all t).
  rule 2  gamlss NBI fit -> exact NB MLE; vcov(type="all") -> analytic joint information;
          vcov() failure <=> Poisson regime (oracle/nbfit.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import flavours
from .eigen_helpers import eigen_map_mat_mult, eigen_map_matrix_invert
from .knots import candidate_knots, tp1, tp2
from .nbfit import NbMle, glm_nb, nb_mle, wald_nb
from .rmath import r_format_num, r_round, r_signif

TP1, TP2 = 1, 2  # hinge directions: tp1 = (x-t)+ "Right", tp2 = (t-x)+ "Left"


@dataclass(frozen=True)
class Term:
    """One basis column: the intercept (kind 0) or a hinge of direction TP1/TP2 at knots[idx]."""
    kind: int
    knot_idx: int = -1


def build_column(term: Term, X: np.ndarray, knots: np.ndarray) -> np.ndarray:
    if term.kind == 0:
        return np.ones_like(X)
    t = knots[term.knot_idx]
    return tp1(X, t) if term.kind == TP1 else tp2(X, t)


def build_design(terms, X, knots) -> np.ndarray:
    return np.stack([build_column(t, X, knots) for t in terms], axis=1)


# ---- stat_out_score_glm_null.R:18-48 ---------------------------------------------------
@dataclass
class NullStats:
    VS: np.ndarray       # (Y - mu) / V
    A: np.ndarray        # (B' W B)^-1
    B1: np.ndarray       # B' W  (p x N)
    mu: np.ndarray
    V: np.ndarray
    sigma: float
    fit: NbMle


def stat_out_score_glm_null(Y: np.ndarray, B_null: np.ndarray, beta0=None) -> NullStats:
    if flavours.STATE["fitter"] == "emulate_gamlss":      # the reference's own fitter, emulated (oracle/flavours.py)
        b, sg, mu_g, G, _ = flavours.gamlss_rs_nbi(B_null, Y)
        fit = NbMle(beta=b, sigma=sg, mu=mu_g, loglik=-0.5 * G, poisson=False)
    else:
        fit = nb_mle(B_null, Y, beta0)
    sigma = fit.sigma
    mu = fit.mu
    V = mu * (1.0 + mu * sigma)
    VS = (Y - mu) / V
    w = mu * mu / V                                   # diag(mu^2 / V), never materialised as N x N
    B1 = B_null.T * w[None, :]                        # t(B_null) %*% diag(w)
    A_inv = eigen_map_mat_mult(B1, B_null)
    if A_inv.shape == (1, 1):
        A = 1.0 / A_inv
    else:
        A = eigen_map_matrix_invert(A_inv)
    return NullStats(VS=VS, A=A, B1=B1, mu=mu, V=V, sigma=sigma, fit=fit)


# ---- score_fun_glm.R:21-54 ----------------------------------------------------------------
def score_fun_glm(ns: NullStats, XA: np.ndarray, rank_deficient: bool) -> float:
    if rank_deficient:                               # rule 1 (stands in for fastLmPure's NA coefficients)
        return float("nan")
    w = ns.mu * ns.mu / ns.V
    Bl = eigen_map_mat_mult(ns.B1, XA)               # p x c
    D = eigen_map_mat_mult(XA.T, XA * w[:, None])    # c x c
    tmp = eigen_map_mat_mult(eigen_map_mat_mult(Bl.T, ns.A), Bl)
    inv_XVX_22 = D - tmp
    B_est = eigen_map_mat_mult((ns.mu * ns.VS)[None, :], XA)   # 1 x c
    XVX_22 = eigen_map_matrix_invert(inv_XVX_22)
    score = eigen_map_mat_mult(eigen_map_mat_mult(B_est, XVX_22), B_est.T)
    return float(score[0, 0])


# ---- stat_out.R:23-47 -----------------------------------------------------------------------
def stat_out(Y, B1, TSS, GCV_null, pen=2.0):
    N = Y.size
    coef, *_ = np.linalg.lstsq(B1, Y, rcond=None)
    p = B1.shape[1]
    df1a = p + pen * (p - 1) / 2.0
    RSS1 = float(np.sum((Y - B1 @ coef) ** 2))
    RSSq1 = 1.0 - RSS1 / TSS
    GCV1 = RSS1 / (N * (1.0 - df1a / N) ** 2)
    GCVq1 = 1.0 - GCV1 / GCV_null
    return {"RSS1": RSS1, "RSSq1": r_round(RSSq1, 10), "GCV1": GCV1, "GCVq1": r_round(GCVq1, 10)}


# ---- selection tree, marge2.R:368-587 ------------------------------------------------------
def _isna(v):
    return np.isnan(v)


def _rmax(v):
    """max(v, na.rm = TRUE) (-Inf when everything is NA)."""
    ok = ~np.isnan(v)
    return float(np.max(v[ok])) if ok.any() else float("-inf")


def _which_max_round6(v):
    """which.max(round(v, 6)): FIRST index of the maximum, NA skipped; None if all NA."""
    r = r_round(v, 6)
    ok = ~np.isnan(r)
    if not ok.any():
        return None
    m = np.max(r[ok])
    return int(np.flatnonzero(ok & (r == m))[0])


def _last_max_round6(v):
    """utils::tail(which(max(round(v,6), na.rm=TRUE) == round(v,6), arr.ind=TRUE), n=1)[1]: LAST index."""
    r = r_round(v, 6)
    ok = ~np.isnan(r)
    if not ok.any():
        return None
    m = np.max(r[ok])
    return int(np.flatnonzero(ok & (r == m))[-1])


def select_candidate(both_int, both_add, one_int, one_add):
    """Returns (trunc_type, is_int, knot_index) or None for the `breakFlag` exit (marge2.R:388-391).
    Inputs are T-vectors (T x 1 matrices in the reference because q = 1 and B2 has one column)."""
    bi_all_na, oi_all_na = bool(np.all(_isna(both_int))), bool(np.all(_isna(one_int)))
    ba_all_na, oa_all_na = bool(np.all(_isna(both_add))), bool(np.all(_isna(one_add)))
    mx_bi, mx_oi, mx_ba, mx_oa = _rmax(both_int), _rmax(one_int), _rmax(both_add), _rmax(one_add)
    if bi_all_na and oi_all_na:                                        # :368
        if (not ba_all_na) and (not oa_all_na):
            if mx_ba > mx_oa:
                return 2, False, _last_max_round6(both_add)            # :372-374
            return 1, False, _last_max_round6(one_add)                 # :376-378
        if ba_all_na and not oa_all_na:
            return 1, False, _which_max_round6(one_add)                # :380-383
        if (not ba_all_na) and oa_all_na:
            return 2, False, _which_max_round6(one_add)                # :384-387 (sic: one_add -> NA -> error)
        return None                                                    # :388-391
    if bi_all_na and not oi_all_na:                                    # :392
        if ba_all_na:
            if oa_all_na:
                return 1, True, _last_max_round6(one_int)
            if mx_oi > mx_oa:
                return 1, True, _last_max_round6(one_int)
            return 1, False, _which_max_round6(one_add)                # :409-411
        if mx_oi > mx_ba:
            if mx_oi > mx_oa:
                return 1, True, _last_max_round6(one_int)
            return 1, False, _last_max_round6(one_add)                 # :424-428
        if mx_ba <= mx_oa:
            return 1, False, _which_max_round6(one_add)
        return 2, False, _which_max_round6(both_add)
    if (not bi_all_na) and oi_all_na:                                  # :443
        if ba_all_na:
            if mx_bi > mx_oa:
                return 2, True, _last_max_round6(both_int)
            return 1, False, _which_max_round6(one_add)
        if mx_bi > mx_ba:
            return 2, True, _last_max_round6(both_int)
        if mx_ba <= mx_oa:
            return 1, False, _which_max_round6(one_add)
        return 2, False, _which_max_round6(both_add)
    # :493 general case
    if mx_bi > mx_oi:
        if not ba_all_na:
            if mx_ba >= mx_bi:
                if mx_ba > mx_oa:
                    return 2, False, _which_max_round6(both_add)
                return 1, False, _which_max_round6(one_add)
            if mx_bi > mx_oa:
                return 2, True, _last_max_round6(both_int)
            return 1, False, _which_max_round6(one_add)
        if mx_oa >= mx_bi:
            return 1, False, _which_max_round6(one_add)
        return 2, True, _last_max_round6(both_int)
    if not ba_all_na:                                                  # :538
        if mx_ba >= mx_oi:
            if mx_ba <= mx_oa:
                return 1, False, _which_max_round6(one_add)            # :541-544
            return 2, False, _which_max_round6(both_add)               # :545-548
        if mx_oi > mx_oa:
            return 1, True, _last_max_round6(one_int)
        return 1, False, _which_max_round6(one_add)
    if oa_all_na:                                                      # :565-571
        return 1, True, _last_max_round6(one_int)
    if mx_oa >= mx_oi:
        return 1, False, _which_max_round6(one_add)
    return 1, True, _last_max_round6(one_int)


def _collinear(terms, cand_kinds, k) -> bool:
    """Rule 1: is [B, XA] exactly rank deficient for candidate hinge(s) `cand_kinds` at knot k?"""
    have = {(t.kind, t.knot_idx) for t in terms}
    for kind in cand_kinds:
        if (kind, k) in have:
            return True
    if len(cand_kinds) == 2:
        knots_tp1 = {t.knot_idx for t in terms if t.kind == TP1}
        knots_tp2 = {t.knot_idx for t in terms if t.kind == TP2}
        if knots_tp1 & knots_tp2:
            return True
    return False


@dataclass
class MargeTrace:
    forward_terms: list = field(default_factory=list)
    forward_scores: list = field(default_factory=list)    # score2 of every accepted/evaluated step
    score_tables: list = field(default_factory=list)      # per iteration dict of the four T-vectors
    null_fits: list = field(default_factory=list)
    wald_steps: list = field(default_factory=list)
    WIC_vec: list = field(default_factory=list)
    cnames: list = field(default_factory=list)
    stop_reason: str = ""


@dataclass
class MargeResult:
    ok: bool
    error: str | None
    final_terms: list            # list[Term], intercept first
    knots: np.ndarray
    X: np.ndarray
    fit: object | None           # NegBinFit
    coef_names: list
    trace: MargeTrace


def term_name(term: Term, knots, var_name: str) -> str:
    """Cleaned coefficient names, marge2.R:593-594 + :817-821."""
    if term.kind == 0:
        return "Intercept"
    lab = r_format_num(r_signif(float(knots[term.knot_idx]), 4))
    return f"h_{var_name}_{lab}" if term.kind == TP1 else f"h_{lab}_{var_name}"


def forward_pass(Y, X, knots, M=5, tols_score=1e-5, trace: MargeTrace | None = None):
    """marge2.R:80-759 for q = 1.  Returns the list of forward-pass terms."""
    NN = Y.size
    pen = 2.0
    TSS = float(np.sum((Y - Y.mean()) ** 2))
    GCV_null = TSS / (NN * (1.0 - 1.0 / NN) ** 2)
    terms = [Term(0)]
    m = k = 1
    T = knots.size
    hinge1 = [tp1(X, t) for t in knots]
    hinge2 = [tp2(X, t) for t in knots]
    beta0 = None
    while True:
        B = build_design(terms, X, knots)
        ns = stat_out_score_glm_null(Y, B, beta0)
        if trace is not None:
            trace.null_fits.append(ns.fit)
        in_set = 1 if len(terms) > 1 else 0      # marge2.R:180 (q = 1: only "Intercept" differs from var_name)
        both_add = np.full(T, np.nan)
        one_add = np.full(T, np.nan)
        colpiv = flavours.STATE["rank_rule"] == "colpiv_qr"      # the reference's numerical rank check, emulated

        unchecked = flavours.STATE["rank_rule"] == "pair_unchecked"

        def rank_def(XA, kinds, t):
            if colpiv:
                return flavours.rank_deficient_colpiv(B, XA)
            if unchecked:                                        # only exactly duplicated columns are NA
                return any((t2.kind, t2.knot_idx) == (kind, t) for t2 in terms for kind in kinds)
            return _collinear(terms, kinds, t)

        for t in range(T):
            XA2 = np.stack([hinge1[t], hinge2[t]], 1)
            both_add[t] = score_fun_glm(ns, XA2, rank_def(XA2, (TP1, TP2), t))
            one_add[t] = score_fun_glm(ns, hinge1[t][:, None], rank_def(hinge1[t][:, None], (TP1,), t))
        if in_set == 0:
            both_int = np.full(T, -1e5)          # marge2.R:244
            one_int = np.full(T, -1e5)
        else:
            both_int = both_add.copy()           # B2 is the intercept column: identical designs (:247-307)
            one_int = one_add.copy()
        if trace is not None:
            trace.score_tables.append({"both_add": both_add, "one_add": one_add})
        sel = select_candidate(both_int, both_add, one_int, one_add)
        if sel is None:
            if trace is not None:
                trace.stop_reason = "all candidates NA"
            break
        trunc_type, is_int, kidx = sel
        if kidx is None:
            raise RuntimeError("selection produced no knot (all-NA score vector)")   # R: subscript error
        # q = 1: an "interaction with the intercept" builds the same columns as the additive case
        new_terms = [Term(TP1, kidx), Term(TP2, kidx)] if trunc_type == 2 else [Term(TP1, kidx)]
        XA = build_design(new_terms, X, knots)
        score2 = score_fun_glm(ns, XA, rank_def(XA, tuple(t.kind for t in new_terms), kidx))
        B_temp = np.concatenate([B, XA], axis=1)
        so = stat_out(Y, B_temp, TSS, GCV_null, pen)
        if trace is not None:
            trace.forward_scores.append(score2)
        if np.isnan(score2):
            raise RuntimeError("missing value where TRUE/FALSE needed")              # if (NA) in R
        if so["GCVq1"] < -10 or r_round(score2, 4) <= 0:                            # :681
            if trace is not None:
                trace.stop_reason = "rejected (GCV / non-positive score)"
            break
        if score2 < tols_score:                                                     # :737
            if trace is not None:
                trace.stop_reason = "score below tols_score"
            break
        beta0 = np.concatenate([ns.fit.beta, np.zeros(len(new_terms))])
        terms = terms + new_terms
        k += 1
        m += 2
        if NN <= len(terms) + 2 or m >= M:                                          # :756
            if trace is not None:
                trace.stop_reason = "M reached"
            break
    if trace is not None:
        trace.forward_terms = list(terms)
    return terms


def backward_sel_wic(Y, B):
    """backward_sel_WIC.R:44-52 (GLM branch)."""
    if flavours.STATE["fitter"] == "emulate_gamlss":
        wald, _ = flavours.gamlss_wald(B, Y)
        return wald, None
    wald, fit = wald_nb(B, Y)
    return wald, fit


def _first_min(wic1) -> int:
    """which(wic1 == min(wic1, na.rm = TRUE))[1] - 1; with every entry NA, R's NA subscript turns the basis into NAs and
    the next refit fails inside try() -> "MARGE model error" (marge2.R:781-783)."""
    hit = np.flatnonzero(np.asarray(wic1) == _nanmin(wic1))
    if hit.size == 0:
        raise RuntimeError("missing value where TRUE/FALSE needed (all Wald statistics NA)")
    return int(hit[0])


def _nanmin(v):
    ok = ~np.isnan(v)
    return float(np.min(v[ok])) if ok.any() else float("inf")


def backward_pass(Y, X, knots, terms, trace: MargeTrace | None = None):
    """marge2.R:763-812 with its index quirks.  Returns the final term list.
    Raises for ncol(B) <= 2, where the reference hits a subscript error (-> "MARGE model error")."""
    ncol_B = len(terms)
    if ncol_B <= 2:
        raise RuntimeError("subscript out of bounds")          # marge2.R:777 / :788
    WIC_vec = [np.nan]
    cur = list(terms)
    models = [list(cur)]
    wic1, fit = backward_sel_wic(Y, build_design(cur, X, knots))
    if trace is not None:
        trace.wald_steps.append(np.array(wic1))
    cum = _nanmin(wic1)
    WIC_vec.append(cum + 2.0 * len(cur))
    lowest = _first_min(wic1) + 1      # 1-based position among non-intercept cols
    cur = [t for j, t in enumerate(cur) if j != lowest]
    models.append(list(cur))
    for i in range(2, ncol_B):
        wic1, fit = backward_sel_wic(Y, build_design(cur, X, knots))
        if trace is not None:
            trace.wald_steps.append(np.array(wic1))
        cum += _nanmin(wic1)
        WIC_vec.append(cum + 2.0 * len(cur))
        if i != ncol_B - 1:
            lowest = _first_min(wic1) + 1
            cur = [t for j, t in enumerate(cur) if j != lowest]
        else:
            cur = [Term(0)]                                       # renamed "Intercept" (:798-799)
        models.append(list(cur))
    WIC = np.array(WIC_vec)
    ok = ~np.isnan(WIC)
    best = int(np.flatnonzero(ok & (WIC == np.min(WIC[ok])))[0])
    final = models[best]
    # B[, colnames(B) %in% cnames]: forward order is preserved; duplicated columns dropped (:807-812)
    final_set = set(final)
    out = []
    for t in terms:
        if t in final_set and t not in out:
            out.append(t)
    if trace is not None:
        trace.WIC_vec = WIC_vec
        trace.cnames = models
    return out


def marge2(x_pred, Y, Y_offset=None, M=5, approx_knot=True, n_knot_max=25, tols_score=1e-5,
           minspan=None, var_name="Lineage_A", knots_cache=None, want_trace=True) -> MargeResult:
    """marge2.R:56-889, GLM mode.  `Y_offset` is the size factor (enters as offset(log(1/sf)), :826-829)."""
    Y = np.asarray(Y, dtype=np.float64)
    if knots_cache is None:
        X, knots = candidate_knots(np.asarray(x_pred, dtype=np.float64), 1, approx_knot, n_knot_max, minspan)
    else:
        X, knots = knots_cache
    trace = MargeTrace()
    try:
        terms = forward_pass(Y, X, knots, M, tols_score, trace)
        final = backward_pass(Y, X, knots, terms, trace)
        Bf = build_design(final, X, knots)
        off = None if Y_offset is None else np.log(1.0 / np.asarray(Y_offset, dtype=np.float64))
        fit = glm_nb(Bf, Y, off)
    except (RuntimeError, FloatingPointError, np.linalg.LinAlgError) as e:   # try(..., silent=TRUE), testDynamic.R:188
        return MargeResult(False, str(e), [], knots, X, None, [], trace)
    names = [term_name(t, knots, var_name) for t in final]
    return MargeResult(True, None, final, knots, X, fit, names, trace)
