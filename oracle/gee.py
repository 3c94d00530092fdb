"""GEE mode of the path (is.gee = TRUE): dense restatement of R/stat_out_score_gee_null.R, R/score_fun_gee.R,
the GEE branches of R/backward_sel_WIC.R:36-42 and R/marge2.R:831-838, R/testDynamic.R:235-244 and
R/waldTestGEE.R:24-93.

PARITY UNPINNED for this mode: the reference bundles no GEE results and its tests assert only shapes
(tests/testthat/test_scLANE.R:118-157,405-430), and the fitter -- geeM::geem -- is third-party code that is not in the
reference tree.  `geem()` below restates geeM 0.10.x's algorithm from its published source as recalled:
moment updates of phi and alpha from Pearson residuals, closed-form R(alpha)^-1, <= 10 inner Newton steps per outer
iteration, <= 20 outer iterations, relative-change stopping rule 1e-5.  Because the fit stops at a 1e-5 relative change,
its result depends on the iteration path; the CUDA path follows the same schedule step by step.

Dense on purpose: per-subject n_i x n_i working covariances are built and inverted exactly as the reference does
(O(n_i^3)), so this file also validates the O(n_i) closed forms the kernels use.  Use small subjects here.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .eigen_helpers import eigen_map_mat_mult as mm, eigen_map_matrix_invert as llt_inv
from .knots import candidate_knots, tp1, tp2
from .marge import TP1, TP2, MargeTrace, Term, _collinear, _first_min, _nanmin, build_design, select_candidate, stat_out, term_name
from .rmath import pchisq_upper, pnorm_two_sided, r_round

NB_THETA = 50.0      # family = MASS::negative.binomial(50, link = "log") at every geem call of the path


def _corr_matrix(cor: str, alpha: float, n: int) -> np.ndarray:
    """stat_out_score_gee_null.R:51-60."""
    if cor == "independence":
        return np.eye(n)
    if cor == "exchangeable":
        return np.full((n, n), alpha) + np.eye(n) * (1.0 - alpha)
    if cor == "ar1":
        idx = np.arange(n)
        return alpha ** np.abs(idx[:, None] - idx[None, :])
    raise ValueError("Currently unsupported correlation structure.")


def _rinv_blocks(cor: str, alpha: float, n_vec) -> list[np.ndarray]:
    """geeM getAlphaInvAR / getAlphaInvEX: closed-form R(alpha)^-1 per cluster."""
    out = []
    for n in n_vec:
        n = int(n)
        if cor == "independence" or n == 1:
            out.append(np.eye(n))
        elif cor == "exchangeable":
            out.append((np.eye(n) - alpha / (1.0 + (n - 1) * alpha) * np.ones((n, n))) / (1.0 - alpha))
        else:
            m = np.zeros((n, n))
            d = np.full(n, 1.0 + alpha * alpha)
            d[0] = d[-1] = 1.0
            m[np.arange(n), np.arange(n)] = d
            m[np.arange(n - 1), np.arange(1, n)] = -alpha
            m[np.arange(1, n), np.arange(n - 1)] = -alpha
            out.append(m / (1.0 - alpha * alpha))
    return out


def _rcond_ok(A: np.ndarray) -> np.ndarray:
    """Inverse of a small system matrix with the reciprocal-condition-number guard of R's solve(): 1-norm rcond below
    .Machine$double.eps (or a non-finite entry) is an error ("system is computationally singular"), which try() turns
    into a model failure.  rcond is computed exactly here (LAPACK estimates it); DESIGN.md section 7."""
    if not np.all(np.isfinite(A)):
        raise RuntimeError("system is computationally singular (non-finite entries)")
    Ainv = np.linalg.inv(A)
    n1 = np.abs(A).sum(axis=0).max() * np.abs(Ainv).sum(axis=0).max()
    if not np.isfinite(n1) or 1.0 / n1 < np.finfo(float).eps:
        raise RuntimeError("system is computationally singular")
    return Ainv


@dataclass
class GeemFit:
    beta: np.ndarray
    alpha: float
    phi: float
    naiv_var: np.ndarray
    var: np.ndarray | None      # robust (sandwich) variance when sandwich = TRUE
    eta: np.ndarray
    mu: np.ndarray
    converged: bool
    niter: int


def geem(X, y, n_vec, cor: str, offset=None, maxit: int = 20, tol: float = 1e-5, sandwich: bool = False) -> GeemFit:
    """geeM::geem(y ~ X - 1, id, family = negative.binomial(50), corstr = cor, scale.fix = FALSE) [3P, from memory]."""
    y = np.asarray(y, dtype=np.float64)
    N, p = X.shape
    off = np.zeros(N) if offset is None else np.asarray(offset, dtype=np.float64)
    beta = np.zeros(p)
    ones = [j for j in range(p) if np.all(X[:, j] == 1.0)]
    if not ones:
        raise RuntimeError("Must supply an initial beta if not using an intercept.")
    beta[ones[0]] = np.log(y.mean()) - off.mean()
    starts = np.concatenate([[0], np.cumsum(n_vec)]).astype(int)
    var = lambda m: m + m * m / NB_THETA
    phi, alpha = 1.0, 0.0
    beta_old = beta.copy()
    converged = False
    hess = None
    count = 0
    while True:
        count += 1
        eta = X @ beta + off
        mu = np.exp(eta)
        se = np.sqrt(1.0 / var(mu))
        resid = se * (y - mu)
        phi = float(resid @ resid) / (N - p)                                        # updatePhi, useP = TRUE
        if cor == "ar1":
            num = sum(float(resid[a:b - 1] @ resid[a + 1:b]) for a, b in zip(starts[:-1], starts[1:]))
            den = phi * (float(np.sum(np.asarray(n_vec) - 1)) - p)
            alpha = num / den
        elif cor == "exchangeable":
            num = sum(0.5 * (float(resid[a:b].sum()) ** 2 - float(resid[a:b] @ resid[a:b])) for a, b in zip(starts[:-1], starts[1:]))
            den = phi * (float(np.sum(np.asarray(n_vec) * (np.asarray(n_vec) - 1) / 2.0)) - p)
            alpha = num / den
        else:
            alpha = 0.0
        rinv = _rinv_blocks(cor, alpha, n_vec)
        # updateBeta: up to 10 Newton steps at fixed R(alpha)^-1 / phi
        bnew = beta.copy()
        for _ in range(10):
            eta = X @ bnew + off
            mu = np.exp(eta)
            se = np.sqrt(1.0 / var(mu))
            Z = (se * mu)[:, None] * X                                              # StdErr %*% dInvLinkdEta %*% X
            rr = se * (y - mu)
            hess = np.zeros((p, p))
            esteq = np.zeros(p)
            for (a, b), Ri in zip(zip(starts[:-1], starts[1:]), rinv):
                hess += Z[a:b].T @ (Ri / phi) @ Z[a:b]
                esteq += Z[a:b].T @ (Ri / phi) @ rr[a:b]
            _rcond_ok(hess)
            update = np.linalg.solve(hess, esteq)
            bnew = bnew + update
            with np.errstate(all="ignore"):
                if np.max(np.abs(update) / bnew) < 100.0 * tol:                      # sic: no abs() on beta.new
                    break
        beta = bnew
        with np.errstate(all="ignore"):
            if np.max(np.abs((beta - beta_old) / (beta_old + np.finfo(float).eps))) < tol:
                converged = True
                break
        if count >= maxit:
            break
        beta_old = beta.copy()
    eta = X @ beta + off
    naiv = _rcond_ok(hess)
    var_rob = None
    if sandwich:
        # geeM getSandwich(): numsand = sum_i e_i e_i', e_i = Z_i' (R_i^-1 / phi) rr_i at the FINAL eta with the alpha / phi
        # of the last outer iteration; var = hess^-1 numsand hess^-T with the hessian of the last inner step
        mu = np.exp(eta)
        se = np.sqrt(1.0 / var(mu))
        Z = (se * mu)[:, None] * X
        rr = se * (y - mu)
        numsand = np.zeros((p, p))
        for (a, b), Ri in zip(zip(starts[:-1], starts[1:]), rinv):
            e = Z[a:b].T @ (Ri / phi) @ rr[a:b]
            numsand += np.outer(e, e)
        var_rob = naiv @ numsand @ naiv.T
    return GeemFit(beta=beta, alpha=float(alpha), phi=float(phi), naiv_var=naiv, var=var_rob, eta=eta, mu=np.exp(eta),
                   converged=converged, niter=count)


# ---- stat_out_score_gee_null.R:22-112 ----------------------------------------------------------------------
@dataclass
class GeeNull:
    VS: list
    AWA: list
    J2: list
    Sigma2: list
    J11_inv: np.ndarray
    JSigma11: np.ndarray
    mu: np.ndarray
    V: np.ndarray
    fit: GeemFit


def stat_out_score_gee_null(Y, B, n_vec, cor: str) -> GeeNull:
    fit = geem(B, Y, n_vec, cor)
    alpha, sigma = fit.alpha, fit.phi                                     # :41-42 (phi is then used as the NB dispersion)
    mu = fit.mu
    V = mu * (1.0 + mu * sigma)                                           # :44
    p = B.shape[1]
    starts = np.concatenate([[0], np.cumsum(n_vec)]).astype(int)
    VS, AWA, J2, S2 = [], [], [], []
    J11 = np.zeros((p, p))
    S11 = np.zeros((p, p))
    for a, b in zip(starts[:-1], starts[1:]):
        n = b - a
        R = _corr_matrix(cor, alpha, n)
        dsv = np.diag(np.sqrt(V[a:b]))
        Vi = mm(mm(dsv, R), dsv)
        Vinv = llt_inv(Vi)
        S = (Y[a:b] - mu[a:b])[:, None]
        AWAi = mm(mm(Vinv, mm(S, S.T)), Vinv)
        D = mm(np.diag(mu[a:b]), B[a:b])
        J2i = mm(D.T, Vinv)
        J11 += mm(J2i, D)
        S2i = mm(D.T, AWAi)
        S11 += mm(S2i, D)
        VS.append(mm(Vinv, S))
        AWA.append(AWAi)
        J2.append(J2i)
        S2.append(S2i)
    J11_inv = 1.0 / J11 if p == 1 else llt_inv(J11)
    JS = mm(mm(J11_inv, S11), J11_inv)
    return GeeNull(VS, AWA, J2, S2, J11_inv, JS, mu, V, fit)


# ---- score_fun_gee.R:28-85 -----------------------------------------------------------------------------------
def score_fun_gee(gn: GeeNull, n_vec, XA: np.ndarray, rank_deficient: bool) -> float:
    if rank_deficient:
        return float("nan")
    c = XA.shape[1]
    p1 = gn.J11_inv.shape[0]
    starts = np.concatenate([[0], np.cumsum(n_vec)]).astype(int)
    Best = np.zeros((c, 1))
    S22 = np.zeros((c, c))
    J21 = np.zeros((c, p1))
    S21 = np.zeros((c, p1))
    for i, (a, b) in enumerate(zip(starts[:-1], starts[1:])):
        D = mm(np.diag(gn.mu[a:b]), XA[a:b])
        J21 += mm(D.T, gn.J2[i].T)
        S21 += mm(D.T, gn.Sigma2[i].T)
        Best += mm(D.T, gn.VS[i])
        S22 += mm(mm(D.T, gn.AWA[i]), D)
    t1 = mm(mm(J21, gn.J11_inv), S21.T)
    t2 = mm(mm(S21, gn.J11_inv), J21.T)
    t3 = mm(mm(J21, gn.JSigma11), J21.T)
    Sigma = S22 - t1 - t2 + t3
    Sinv = llt_inv(Sigma)
    return float(mm(mm(Best.T, Sinv), Best)[0, 0])


# ---- marge2(is.gee = TRUE) ----------------------------------------------------------------------------------------
def forward_pass_gee(Y, X, knots, n_vec, cor, M=5, tols_score=1e-5, trace=None):
    """marge2.R:109-759 with the GEE score functions (q = 1)."""
    NN = Y.size
    TSS = float(np.sum((Y - Y.mean()) ** 2))
    GCV_null = TSS / (NN * (1.0 - 1.0 / NN) ** 2)
    terms = [Term(0)]
    m = 1
    T = knots.size
    h1 = [tp1(X, t) for t in knots]
    h2 = [tp2(X, t) for t in knots]
    while True:
        B = build_design(terms, X, knots)
        gn = stat_out_score_gee_null(Y, B, n_vec, cor)
        if trace is not None:
            trace.null_fits.append(gn.fit)
        in_set = 1 if len(terms) > 1 else 0
        both = np.full(T, np.nan)
        one = np.full(T, np.nan)
        for t in range(T):
            both[t] = score_fun_gee(gn, n_vec, np.stack([h1[t], h2[t]], 1), _collinear(terms, (TP1, TP2), t))
            one[t] = score_fun_gee(gn, n_vec, h1[t][:, None], _collinear(terms, (TP1,), t))
        bi = both.copy() if in_set else np.full(T, -1e5)
        oi = one.copy() if in_set else np.full(T, -1e5)
        if trace is not None:
            trace.score_tables.append({"both_add": both, "one_add": one})
        sel = select_candidate(bi, both, oi, one)
        if sel is None:
            break
        tt, _, kidx = sel
        if kidx is None:
            raise RuntimeError("selection produced no knot (all-NA score vector)")
        new_terms = [Term(TP1, kidx), Term(TP2, kidx)] if tt == 2 else [Term(TP1, kidx)]
        score2 = both[kidx] if tt == 2 else one[kidx]
        if trace is not None:
            trace.forward_scores.append(score2)
        if np.isnan(score2):
            raise RuntimeError("missing value where TRUE/FALSE needed")
        so = stat_out(Y, np.concatenate([B, build_design(new_terms, X, knots)], 1), TSS, GCV_null)
        if so["GCVq1"] < -10 or r_round(score2, 4) <= 0:
            break
        if score2 < tols_score:
            break
        terms = terms + new_terms
        m += 2
        if NN <= len(terms) + 2 or m >= M:
            break
    if trace is not None:
        trace.forward_terms = list(terms)
    return terms


def backward_sel_wic_gee(Y, B, n_vec, cor, sandwich=False):
    """backward_sel_WIC.R:36-42: geem refit, summary(fit)$wald.test[-1]^2 -- beta / se.model, or beta / se.robust when the
    fit carries a sandwich variance (summary.geem)."""
    fit = geem(B, Y, n_vec, cor, sandwich=sandwich)
    se = np.sqrt(np.diag(fit.var if sandwich else fit.naiv_var))
    return (fit.beta / se)[1:] ** 2, fit


def backward_pass_gee(Y, X, knots, terms, n_vec, cor, trace=None, sandwich=False):
    """marge2.R:763-812 (same bookkeeping as the GLM branch)."""
    ncol_B = len(terms)
    if ncol_B <= 2:
        raise RuntimeError("subscript out of bounds")
    WIC = [np.nan]
    cur = list(terms)
    models = [list(cur)]
    cum = 0.0
    lowest = 0
    for i in range(1, ncol_B):
        wic1, _ = backward_sel_wic_gee(Y, build_design(cur, X, knots), n_vec, cor, sandwich)
        if trace is not None:
            trace.wald_steps.append(np.array(wic1))
        cum += _nanmin(wic1)
        WIC.append(cum + 2.0 * len(cur))
        if i != ncol_B - 1:
            lowest = _first_min(wic1) + 1
            cur = [t for j, t in enumerate(cur) if j != lowest]
        else:
            cur = [Term(0)]
        models.append(list(cur))
    W = np.array(WIC)
    ok = ~np.isnan(W)
    best = int(np.flatnonzero(ok & (W == np.min(W[ok])))[0])
    fin = set(models[best])
    out = []
    for t in terms:
        if t in fin and t not in out:
            out.append(t)
    if trace is not None:
        trace.WIC_vec = WIC
    return out


def bias_correct_gee(fit: GeemFit, method: str, n_subjects: int, n_obs: int) -> np.ndarray:
    """biasCorrectGEE.R:17-158.  "df": n_s / (n_s - p) x sandwich (unchanged when p >= n_s, :36-38).  "kc":
    (n_s / (n_s - 1)) / (1 - tr(H) / n) x sandwich with H = X (X'WX)^-1 X'W a projection of rank p, so tr(H) = p whatever
    the per-subject correlation estimates in W are (:125-137)."""
    p = fit.beta.size
    if method == "df":
        return fit.var if p >= n_subjects else (n_subjects / (n_subjects - p)) * fit.var
    if method == "kc":
        return ((n_subjects / (n_subjects - 1.0)) / (1.0 - p / n_obs)) * fit.var
    raise ValueError("Unsupported bias correction method in waldTestGEE().")


def wald_test_gee(fit: GeemFit, vcov=None):
    """waldTestGEE.R:24-93 (correction.method = NULL => naive variance; otherwise the bias-corrected sandwich)."""
    p = fit.beta.size
    if p <= 1:
        return {"Wald_Stat": 0.0, "DF": 0, "P_Val": 1.0}
    L = np.eye(p)[1:]
    middle = L @ (fit.naiv_var if vcov is None else vcov) @ L.T
    sides = L @ fit.beta
    w = float(sides @ llt_inv(middle) @ sides)
    return {"Wald_Stat": w, "DF": p - 1, "P_Val": 1.0 - (1.0 - pchisq_upper(w, p - 1))}


def score_test_gee(X_alt: np.ndarray, y: np.ndarray, null: GeemFit, n_vec, cor: str):
    """scoreTestGEE.R:26-139, dense like the reference: per-subject V_i = A^1/2 R(rho) A^1/2 with the NB(50) variance
    function at the null fit, W_i = K V_i^-1 K, U = t(X) W K^-1 r / phi, V_U = t(X) W X, all inverses through
    eigenMapMatrixInvert (LLT)."""
    p_alt = X_alt.shape[1]
    if p_alt <= 1:
        return {"Score_Stat": 0.0, "DF": 0, "P_Val": 1.0}
    rho, phi = null.alpha, null.phi
    mu = np.exp(null.eta)
    r_null = y - mu
    starts = np.concatenate([[0], np.cumsum(n_vec)]).astype(int)
    U = np.zeros((p_alt, 1))
    V_U = np.zeros((p_alt, p_alt))
    for a, b in zip(starts[:-1], starts[1:]):
        n = b - a
        R = _corr_matrix(cor, rho, n)
        A_sqrt = np.diag(np.sqrt(mu[a:b] + mu[a:b] ** 2 / NB_THETA))
        V_inv = llt_inv(A_sqrt @ R @ A_sqrt)
        K = np.diag(mu[a:b])                                   # mu.eta of the log link
        W = K @ V_inv @ K
        K_inv = np.diag(1.0 / mu[a:b])
        U += X_alt[a:b].T @ W @ K_inv @ r_null[a:b, None] / phi
        V_U += X_alt[a:b].T @ W @ X_alt[a:b]
    VarM = phi * llt_inv(V_U)
    Lp = np.eye(p_alt)[1:]
    mid_inv = llt_inv(Lp @ VarM @ Lp.T)
    full_mid = VarM @ Lp.T @ mid_inv @ Lp @ VarM
    S = float((U.T @ full_mid @ U)[0, 0])
    return {"Score_Stat": S, "DF": p_alt - 1, "P_Val": 1.0 - (1.0 - pchisq_upper(S, p_alt - 1))}


def test_dynamic_gee(counts, pt, subject, genes=None, size_factor_offset=None, cor="ar1", M=5, gee_test="wald", bias_correction=None):
    """testDynamic(is.gee = TRUE, gee.test = "wald") for one lineage; `subject` must be sorted (testDynamic.R:115)."""
    counts = np.asarray(counts)
    pt = np.asarray(pt, dtype=np.float64).reshape(-1)
    subject = np.asarray(subject)
    assert np.all(np.diff(subject) >= 0), "data must be ordered by subject"
    n_vec = np.bincount(np.unique(subject, return_inverse=True)[1])
    X, knots = candidate_knots(pt)
    off = None if size_factor_offset is None else np.log(1.0 / np.asarray(size_factor_offset, dtype=np.float64))
    G = counts.shape[0]
    genes = genes or [f"G{i}" for i in range(G)]
    sandwich = bias_correction is not None          # testDynamic.R:195,243,286
    if sandwich and cor not in ("ar1", "exchangeable"):
        raise ValueError("Unrecognized correlation structure in biasCorrectGEE().")
    out = {}
    for g in range(G):
        y = counts[g].astype(np.float64)
        rec = {"Gene": genes[g], "Lineage": "A", "Test_Stat_Type": "Wald"}
        trace = MargeTrace()
        null = None
        try:
            if y.sum() == 0:
                raise RuntimeError("all counts are zero")
            null = geem(np.ones((y.size, 1)), y, n_vec, cor, off, sandwich=sandwich)
        except (RuntimeError, np.linalg.LinAlgError, FloatingPointError):
            null = None
        try:
            if y.sum() == 0:
                raise RuntimeError("all counts are zero")
            terms = forward_pass_gee(y, X, knots, n_vec, cor, M, trace=trace)
            final = backward_pass_gee(y, X, knots, terms, n_vec, cor, trace, sandwich)
            fit = geem(build_design(final, X, knots), y, n_vec, cor, off, sandwich=sandwich)
            se = np.sqrt(np.diag(fit.var if sandwich else fit.naiv_var))
            z = fit.beta / se
            rec.update({"Model_Status": f"MARGE model OK, null model {'OK' if null is not None else 'error'}",
                        "MARGE_Summary": {"term": [term_name(t, knots, "Lineage_A") for t in final],
                                          "estimate": list(fit.beta), "std.error": list(se), "statistic": list(z),
                                          "p.value": [r_round(pnorm_two_sided(v), 8) for v in z]},       # summary.geem rounds to 8 digits
                        "alpha": fit.alpha, "phi": fit.phi, "converged": fit.converged, "niter": fit.niter})
            if gee_test == "score":
                if null is None:
                    w = {"Wald_Stat": np.nan, "DF": np.nan, "P_Val": np.nan}
                else:
                    sc = score_test_gee(build_design(final, X, knots), y, null, n_vec, cor)
                    w = {"Wald_Stat": sc["Score_Stat"], "DF": sc["DF"], "P_Val": sc["P_Val"]}
                rec["Test_Stat_Type"] = "Score"
            else:
                w = wald_test_gee(fit, None if not sandwich else bias_correct_gee(fit, bias_correction, len(n_vec), y.size))
            rec.update({"Test_Stat": w["Wald_Stat"], "Degrees_Freedom": w["DF"], "P_Val": w["P_Val"]})
        except (RuntimeError, np.linalg.LinAlgError, FloatingPointError) as e:
            rec.update({"Model_Status": f"MARGE model error, null model {'OK' if null is not None else 'error'}",
                        "MARGE_Summary": None, "Test_Stat": np.nan, "Degrees_Freedom": np.nan, "P_Val": np.nan, "error": str(e)})
        rec["Null_Summary"] = (None if null is None else
                               {"term": ["Intercept"], "estimate": [float(null.beta[0])],
                                "std.error": [float(np.sqrt((null.var if sandwich else null.naiv_var)[0, 0]))],
                                "alpha": null.alpha, "phi": null.phi})
        rec["_trace"] = trace
        out[genes[g]] = {"Lineage_A": rec}
    return out
