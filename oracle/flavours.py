"""Oracle flavours: how far do the engine's two deterministic rules sit from what the reference's third-party fitters do?

TEST INFRASTRUCTURE (like everything under oracle/).  The engine and the default oracle use
  rule 1  structural rank check      instead of RcppEigen::fastLmPure's numerical one   (score_fun_glm.R:32-34)
  rule 2  exact NB maximum likelihood instead of gamlss's RS algorithm + optimHess vcov (stat_out_score_glm_null.R:22-25,
                                                                                         backward_sel_WIC.R:44-52)
This module restates the replaced behaviour -- from the published algorithms of Eigen and gamlss, NOT from their sources, which
are not in the reference tree -- so that the effect of each rule can be measured (tools/rule_flavours.py, tests/test_flavours.py):

  rank_rule = "colpiv_qr"      Eigen::ColPivHouseholderQR on the N x (p + c) design [B XA] (column norms with Eigen's
                               down-dating rule, rank = #{|R_ii| > eps * min(N, p + c) * max|R_ii|}); when the check does not fire
                               on an exactly collinear design the "both" score comes out of a numerically singular 2 x 2
                               complement through the unchecked LLT, exactly as the reference computes it.
  rank_rule = "pair_unchecked" the accident itself: the rank check never fires on a hinge pair added to a basis that already holds a
                               full pair (only an exactly duplicated column is NA), so the pair's score is whatever the singular
                               2 x 2 complement gives -- what happened in the reference run for the 11 bundled exception genes.
  fitter = "emulate_gamlss"    RS-style alternation (mu | sigma, sigma | mu) with gamlss's start values (mu = (y + mean y)/2,
                               sigma = max((var - mean)/mean^2, 0.1)) and gamlss's STOPPING RULES (outer c.crit = 0.001 on the global
                               deviance, inner cc = 0.001): the same optimum as the exact MLE, left early.
                               vcov(type = "all") = inverse of a finite-difference Hessian
                               of -loglik over (beta, log sigma) with optimHess's central differences (ndeps = 1e-3); when that
                               fails (singular / negative variance) scLANE's fallback beta_j / Var_j (backward_sel_WIC.R:48-50).

Use:  with flavour(rank_rule="colpiv_qr"): oracle.testdynamic.test_dynamic(...)
"""
from __future__ import annotations

import contextlib
import math

import numpy as np
from scipy.special import digamma, gammaln

STATE = {"rank_rule": "structural", "fitter": "exact_mle"}


@contextlib.contextmanager
def flavour(rank_rule: str = "structural", fitter: str = "exact_mle"):
    assert rank_rule in ("structural", "colpiv_qr", "pair_unchecked") and fitter in ("exact_mle", "emulate_gamlss")
    old = dict(STATE)
    STATE.update(rank_rule=rank_rule, fitter=fitter)
    try:
        yield
    finally:
        STATE.update(old)


# ------------------------------------------------------------------------------------------------
# Eigen::ColPivHouseholderQR rank (Eigen 3.3/3.4 algorithm: LAPACK dgeqp3-style norm down-dating)
# ------------------------------------------------------------------------------------------------
def eigen_colpiv_rank(A: np.ndarray) -> int:
    A = np.array(A, dtype=np.float64, order="F", copy=True)
    rows, cols = A.shape
    size = min(rows, cols)
    norms_direct = np.sqrt(np.sum(A * A, axis=0))
    norms_updated = norms_direct.copy()
    eps = np.finfo(np.float64).eps
    threshold_helper = (norms_updated.max() * eps / rows) ** 2
    downdate_threshold = math.sqrt(eps)
    nonzero_pivots = size
    maxpivot = 0.0
    diag = np.zeros(size)
    for k in range(size):
        j = k + int(np.argmax(norms_updated[k:]))
        biggest_sq = norms_updated[j] ** 2
        if nonzero_pivots == size and biggest_sq < threshold_helper * (rows - k):
            nonzero_pivots = k
        if j != k:
            A[:, [k, j]] = A[:, [j, k]]
            norms_updated[[k, j]] = norms_updated[[j, k]]
            norms_direct[[k, j]] = norms_direct[[j, k]]
        # Householder on A[k:, k]
        x = A[k:, k]
        tail_sq = float(np.sum(x[1:] ** 2))
        c0 = float(x[0])
        if tail_sq <= np.finfo(np.float64).tiny:
            tau, beta = 0.0, c0
            v = np.zeros(x.size)
        else:
            beta = math.sqrt(c0 * c0 + tail_sq)
            if c0 >= 0:
                beta = -beta
            v = x / (c0 - beta)
            tau = (beta - c0) / beta
        v[0] = 1.0
        A[k, k] = beta
        diag[k] = beta
        if abs(beta) > maxpivot:
            maxpivot = abs(beta)
        if tau != 0.0 and k + 1 < cols:
            sub = A[k:, k + 1:]
            sub -= tau * np.outer(v, v @ sub)
        A[k + 1:, k] = 0.0
        for jj in range(k + 1, cols):
            if norms_updated[jj] != 0.0:
                t = abs(A[k, jj]) / norms_updated[jj]
                t = (1.0 + t) * (1.0 - t)
                t = 0.0 if t < 0.0 else t
                t2 = t * (norms_updated[jj] / norms_direct[jj]) ** 2
                if t2 <= downdate_threshold:
                    norms_direct[jj] = float(np.sqrt(np.sum(A[k + 1:, jj] ** 2)))
                    norms_updated[jj] = norms_direct[jj]
                else:
                    norms_updated[jj] *= math.sqrt(t)
    thr = maxpivot * eps * size
    return int(np.sum(np.abs(diag[:nonzero_pivots]) > thr))


def rank_deficient_colpiv(B: np.ndarray, XA: np.ndarray) -> bool:
    """fastLmPure(cbind(B, XA), Y) reports an NA coefficient <=> the column-pivoted QR finds rank < p + c."""
    A = np.concatenate([B, XA], axis=1)
    return eigen_colpiv_rank(A) < A.shape[1]


# ------------------------------------------------------------------------------------------------
# gamlss(Y ~ B - 1, family = "NBI") by the RS algorithm (defaults), and vcov(type = "all")
# ------------------------------------------------------------------------------------------------
def _nbi_logpdf(y, mu, sigma):
    if sigma < 1e-4:                      # gamlss.dist dNBI switches to the Poisson density
        return y * np.log(mu) - mu - gammaln(y + 1.0)
    th = 1.0 / sigma
    return gammaln(y + th) - gammaln(th) - gammaln(y + 1.0) + y * np.log(sigma * mu) - (y + th) * np.log1p(sigma * mu)


def _wls(X, w, z):
    Xw = X * w[:, None]
    return np.linalg.solve(X.T @ Xw, Xw.T @ z)


def gamlss_rs_nbi(X, y, c_crit=1e-3, n_cyc=20, cc=1e-3, cyc=50):
    """Returns (beta, sigma, mu, G, XtWX of the last mu step)."""
    y = np.asarray(y, dtype=np.float64)
    ybar = float(y.mean())
    mu = (y + ybar) / 2.0
    sigma = max((float(y.var(ddof=1)) - ybar) / (ybar * ybar), 0.1)
    eta = np.log(mu)
    eta_s = math.log(sigma)
    beta = np.zeros(X.shape[1])
    G_old = -2.0 * float(np.sum(_nbi_logpdf(y, mu, sigma)))
    G = G_old
    xtwx = None
    for _ in range(n_cyc):
        # ---- mu | sigma: glim.fit
        dev_old = -2.0 * float(np.sum(_nbi_logpdf(y, mu, sigma)))
        for itn in range(1, cyc + 1):
            w = mu / (1.0 + mu * sigma)                     # -d2ldm2 / dr^2, dr = 1 / mu
            w = np.maximum(w, 1e-15)
            z = eta + (y - mu) / mu
            beta_new = _wls(X, w, z)
            eta_new = X @ beta_new
            mu_new = np.exp(eta_new)
            dev = -2.0 * float(np.sum(_nbi_logpdf(y, mu_new, sigma)))
            step = 1.0
            while itn >= 2 and dev > dev_old + 1e-12 and step > 1e-3:     # autostep (from the 2nd inner iteration on): halve towards the previous predictor
                step *= 0.5
                eta_h = eta + step * (eta_new - eta)
                mu_new = np.exp(eta_h)
                dev = -2.0 * float(np.sum(_nbi_logpdf(y, mu_new, sigma)))
                beta_new = beta + step * (beta_new - beta) if np.any(beta) else beta_new
                eta_new = eta_h
            beta, eta, mu = beta_new, eta_new, mu_new
            xtwx = (X * w[:, None]).T @ X
            if abs(dev_old - dev) < cc:
                dev_old = dev
                break
            dev_old = dev
        # ---- sigma | mu: scoring steps on log sigma until the deviance moves by less than cc.  (gamlss's NBI uses the
        # quasi-Newton weight -dldd^2 here; an emulation with that weight -- written from memory -- landed far from the
        # reference's bundled results (35 / 100 term sets), so the step itself is a plain safeguarded Newton step and only
        # the STOPPING RULES are gamlss's: what this flavour measures is the effect of stopping early.)
        from .nbfit import _dl_dlogsigma
        for _ in range(cyc):
            d1, d2 = _dl_dlogsigma(y, mu, sigma)
            step_s = -d1 / d2 if d2 < 0 else math.copysign(1.0, d1)
            step_s = max(-2.0, min(2.0, step_s))
            eta_s_new = min(max(eta_s + step_s, -23.0), 23.0)
            sigma_new = math.exp(eta_s_new)
            dev = -2.0 * float(np.sum(_nbi_logpdf(y, mu, sigma_new)))
            step = 1.0
            while dev > dev_old + 1e-12 and step > 1e-3:
                step *= 0.5
                sigma_new = math.exp(eta_s + step * (eta_s_new - eta_s))
                dev = -2.0 * float(np.sum(_nbi_logpdf(y, mu, sigma_new)))
            eta_s, sigma = math.log(sigma_new), sigma_new
            if abs(dev_old - dev) < cc:
                dev_old = dev
                break
            dev_old = dev
        G = -2.0 * float(np.sum(_nbi_logpdf(y, mu, sigma)))
        if abs(G_old - G) < c_crit:
            break
        G_old = G
    return beta, sigma, mu, G, xtwx


def _optimhess(f, x, ndeps=1e-3):
    """stats::optimHess without an analytic gradient: central-difference gradient, central-difference Hessian of it, symmetrised."""
    n = x.size

    def grad(xx):
        g = np.empty(n)
        for i in range(n):
            e = np.zeros(n)
            e[i] = ndeps
            g[i] = (f(xx + e) - f(xx - e)) / (2.0 * ndeps)
        return g

    H = np.empty((n, n))
    for i in range(n):
        e = np.zeros(n)
        e[i] = ndeps
        H[i] = (grad(x + e) - grad(x - e)) / (2.0 * ndeps)
    return 0.5 * (H + H.T)


def gamlss_wald(X, y):
    """backward_sel_WIC.R:44-52 with the gamlss fit emulated: (coef / se)^2 of the non-intercept coefficients from
    vcov(type = "all"), or scLANE's fallback when vcov() fails.  Returns (wald vector, used_fallback)."""
    beta, sigma, mu, _, xtwx = gamlss_rs_nbi(X, y)
    p = X.shape[1]

    def negll(par):
        return -float(np.sum(_nbi_logpdf(y, np.exp(X @ par[:p]), math.exp(par[p]))))

    par = np.concatenate([beta, [math.log(sigma)]])
    ok = True
    try:
        H = _optimhess(negll, par)
        cov = np.linalg.solve(H, np.eye(p + 1))
        if not np.all(np.isfinite(cov)) or np.any(np.diag(cov) < 0.0):
            ok = False
    except np.linalg.LinAlgError:
        ok = False
    if ok:
        se = np.sqrt(np.diag(cov))
        return ((par / se)[1:p]) ** 2, False
    var = np.diag(np.linalg.inv(xtwx))                       # chol2inv(mu.qr): (B'WB)^-1 of the last mu step
    return (beta / var)[1:], True
