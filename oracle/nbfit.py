"""Negative-binomial fits on the path (dense fp64 restatements).

Two kinds of fit exist on the reference's path:

1. `glm_nb` -- MASS::glm.nb(..., method = "glm.fit2", init.theta = 1, link = log), the
   FINAL alternative fit (marge2.R:841-847) and the intercept-only NULL fit
   (testDynamic.R:254-260).  MASS / glm2 are third-party (not in the reference tree);
   the iteration schedule below restates their published R sources (MASS::glm.nb,
   MASS::theta.ml, MASS::negative.binomial, stats::glm.fit + the glm2 step-halving
   addition).  Pinned against the bundled golden: tests/test_oracle_golden.py.

2. `nb_mle` / `wald_nb` -- stand-in for gamlss::gamlss(Y ~ B - 1, family = "NBI")
   (stat_out_score_glm_null.R:22-25, backward_sel_WIC.R:44-47) and for
   vcov(fit, type = "all") (backward_sel_WIC.R:47-52).  gamlss is third-party and its RS
   algorithm stops early (c.crit = 0.001); the engine replaces it by the exact NB
   maximum-likelihood fit (what RS converges to) and the analytic joint observed
   information over (beta, log sigma) (what optimHess approximates).  Documented rule for
   the Poisson regime (DESIGN.md "Deterministic rules", rule 2): when the MLE of sigma is
   below SIGMA_POISSON = 1e-4 -- the point where gamlss.dist's NBI density switches to
   Poisson -- the fit is the Poisson MLE (sigma := 0) and backward_sel_WIC takes its
   `try-error` fallback branch (backward_sel_WIC.R:48-50), i.e. "Wald" = beta_j / Var_j.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
from scipy.special import digamma, gammaln, polygamma

SIGMA_POISSON = 1e-4  # gamlss.dist::dNBI switches to the Poisson density below this sigma
EPS_GLM = 1e-8        # glm.control(epsilon)
MAXIT_GLM = 25        # glm.control(maxit)


# --------------------------------------------------------------------------------------
# small dense helpers
# --------------------------------------------------------------------------------------
def _wls(X: np.ndarray, w: np.ndarray, z: np.ndarray):
    """Weighted least squares via the normal equations (Cholesky); returns (beta, XtWX)."""
    Xw = X * w[:, None]
    G = X.T @ Xw
    b = Xw.T @ z
    beta = np.linalg.solve(G, b)
    return beta, G


# --------------------------------------------------------------------------------------
# 1. MASS::glm.nb restatement
# --------------------------------------------------------------------------------------
def nb_dev_resids_sum(y, mu, th):
    """sum(negative.binomial(th)$dev.resids(y, mu, 1)):
    2 * (y * log(pmax(1, y) / mu) - (y + th) * log((y + th) / (mu + th)))."""
    return float(np.sum(2.0 * (y * np.log(np.maximum(1.0, y) / mu) - (y + th) * np.log((y + th) / (mu + th)))))


def nb_loglik(y, mu, th):
    """glm.nb's internal loglik(): note the `mu + (y == 0)` guard."""
    return float(np.sum(gammaln(th + y) - gammaln(th) - gammaln(y + 1.0) + th * math.log(th)
                        + y * np.log(mu + (y == 0)) - (th + y) * np.log(th + mu)))


@dataclass
class GlmFit:
    coef: np.ndarray
    mu: np.ndarray
    eta: np.ndarray
    deviance: float
    XtWX: np.ndarray          # of the last weighted least-squares solve (what summary.glm's qr holds)
    iters: int
    converged: bool
    rank: int


def glm_fit2_nb(X, y, offset, th, etastart=None, epsilon=EPS_GLM, maxit=MAXIT_GLM) -> GlmFit:
    """stats::glm.fit with family negative.binomial(theta = th, link = log) plus glm2's
    step-halving on increasing deviance.  Start: family$initialize => mustart = y + (y==0)/6
    unless `etastart` is given (glm.nb passes etastart = log(mu) in its alternation loop)."""
    n, p = X.shape
    if etastart is None:
        mustart = y + (y == 0) / 6.0
        eta = np.log(mustart)
    else:
        eta = np.array(etastart, dtype=np.float64, copy=True)
    mu = np.exp(eta)
    devold = nb_dev_resids_sum(y, mu, th)
    coef = None
    coefold = None
    conv = False
    G = None
    it = 0
    for it in range(1, maxit + 1):
        varmu = mu + mu * mu / th
        mu_eta = mu
        z = (eta - offset) + (y - mu) / mu_eta
        w = (mu_eta * mu_eta) / varmu           # = sqrt(.)^2, the WLS weights
        start, G = _wls(X, w, z)
        eta = X @ start + offset
        mu = np.exp(eta)
        dev = nb_dev_resids_sum(y, mu, th)
        if not np.isfinite(dev):
            if coefold is None:
                raise FloatingPointError("no valid set of coefficients has been found")
            ii = 1
            while not np.isfinite(dev):
                if ii > maxit:
                    raise FloatingPointError("inner loop 1; cannot correct step size")
                ii += 1
                start = (start + coefold) / 2.0
                eta = X @ start + offset
                mu = np.exp(eta)
                dev = nb_dev_resids_sum(y, mu, th)
        # glm2: step-halving when the deviance increased
        if (dev - devold) / (0.1 + abs(dev)) >= epsilon and it > 1 and coefold is not None:
            ii = 1
            while (dev - devold) / (0.1 + abs(dev)) > -epsilon:
                if ii > maxit:
                    break
                ii += 1
                start = (start + coefold) / 2.0
                eta = X @ start + offset
                mu = np.exp(eta)
                dev = nb_dev_resids_sum(y, mu, th)
        if abs(dev - devold) / (abs(dev) + 0.1) < epsilon:
            conv = True
            coef = start
            break
        devold = dev
        coef = coefold = start
    return GlmFit(coef=coef, mu=mu, eta=eta, deviance=dev, XtWX=G, iters=it, converged=conv, rank=p)


def theta_ml(y, mu, limit=MAXIT_GLM, eps=np.finfo(np.float64).eps ** 0.25):
    """MASS::theta.ml(y, mu, n, limit, eps).  Returns (theta, SE, warn, iterations)."""
    n = y.size
    t0 = n / float(np.sum((y / mu - 1.0) ** 2))
    it = 0
    dl = 1.0
    info = np.nan
    while True:
        it += 1
        if not (it < limit and abs(dl) > eps):
            break
        t0 = abs(t0)
        score = float(np.sum(digamma(t0 + y) - digamma(t0) + math.log(t0) + 1.0 - np.log(t0 + mu)
                             - (y + t0) / (mu + t0)))
        info = float(np.sum(-polygamma(1, t0 + y) + polygamma(1, t0) - 1.0 / t0 + 2.0 / (mu + t0)
                            - (y + t0) / (mu + t0) ** 2))
        dl = score / info
        t0 = t0 + dl
    warn = None
    if t0 < 0:
        t0 = 0.0
        warn = "estimate truncated at zero"
    if it == limit:
        warn = "iteration limit reached"
    with np.errstate(all="ignore"):
        se = math.sqrt(1.0 / info) if info == info and info > 0 else float("nan")
    return t0, se, warn, it


@dataclass
class NegBinFit:
    coef: np.ndarray
    se: np.ndarray
    theta: float
    se_theta: float
    loglik: float
    deviance: float
    mu: np.ndarray
    eta: np.ndarray
    cov_unscaled: np.ndarray
    converged: bool
    th_warn: str | None
    alternations: int
    df_residual: int
    irls_total: int = 0
    theta_iters_total: int = 0


def glm_nb(X, y, offset=None, init_theta=1.0, epsilon=EPS_GLM, maxit=MAXIT_GLM) -> NegBinFit:
    """MASS::glm.nb(formula, method = "glm.fit2", init.theta = 1, link = log) as called at
    marge2.R:841-847 and testDynamic.R:254-260.  Note the schedule quirk that must be
    replicated: inside the alternation loop theta.ml() is evaluated at the mu of the
    PREVIOUS alternation, and the returned glm object (coefficients, deviance, qr) is the
    one fitted under the previous theta."""
    y = np.asarray(y, dtype=np.float64)
    n, p = X.shape
    offset = np.zeros(n) if offset is None else np.asarray(offset, dtype=np.float64)
    fit = glm_fit2_nb(X, y, offset, init_theta, None, epsilon, maxit)
    irls = fit.iters
    mu = fit.mu
    th, se_th, warn, tit = theta_ml(y, mu, limit=maxit)
    thits = tit
    it = 0
    d1 = math.sqrt(2.0 * max(1, n - p))
    d2 = dl = 1.0
    Lm = nb_loglik(y, mu, th)
    Lm0 = Lm + 2.0 * d1
    while True:
        it += 1
        if not (it <= maxit and (abs(Lm0 - Lm) / d1 + abs(dl) / d2) > epsilon):
            break
        eta = np.log(mu)
        fit = glm_fit2_nb(X, y, offset, th, eta, epsilon, maxit)
        irls += fit.iters
        t0 = th
        th, se_th, warn, tit = theta_ml(y, mu, limit=maxit)   # stale mu, as in MASS
        thits += tit
        mu = fit.mu
        dl = t0 - th
        Lm0 = Lm
        Lm = nb_loglik(y, mu, th)
    th_warn = warn
    if it > maxit:
        th_warn = "alternation limit reached"
    cov = np.linalg.inv(fit.XtWX)
    se = np.sqrt(np.diag(cov))
    return NegBinFit(coef=fit.coef, se=se, theta=th, se_theta=se_th, loglik=Lm, deviance=fit.deviance,
                     mu=mu, eta=fit.eta, cov_unscaled=cov, converged=fit.converged, th_warn=th_warn,
                     alternations=it - 1, df_residual=n - p, irls_total=irls, theta_iters_total=thits)


# --------------------------------------------------------------------------------------
# 2. exact NB MLE (gamlss NBI stand-in) and the joint-information Wald statistics
# --------------------------------------------------------------------------------------
def _nbi_loglik(y, eta, sigma):
    """NBI log-likelihood in (eta = log mu, sigma = 1/theta); Poisson when sigma == 0."""
    mu = np.exp(eta)
    if sigma == 0.0:
        return float(np.sum(y * eta - mu - gammaln(y + 1.0)))
    th = 1.0 / sigma
    return float(np.sum(gammaln(y + th) - gammaln(th) - gammaln(y + 1.0) + y * (math.log(sigma) + eta)
                        - (y + th) * np.log1p(sigma * mu)))


def _irls_beta(X, y, sigma, beta, tol=1e-13, maxit=100):
    """Fisher scoring for beta at fixed sigma (w = mu/(1+mu*sigma)); iterate until the
    log-likelihood stops increasing (relative change < tol), with step-halving.  The normal
    equations use w*z = w*eta + (y-mu)/(1+mu*sigma), which stays finite when mu underflows
    (quasi-separated low-count genes, where a hinge coefficient drifts to -inf while the
    likelihood converges)."""
    eta = X @ beta
    ll = _nbi_loglik(y, eta, sigma)
    for _ in range(maxit):
        mu = np.exp(eta)
        den = 1.0 + mu * sigma
        w = mu / den
        wz = w * eta + (y - mu) / den
        G = X.T @ (X * w[:, None])
        try:
            new = np.linalg.solve(G, X.T @ wz)
        except np.linalg.LinAlgError:
            break
        if not np.all(np.isfinite(new)):
            break
        eta_n = X @ new
        ll_n = _nbi_loglik(y, eta_n, sigma)
        k = 0
        while not (ll_n >= ll - 1e-12 * (abs(ll) + 1.0)) and k < 40:
            new = 0.5 * (new + beta)
            eta_n = X @ new
            ll_n = _nbi_loglik(y, eta_n, sigma)
            k += 1
        if not (ll_n >= ll - 1e-12 * (abs(ll) + 1.0)):
            break
        done = abs(ll_n - ll) <= tol * (abs(ll) + 1.0)
        beta, eta, ll = new, eta_n, ll_n
        if done:
            break
    return beta, eta, ll


def _dl_dlogsigma(y, mu, sigma):
    """First and second derivative of the log-likelihood w.r.t. s = log sigma at fixed mu."""
    th = 1.0 / sigma
    g = float(np.sum(digamma(y + th) - digamma(th) + math.log(th) + 1.0 - np.log(th + mu) - (y + th) / (th + mu)))
    h = float(np.sum(polygamma(1, y + th) - polygamma(1, th) + 1.0 / th - 2.0 / (th + mu) + (y + th) / (th + mu) ** 2))
    d1 = -th * g
    d2 = th * g + th * th * h
    return d1, d2


@dataclass
class NbMle:
    beta: np.ndarray
    sigma: float            # 0.0 in the Poisson regime
    mu: np.ndarray
    loglik: float
    poisson: bool
    passes: int = 0
    info: dict = field(default_factory=dict)


def nb_mle(X, y, beta0=None) -> NbMle:
    """Exact maximum-likelihood NBI fit of y ~ X - 1 (log link, no offset), sigma >= SIGMA_POISSON,
    by alternating Fisher scoring (beta | sigma) with a safeguarded Newton on log sigma (sigma | mu),
    then polishing with joint Newton steps.  If the constrained optimum sits on the boundary
    sigma = SIGMA_POISSON the fit is declared Poisson: sigma := 0 and beta := Poisson MLE."""
    y = np.asarray(y, dtype=np.float64)
    n, p = X.shape
    ybar = float(y.mean())
    if beta0 is None:
        # least-squares start on log(y + 0.5-ish) keeps hinge coefficients sane
        z0 = np.log((y + ybar) / 2.0 + 1e-8) if ybar > 0 else np.zeros(n)
        beta = np.linalg.lstsq(X, z0, rcond=None)[0]
    else:
        beta = np.array(beta0, dtype=np.float64)
    var = float(y.var(ddof=1)) if n > 1 else 0.0
    sigma = max((var - ybar) / (ybar * ybar), 0.1) if ybar > 0 else 0.1
    s = math.log(sigma)
    s_min = math.log(SIGMA_POISSON)
    passes = 0
    on_boundary = False
    for outer in range(200):
        beta, eta, ll = _irls_beta(X, y, math.exp(s), beta)
        mu = np.exp(eta)
        s_old = s
        # Newton on s with bisection-style safeguards, sigma clamped at the Poisson switch
        for _ in range(100):
            d1, d2 = _dl_dlogsigma(y, mu, math.exp(s))
            passes += 1
            if d2 < 0:
                step = -d1 / d2
            else:
                step = math.copysign(1.0, d1)
            step = max(-2.0, min(2.0, step))
            s_new = max(s + step, s_min)
            if abs(s_new - s) < 1e-12:
                s = s_new
                break
            s = s_new
        on_boundary = (s <= s_min + 1e-12)
        if abs(s - s_old) < 1e-10:
            break
    sigma = math.exp(s)
    if on_boundary:
        d1, _ = _dl_dlogsigma(y, mu, sigma)
        if d1 < 0.0:
            beta, eta, ll = _irls_beta(X, y, 0.0, beta)
            return NbMle(beta=beta, sigma=0.0, mu=np.exp(eta), loglik=ll, poisson=True, passes=passes)
    # joint Newton polish on (beta, log sigma)
    for _ in range(4):
        H, grad = nb_joint_hessian(X, y, beta, sigma, want_grad=True)
        try:
            step = np.linalg.solve(H, -grad)
        except np.linalg.LinAlgError:
            break
        if not np.all(np.isfinite(step)) or np.max(np.abs(step)) > 0.1:
            break
        beta = beta + step[:p]
        sigma = sigma * math.exp(step[p])
        if np.max(np.abs(step)) < 1e-13:
            break
    eta = X @ beta
    return NbMle(beta=beta, sigma=sigma, mu=np.exp(eta), loglik=_nbi_loglik(y, eta, sigma), poisson=False,
                 passes=passes)


def nb_joint_hessian(X, y, beta, sigma, want_grad=False):
    """Hessian of the NBI log-likelihood w.r.t. (beta, s = log sigma) -- analytic form of what
    gamlss's vcov(type = "all") obtains by finite differences (optimHess)."""
    p = X.shape[1]
    eta = X @ beta
    mu = np.exp(eta)
    th = 1.0 / sigma
    den = 1.0 + sigma * mu
    d_eta = (y - mu) / den
    d2_eta = -mu * (1.0 + sigma * y) / den ** 2
    d2_eta_s = -sigma * mu * (y - mu) / den ** 2
    g = digamma(y + th) - digamma(th) + math.log(th) + 1.0 - np.log(th + mu) - (y + th) / (th + mu)
    h = polygamma(1, y + th) - polygamma(1, th) + 1.0 / th - 2.0 / (th + mu) + (y + th) / (th + mu) ** 2
    d_s = -th * g
    d2_s = th * g + th * th * h
    H = np.zeros((p + 1, p + 1))
    H[:p, :p] = X.T @ (X * d2_eta[:, None])
    H[:p, p] = H[p, :p] = X.T @ d2_eta_s
    H[p, p] = float(np.sum(d2_s))
    if want_grad:
        grad = np.concatenate([X.T @ d_eta, [float(np.sum(d_s))]])
        return H, grad
    return H


def wald_nb(X, y, fit: NbMle | None = None) -> tuple[np.ndarray, NbMle]:
    """backward_sel_WIC.R:44-52 (GLM branch): Wald statistics of the non-intercept columns.
    Regular case: (coef / se)^2 with se from the inverse joint observed information over
    (beta, log sigma).  Poisson regime (vcov() failure): the reference's fallback
    coef / sqrt(diag(chol2inv(mu.qr)))^2 = beta_j / Var_j with Var = diag((X'WX)^-1), W = mu."""
    if fit is None:
        fit = nb_mle(X, y)
    p = X.shape[1]
    if fit.poisson:
        G = X.T @ (X * fit.mu[:, None])
        cov = np.linalg.inv(G)
        wald = fit.beta / np.diag(cov)
        return wald[1:], fit
    H = nb_joint_hessian(X, y, fit.beta, fit.sigma)
    cov = np.linalg.inv(-H)
    with np.errstate(all="ignore"):
        se = np.sqrt(np.diag(cov))
    wald = (fit.beta / se[:p]) ** 2
    return wald[1:], fit
