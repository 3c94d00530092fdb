"""Scalar R semantics the path depends on (base R / nmath, restated; not in the reference tree)."""
from __future__ import annotations

import math

import numpy as np
from scipy import stats as _st


def r_round(x, digits: int = 0):
    """R >= 4.0.0 `round(x, digits)` (nmath/fround.c, "round to nearest representable,
    ties to even"): candidates floor(x*10^d)/10^d and ceil(x*10^d)/10^d, pick the closer
    one to x; on an exact tie pick the one whose scaled integer is even.
    Used at marge2.R:154,165 (round(x,4)), :374-576 (round(score,6)), :681 (round(score2,4)),
    stat_out.R:41-43 (round(.,10))."""
    a = np.asarray(x, dtype=np.float64)
    scalar = a.ndim == 0
    a = np.atleast_1d(a).copy()
    out = a.copy()
    fin = np.isfinite(a) & (a != 0.0)
    if digits == 0:
        out[fin] = np.rint(a[fin])
    else:
        p10 = 10.0 ** digits
        s = np.sign(a[fin])
        v = np.abs(a[fin])
        x10 = p10 * v
        i10 = np.floor(x10)
        xd = i10 / p10
        xu = np.ceil(x10) / p10
        du = xu - v
        dd = v - xd
        even = np.fmod(i10, 2.0) == 0.0
        take_d = (dd < du) | ((dd == du) & even)
        r = np.where(take_d, xd, xu)
        # values with more magnitude than 15 significant digits allow are returned unchanged
        big = x10 >= 2.0 ** 52
        r = np.where(big, v, r)
        out[fin] = s * r
    return float(out[0]) if scalar else out


def r_signif(x: float, digits: int = 4) -> float:
    """R `signif(x, digits)` (nmath/fprec.c) for finite doubles; used for knot labels
    at marge2.R:593-594."""
    if x == 0.0 or not math.isfinite(x):
        return x
    e10 = digits - 1 - math.floor(math.log10(abs(x)))
    if e10 >= 0:
        p = 10.0 ** e10
        return float(np.rint(x * p) / p)
    p = 10.0 ** (-e10)
    return float(np.rint(x / p) * p)


def r_format_num(x: float) -> str:
    """as.character(<double>) with 15 significant digits (what paste() does at marge2.R:593)."""
    s = f"{x:.15g}"
    return s


def quantile7(x: np.ndarray, p: float) -> float:
    """stats::quantile(type = 7) (the default) as called at marge2.R:158-159."""
    xs = np.sort(np.asarray(x, dtype=np.float64))
    n = xs.size
    index = 1.0 + max(n - 1, 0) * p
    lo = int(math.floor(index))
    hi = int(math.ceil(index))
    q = xs[lo - 1]
    h = index - lo
    if index > lo and xs[hi - 1] != q:
        q = (1.0 - h) * q + h * xs[hi - 1]
    return float(q)


def r_seq_len_out(frm: float, to: float, n: int) -> np.ndarray:
    """base::seq.default(from, to, length.out = n): c(from, from + (1:(n-2)) * by, to)."""
    if n == 1:
        return np.array([frm])
    if n == 2:
        return np.array([frm, to])
    if frm == to:
        return np.full(n, frm)
    by = (to - frm) / (n - 1)
    mid = frm + np.arange(1, n - 1, dtype=np.float64) * by
    return np.concatenate([[frm], mid, [to]])


def pchisq_upper(q: float, df: float) -> float:
    """stats::pchisq(q, df, lower.tail = FALSE) (modelLRT.R:41-43)."""
    if q is None or (isinstance(q, float) and math.isnan(q)):
        return float("nan")
    return float(_st.chi2.sf(q, df))


def pnorm_two_sided(z: float) -> float:
    """2 * pnorm(-abs(z)) (summary.glm coefficient table; broom tidy p.value)."""
    return float(2.0 * _st.norm.cdf(-abs(z)))


def p_adjust_bh(p: np.ndarray) -> np.ndarray:
    """stats::p.adjust(method = "fdr"/"BH"); NA (NaN) entries are excluded from n and
    stay NA (getResultsDE.R:39)."""
    p = np.asarray(p, dtype=np.float64)
    out = np.full(p.shape, np.nan)
    ok = ~np.isnan(p)
    pv = p[ok]
    n = pv.size
    if n == 0:
        return out
    o = np.argsort(-pv, kind="stable")  # decreasing
    ro = np.argsort(o, kind="stable")
    i = np.arange(n, 0, -1, dtype=np.float64)
    adj = np.minimum(1.0, np.minimum.accumulate(n / i * pv[o]))[ro]
    out[ok] = adj
    return out
