"""Hinge functions and the gene-independent candidate-knot set (R/tp1.R, R/tp2.R,
R/min_span.R, R/max_span.R, R/marge2.R:150-173)."""
from __future__ import annotations

import math

import numpy as np

from .rmath import quantile7, r_round, r_seq_len_out


def tp1(x: np.ndarray, t: float) -> np.ndarray:
    """tp1.R:14-19: (x - t) * (x > t)."""
    return (x - t) * (x > t)


def tp2(x: np.ndarray, t: float) -> np.ndarray:
    """tp2.R:14-19: (t - x) * (x < t)."""
    return (t - x) * (x < t)


def min_span(x_red: np.ndarray, q: int = 1, minspan: int | None = None, alpha: float = 0.05) -> np.ndarray:
    """min_span.R:15-36.  Every (minspan+1)-th order statistic of the (duplicated) sorted
    values, starting from the minimum; the last index may run past N (R yields NA, which the
    later intersect() discards -> represented here as NaN and dropped by the caller)."""
    n = x_red.size
    x = np.sort(x_red)
    if minspan is None:
        minspan = int(np.rint(-math.log2(-(1.0 / (q * n)) * math.log(1.0 - alpha)) / 2.5))
    out = [x[0]]
    cc = 1
    while cc + minspan <= n:
        idx = cc + minspan + 1  # 1-based
        out.append(x[idx - 1] if idx <= n else np.nan)
        cc += minspan + 1
    out = np.array(out)
    # unique(), keeping first occurrences (values are ascending so order is preserved)
    _, first = np.unique(out, return_index=True)
    return out[np.sort(first)]


def max_span(x_red: np.ndarray, q: int = 1, alpha: float = 0.05) -> np.ndarray:
    """max_span.R:14-26: unique sorted values minus `maxspan` values at each end."""
    x = np.unique(x_red)
    n = x.size
    maxspan = int(np.rint(3.0 - math.log2(alpha / q)))
    drop = set(range(1, maxspan + 1)) | set(range(int(math.floor(n - maxspan + 1)), n + 1))
    keep = [i for i in range(1, n + 1) if i not in drop]
    x_new = x[np.array(keep, dtype=int) - 1] if keep else x
    return x_new


def candidate_knots(x_pred: np.ndarray, q: int = 1, approx_knot: bool = True, n_knot_max: int = 25,
                    minspan: int | None = None):
    """marge2.R:152-173.  Returns (X, X_red): the rounded predictor the hinges are built
    from and the candidate knots in evaluation order."""
    X = r_round(x_pred, 4)
    r1 = min_span(X, q=q, minspan=minspan)
    r2 = max_span(X, q=q)
    r1 = r1[~np.isnan(r1)]
    x_red = r1[np.isin(r1, r2)]  # intersect(): order of the first argument, unique
    if approx_knot:
        q05 = quantile7(x_pred, 0.05)
        q95 = quantile7(x_pred, 0.95)
        x_red = x_red[(x_red > q05) & (x_red < q95)]
        if x_red.size > n_knot_max:
            x_red = r_round(r_seq_len_out(x_red.min(), x_red.max(), n_knot_max), 4)
    return X, x_red
