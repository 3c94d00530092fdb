"""Restatement of the reference's three RcppEigen helpers (the only native code in the
reference: src/eigenMapMatMult.cpp:6-15, src/eigenMapInvert.cpp:6-11,
src/eigenMapPseudoInverse.cpp:6-21).  Eigen itself is third-party and not in the
reference tree; its LLT (unblocked path, used for n < 32) is restated from the
published algorithm, including its behaviour on non-positive-definite input: the
factorisation stops at the first non-positive pivot, leaves the rest of the matrix
untouched, and the reference never checks info(), so solve() proceeds on whatever is there."""
from __future__ import annotations

import numpy as np


def eigen_map_mat_mult(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """eigenMapMatMult.cpp:13-14: C.noalias() = A * B."""
    return np.asarray(a, dtype=np.float64) @ np.asarray(b, dtype=np.float64)


def _llt_inplace_unblocked(m: np.ndarray) -> int:
    n = m.shape[0]
    for k in range(n):
        x = m[k, k]
        if k > 0:
            x = x - float(np.dot(m[k, :k], m[k, :k]))
        if x <= 0.0:  # NaN falls through, exactly as `if (x<=RealScalar(0)) return k;`
            return k
        x = np.sqrt(x)
        m[k, k] = x
        rs = n - k - 1
        if k > 0 and rs > 0:
            m[k + 1:, k] -= m[k + 1:, :k] @ m[k, :k]
        if rs > 0:
            m[k + 1:, k] /= x
    return -1


def eigen_map_matrix_invert(a: np.ndarray) -> np.ndarray:
    """eigenMapInvert.cpp:10: A.llt().solve(Identity).  Only the lower triangle of A is read."""
    a = np.array(a, dtype=np.float64, copy=True)
    n = a.shape[0]
    _llt_inplace_unblocked(a)
    L = np.tril(a)
    with np.errstate(all="ignore"):
        # forward then backward substitution against identity, no pivot checks (as Eigen)
        y = np.zeros((n, n))
        for c in range(n):
            for i in range(n):
                s = (1.0 if i == c else 0.0) - float(np.dot(L[i, :i], y[:i, c]))
                y[i, c] = s / L[i, i]
        xinv = np.zeros((n, n))
        for c in range(n):
            for i in range(n - 1, -1, -1):
                s = y[i, c] - float(np.dot(L[i + 1:, i], xinv[i + 1:, c]))
                xinv[i, c] = s / L[i, i]
    return xinv


def eigen_map_pseudo_inverse(a: np.ndarray, tolerance: float = 1e-6) -> np.ndarray:
    """eigenMapPseudoInverse.cpp:13-20: thin SVD, invert singular values > tolerance (absolute)."""
    a = np.asarray(a, dtype=np.float64)
    u, s, vt = np.linalg.svd(a, full_matrices=False)
    sinv = np.where(s > tolerance, 1.0 / np.where(s > tolerance, s, 1.0), 0.0)
    return (vt.T * sinv) @ u.T
