"""Deterministic synthetic inputs of the shapes BASELINE.json names.

pt: x_i = (rank_i - 0.5)/N after a seeded permutation; genes: 50 % dynamic with up to two hinges
(log mu = b0 + sum a_k (x - tau_k)+-), 50 % static; theta ~ logU(0.5, 20) with a 5 % near-Poisson
stratum (theta = 1e3); per-cell depth factor exp(N(0, 0.3^2)); Y ~ NB(mean = mu * d, size = theta), int32.
Seed 312 is the reference's default random.seed (testDynamic.R:75)."""
from __future__ import annotations

import numpy as np


def make_pt(n_cells: int, rng: np.random.Generator) -> np.ndarray:
    perm = rng.permutation(n_cells)
    x = np.empty(n_cells)
    x[perm] = (np.arange(n_cells) + 0.5) / n_cells
    return x


def make_counts(n_genes: int, n_cells: int, seed: int = 312, n_lineages: int = 1, chunk: int = 256):
    """Returns (counts int32 [n_genes, n_cells] gene-major, pt float64 [n_cells, n_lineages], depth [n_cells])."""
    rng = np.random.default_rng(seed)
    if n_lineages == 1:
        pt = make_pt(n_cells, rng)[:, None]
    else:
        # two lineages sharing a root: v < 0.6 -> A, v > 0.4 -> B (20 % of cells in both)
        v = rng.random(n_cells)
        pt = np.full((n_cells, n_lineages), np.nan)
        for l in range(n_lineages):
            member = (v < 0.6) if l == 0 else (v > 0.4)
            idx = np.flatnonzero(member)
            pt[idx, l] = make_pt(idx.size, rng)
    depth = np.exp(rng.normal(0.0, 0.3, n_cells))
    x = np.nan_to_num(pt[:, 0], nan=0.5) if n_lineages == 1 else np.nanmean(pt, axis=1)
    counts = np.empty((n_genes, n_cells), dtype=np.int32)
    for g0 in range(0, n_genes, chunk):
        g1 = min(g0 + chunk, n_genes)
        n = g1 - g0
        b0 = rng.normal(0.0, 1.0, n)
        dynamic = rng.random(n) < 0.5
        eta = np.repeat(b0[:, None], n_cells, axis=1)
        for _ in range(2):
            tau = rng.uniform(0.1, 0.9, n)
            a = rng.normal(0.0, 2.0, n) * dynamic
            right = rng.random(n) < 0.5
            h = np.where(right[:, None], np.maximum(x[None, :] - tau[:, None], 0.0), np.maximum(tau[:, None] - x[None, :], 0.0))
            eta += a[:, None] * h
        theta = np.exp(rng.uniform(np.log(0.5), np.log(20.0), n))
        theta[rng.random(n) < 0.05] = 1e3
        mu = np.exp(np.clip(eta, -8.0, 8.0)) * depth[None, :]
        lam = rng.gamma(shape=theta[:, None], scale=mu / theta[:, None])
        counts[g0:g1] = rng.poisson(lam).astype(np.int32)
    return counts, pt, depth
