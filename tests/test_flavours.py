"""The oracle flavours (oracle/flavours.py) that restate the third-party behaviour the engine's deterministic rules replace: the
Eigen column-pivoted QR rank check (rule 1) and gamlss's early-stopped RS fit + finite-difference vcov (rule 2).  Full report:
tools/rule_flavours.py -> profiles/r02_rule_flavours.txt."""
import numpy as np

from conftest import GOLDEN_EXCEPTIONS
from oracle.flavours import eigen_colpiv_rank, flavour, gamlss_rs_nbi, gamlss_wald
from oracle.knots import candidate_knots, tp1, tp2
from oracle.nbfit import nb_mle, wald_nb
from oracle.testdynamic import test_dynamic as oracle_test_dynamic


def test_eigen_colpiv_rank_on_hinge_designs(golden_inputs):
    X, knots = candidate_knots(golden_inputs["pt"].reshape(-1))
    one = np.ones_like(X)
    full = np.stack([one, tp1(X, knots[5]), tp2(X, knots[5]), tp1(X, knots[12])], 1)
    assert eigen_colpiv_rank(full) == 4
    # h1(t2) - h2(t2) = h1(t1) - h2(t1) + (t1 - t2): two full pairs + intercept are exactly collinear
    coll = np.stack([one, tp1(X, knots[5]), tp2(X, knots[5]), tp1(X, knots[12]), tp2(X, knots[12])], 1)
    assert eigen_colpiv_rank(coll) == 4
    assert eigen_colpiv_rank(np.stack([one, tp1(X, knots[5]), tp1(X, knots[5])], 1)) == 2
    rng = np.random.default_rng(0)
    assert eigen_colpiv_rank(rng.normal(size=(300, 6))) == 6


def test_gamlss_emulation_lands_next_to_the_mle(golden_inputs):
    """RS-style alternation with gamlss's start values and stopping rules (c.crit = cc = 0.001): the same optimum as the exact MLE,
    left early -- coefficients within 1e-3, Wald statistics within 1e-2 relative on a typical bundled gene."""
    X, knots = candidate_knots(golden_inputs["pt"].reshape(-1))
    B = np.stack([np.ones_like(X), tp1(X, knots[10]), tp2(X, knots[10])], 1)
    for gi in (0, 5, 17):
        y = golden_inputs["counts"][gi].astype(np.float64)
        f = nb_mle(B, y)
        b, s, _, _, _ = gamlss_rs_nbi(B, y)
        assert np.max(np.abs(b - f.beta)) < 1e-3 and abs(s - f.sigma) < 1e-3 * max(f.sigma, 1e-3)
        w_mle, _ = wald_nb(B, y)
        w_rs, fallback = gamlss_wald(B, y)
        if not f.poisson and not fallback:
            assert np.allclose(w_rs, w_mle, rtol=1e-2, atol=1e-4)


def test_exception_genes_are_not_recovered_by_the_emulated_rank_check(golden_inputs, golden_models):
    """Rule 1 evidence: an emulation of Eigen's ColPivHouseholderQR rank decision (threshold eps * diagSize * max pivot) fires on the
    exactly collinear 5-column designs (test above), loses no other gene and recovers none of the exception genes -- the
    reference's 11 exception genes need the check NOT to fire, which depends on the build's floating-point details (SIMD norm
    updates, src/Makevars:10-11) and cannot be restated deterministically.  Counts over all 100 genes: profiles/r02_rule_flavours.txt."""
    genes = golden_inputs["genes"]
    sel = [i for i, g in enumerate(genes) if g in GOLDEN_EXCEPTIONS][:6] + [i for i, g in enumerate(genes) if g not in GOLDEN_EXCEPTIONS][:6]
    counts, names = golden_inputs["counts"][sel], [genes[i] for i in sel]
    base = oracle_test_dynamic(counts, golden_inputs["pt"], names, golden_inputs["size_factor"], keep_preds=False)
    with flavour(rank_rule="colpiv_qr"):
        piv = oracle_test_dynamic(counts, golden_inputs["pt"], names, golden_inputs["size_factor"], keep_preds=False)
    for g in names:
        gold = golden_models[g]["MARGE_Summary"]["term"]
        if g in GOLDEN_EXCEPTIONS:            # recovered by neither rule
            assert piv[g]["Lineage_A"]["MARGE_Summary"]["term"] != gold and base[g]["Lineage_A"]["MARGE_Summary"]["term"] != gold, g
        else:                                 # and no other gene is lost
            assert piv[g]["Lineage_A"]["MARGE_Summary"]["term"] == base[g]["Lineage_A"]["MARGE_Summary"]["term"] == gold, g
