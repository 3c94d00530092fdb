"""CPU-side checks: the C ABI library loads and exports every declared symbol, fails loudly without a GPU, the
host-side K0 (knots / sort / buckets) agrees with the oracle, and the shared host/device algorithm core (run through
the serial test executor in tests/host_emu) reproduces the oracle on the bundled data and on synthetic data."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu, record_terms, rel
from oracle import knots as oknots
from oracle import rmath
from oracle.testdynamic import test_dynamic as oracle_test_dynamic


def test_library_exports_every_declared_symbol():
    import sclane_b200 as sb
    from sclane_b200 import _abi
    header = open(f"{ROOT}/include/sclane_b200.h").read()
    declared = set(re.findall(r"\b(sclane_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_abi.EXPORTED_SYMBOLS)
    lib = C.CDLL(sb.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.sclane_abi_version() == _abi.SCLANE_ABI_VERSION
    lib.sclane_record_size.restype = C.c_size_t
    assert lib.sclane_record_size() == C.sizeof(_abi.SclaneRecord)
    o = _abi.SclaneOpts()
    lib.sclane_default_opts(C.byref(o))
    assert (o.M, o.approx_knot, o.n_knot_max, o.minspan, o.tols_score) == (5, 1, 25, -1, 1e-5)


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
def test_no_cpu_fallback(golden_inputs):
    import sclane_b200 as sb
    with pytest.raises(sb.SclaneError) as e:
        sb.testDynamic(golden_inputs["counts"][:2], golden_inputs["pt"])
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)
    with pytest.raises(sb.SclaneError):
        sb.eigenMapMatMult(np.eye(3), np.eye(3))


def test_bad_arguments_are_reported(golden_inputs):
    import sclane_b200 as sb
    with pytest.raises(ValueError):
        sb.testDynamic(None, None)
    with pytest.raises(ValueError):
        sb.testDynamic(golden_inputs["counts"][:2], golden_inputs["pt"], is_gee=True)     # id.vec missing
    with pytest.raises(sb.SclaneError):
        sb.testDynamic(golden_inputs["counts"][:2], golden_inputs["pt"], n_potential_basis_fns=11)


@pytest.mark.parametrize("n,seed", [(1198, 0), (300, 1), (10000, 2), (50000, 3)])
def test_candidate_knots_host_vs_oracle(n, seed, golden_inputs):
    import sclane_b200 as sb
    rng = np.random.default_rng(seed)
    x = golden_inputs["pt"] if seed == 0 else rng.permutation((np.arange(n) + 0.5) / n)
    if seed == 1:
        x = np.round(rng.random(n) * 37.0, 2)          # unnormalised pseudotime with ties
    k, ms = sb.candidate_knots(x)
    _, ko = oknots.candidate_knots(x)
    assert np.array_equal(k, np.sort(ko))
    assert ms == int(np.rint(-np.log2(-(1.0 / n) * np.log(0.95)) / 2.5))


@pytest.mark.parametrize("case", ["uniform", "nan-membership", "ties"])
def test_lineage_plan_host_tables(case):
    """sclane_lineage_plan (host build; the device build is compared with it bit for bit in tests/test_parity_r02.py): knots == the
    oracle's, the permutation is the stable order by round(x, 4), buckets / coordinates / abscissa tables are consistent."""
    from sclane_b200 import api
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 1, 6000)
    if case == "nan-membership":
        x[rng.uniform(size=x.size) < 0.3] = np.nan
    if case == "ties":
        x = np.round(x, 2)
    P = api.lineage_plan(x)
    cells = np.flatnonzero(~np.isnan(x))
    _, ko = oknots.candidate_knots(x[cells])
    assert P["N"] == cells.size and np.array_equal(P["knots"], np.sort(ko))
    X = np.array([oknots.r_round(v, 4) for v in x[cells]])
    order = cells[np.argsort(X, kind="stable")]
    assert np.array_equal(P["perm"], order)
    Xs = np.array([oknots.r_round(v, 4) for v in x[P["perm"]]])
    T = P["T"]
    bstart = P["bstart"]
    assert bstart[0] == 0 and bstart[T + 1] == P["N"] and np.array_equal(bstart[1:T + 1], np.searchsorted(Xs, P["knots"], side="left"))
    ref = np.concatenate([[P["knots"][0]], P["knots"]])
    bucket = np.searchsorted(bstart[1:T + 1], np.arange(P["N"]), side="right")
    assert np.array_equal(P["u"], Xs - ref[bucket])
    for b in range(T + 1):
        ub = P["u"][bstart[b]:bstart[b + 1]]
        assert P["U"][0, b] == ub.size and abs(P["U"][1, b] - ub.sum()) <= 1e-12 * max(1.0, abs(ub).sum()) and abs(P["U"][2, b] - (ub * ub).sum()) <= 1e-12
    starts = np.flatnonzero(np.concatenate([[True], Xs[1:] != Xs[:-1]]))
    assert P["D"] == starts.size and np.array_equal(P["pstart"], np.concatenate([starts, [P["N"]]]))
    assert np.array_equal(P["pn"], np.diff(P["pstart"])) and np.array_equal(P["pu"], P["u"][starts]) and np.array_equal(P["pb"], bucket[starts])
    assert np.array_equal(P["pbstart"], np.searchsorted(starts, bstart, side="left"))


def test_scalar_helpers_vs_scipy(emu):
    from scipy.special import digamma, gammaln, polygamma
    from scipy.stats import chi2
    for x in (0.003, 0.7, 1.0, 5.5, 9.99, 10.0, 37.2, 5e3, 1e7):
        assert rel(emu.emu_digamma(x), digamma(x)) < 5e-15 * max(1.0, abs(np.log(x)) * 10)
        assert rel(emu.emu_trigamma(x), polygamma(1, x)) < 1e-13
    for y in (0, 1, 2, 5, 8, 9, 30, 500):
        for sigma in (1e-6, 1e-3, 0.04, 0.1, 3.0, 200.0):
            th = 1.0 / sigma
            ref = gammaln(th + y) - gammaln(th) + y * np.log(sigma)
            if th > 1e4 and y > 1:       # the lgamma difference itself loses digits here; compare with the exact sum
                ref = np.sum(np.log1p(np.arange(1, y) * sigma))
            assert abs(emu.emu_nb_lambda(y, sigma) - ref) < 1e-10 * max(1.0, abs(ref)), (y, sigma)
    for q, df in ((0.0, 1), (1e-3, 1), (3.84, 1), (7.8, 1), (50.0, 2), (700.0, 3), (1500.0, 2), (0.5, 4)):
        assert rel(emu.emu_pchisq_upper(q, df), chi2.sf(q, df), 1e-320) < 1e-10
    for v, d in ((0.12345, 4), (2.675, 2), (1e-7, 6), (-3.14159265, 4), (123456.7891, 2), (0.5, 0 + 1)):
        assert emu.emu_r_round(v, d) == rmath.r_round(v, d)


def _compare(recs, oracle, genes, tol=1e-6):
    worst = 0.0
    for gi, g in enumerate(genes):
        r, o = recs[gi], oracle[g]["Lineage_A"]
        ot = o["MARGE_Summary"]["term"] if o["MARGE_Summary"] else None
        if ot is None:
            assert not (r.status & 1), g
            continue
        assert r.status == 3, g
        assert record_terms(r) == ot, g
        if abs(o["Test_Stat"]) > 1e-6:
            worst = max(worst, rel(r.test_stat, o["Test_Stat"]))
        else:
            assert abs(r.test_stat) < 1e-6
        assert r.df == o["Degrees_Freedom"]
        for j in range(r.n_terms):
            worst = max(worst, rel(r.coef[j], o["MARGE_Summary"]["estimate"][j]), rel(r.se[j], o["MARGE_Summary"]["std.error"][j]))
        worst = max(worst, rel(r.loglik, o["LogLik_MARGE"]), rel(r.null_loglik, o["LogLik_Null"]),
                    rel(r.deviance, o["Dev_MARGE"]), rel(r.null_deviance, o["Dev_Null"]), rel(r.null_coef, o["Null_Summary"]["estimate"][0]))
    assert worst < tol, worst


def test_core_algorithm_vs_oracle_bundled(emu, golden_inputs):
    """Bucket-moment formulation + joint-Newton NB fits + glm.nb schedule (the code the CUDA kernels run) against the
    dense numpy oracle on a slice of the bundled data (with size-factor offsets)."""
    sel = list(range(0, 100, 4))
    genes = [golden_inputs["genes"][i] for i in sel]
    counts = golden_inputs["counts"][sel]
    recs, infos = emu.run(counts, golden_inputs["pt"], golden_inputs["size_factor"])
    oracle = oracle_test_dynamic(counts, golden_inputs["pt"], genes, golden_inputs["size_factor"], keep_preds=False)
    assert infos[0].n_knots == 25 and infos[0].n_cells == 1198
    _compare(recs, oracle, genes)


def test_core_algorithm_vs_oracle_synthetic(emu):
    from tools import synth
    counts, pt, _ = synth.make_counts(16, 2500, seed=7)
    counts[3] = 0                      # all-zero gene
    counts[5] = (np.arange(2500) % 97 == 0).astype(np.int32)   # nearly empty gene
    genes = [f"g{i}" for i in range(16)]
    recs, _ = emu.run(counts, pt)
    oracle = oracle_test_dynamic(counts, pt, genes, None, keep_preds=False)
    assert recs[3].status == 0 and (recs[3].marge_notes & 32)
    _compare([recs[i] for i in range(16) if i != 3], {g: oracle[g] for i, g in enumerate(genes) if i != 3},
             [g for i, g in enumerate(genes) if i != 3])


def test_theta_ml_on_chebyshev_nodes_vs_points(emu, capfd):
    """theta.ml at fixed mu on the condensed node list (sclane_comp.cuh: ThetaNodes) against the same Newton steps walked over
    every abscissa: identical schedules (alternations, IRLS iterations) and statistics equal far below the 1e-6 parity bar."""
    from tools import synth
    counts, pt, _ = synth.make_counts(24, 20000, seed=11)
    os.environ["SCLANE_EMU_STAGES"] = "1"
    try:
        os.environ["SCLANE_EMU_THETA_NODES"] = "1"
        a, _ = emu.run(counts, pt)
        err = capfd.readouterr().err
        os.environ["SCLANE_EMU_THETA_NODES"] = "0"
        b, _ = emu.run(counts, pt)
    finally:
        os.environ.pop("SCLANE_EMU_THETA_NODES", None)
        os.environ.pop("SCLANE_EMU_STAGES", None)
    capfd.readouterr()
    on_nodes = sum(int(m) for m in re.findall(r"theta passes on nodes (\d+)", err))
    on_points = sum(int(m) for m in re.findall(r"on points (\d+)", err))
    entries = [float(m) for m in re.findall(r"mean entries ([0-9.]+)", err) if float(m) > 0]
    assert on_nodes > 5 * max(on_points, 1) and np.mean(entries) < 1024, (on_nodes, on_points)
    worst = 0.0
    n_fit = 0
    for ra, rb in zip(a, b):
        assert ra.status == rb.status and ra.n_terms == rb.n_terms
        assert ra.alternations == rb.alternations and ra.irls_iters == rb.irls_iters and ra.fit_passes == rb.fit_passes
        if ra.status != 3:
            continue
        n_fit += 1
        runaway = ra.theta > 1e3                 # theta.ml's likelihood is flat there: theta itself is ill-conditioned
        for x, y in ((ra.loglik, rb.loglik), (ra.deviance, rb.deviance), (ra.test_stat, rb.test_stat)) + tuple(zip(ra.coef[:ra.n_terms], rb.coef[:rb.n_terms])):
            worst = max(worst, rel(x, y, 1e-3))
        if not runaway:
            worst = max(worst, rel(ra.theta, rb.theta))
    assert n_fit >= 12 and worst < 1e-9, (n_fit, worst)


def test_fastmath_tables_vs_libm():
    """sclane_fastmath.cuh (the exp / log / log1p every pass of the fit kernels evaluates) against numpy on 10^6 arguments each:
    a few ulp at most inside the documented range, the documented values outside."""
    import __graft_entry__ as ge
    lib = C.CDLL(ge.EMU)
    lib.emu_fastmath.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_long]

    def call(which, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(x)
        lib.emu_fastmath(which, x.ctypes.data, out.ctypes.data, x.size)
        return out

    rng = np.random.default_rng(5)
    with np.errstate(over="ignore", under="ignore"):
        x = np.concatenate([rng.uniform(-700.0, 700.0, 500000), rng.normal(0.0, 3.0, 500000), [0.0, -0.0, 1e-300, -1e-300, 700.0, -700.0]])
        got, ref = call(0, x), np.exp(x)
        ulp = np.abs(got - ref) / np.spacing(ref)
        assert ulp.max() <= 2.0, ulp.max()
        # the contract outside +-700 (a mean above 1e304 is a diverged fit either way): +inf / 0, NaN propagates
        sp = call(0, np.array([np.inf, -np.inf, 700.5, -700.5, np.nan]))
        assert sp[0] == np.inf and sp[1] == 0.0 and sp[2] == np.inf and sp[3] == 0.0 and np.isnan(sp[4])
        t = np.concatenate([1.0 + rng.exponential(1.0, 400000), np.exp(rng.uniform(0.0, 690.0, 400000)), 1.0 + rng.uniform(0.0, 1e-6, 200000), [1.0]])
        got, ref = call(1, t), np.log(t)
        err = np.abs(got - ref) / np.maximum(np.spacing(ref), np.spacing(1.0) * 0.5 * np.abs(t - 1.0))
        assert err.max() <= 4.0, err.max()
        v = np.concatenate([np.exp(rng.uniform(-40.0, 40.0, 900000)), rng.uniform(0.0, 1e-8, 100000), [0.0]])
        got, ref = call(2, v), np.log1p(v)
        # log1p(x) computed from t = 1 + x as the kernels have it: the rounding of t is part of the contract for tiny x
        tol = np.maximum(4.0 * np.spacing(ref), 1.5 * np.spacing(1.0) * np.minimum(1.0, ref + 1e-300) / np.maximum(v, 1e-300) * ref)
        assert np.all(np.abs(got - ref) <= tol), float(np.max(np.abs(got - ref) / np.maximum(ref, 1e-300)))
