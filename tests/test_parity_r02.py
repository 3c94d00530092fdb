"""Round-2 parity tests of the CUDA path (through the C ABI) against the independent dense oracle at the BASELINE shapes and on
the edge cases the selection rules of marge2.R care about.  Everything here needs a B200 (pytest -m gpu)."""
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np
import pytest

from conftest import GOLDEN_EXCEPTIONS, rel
from oracle import marge as omarge
from oracle.testdynamic import test_dynamic as oracle_test_dynamic
from test_gpu_parity import TOL, _compare_with_oracle

pytestmark = pytest.mark.gpu


def _oracle_chunk(a):
    from threadpoolctl import threadpool_limits
    counts, pt, genes, sf = a
    with threadpool_limits(limits=1):
        return oracle_test_dynamic(counts, pt, genes, sf, keep_preds=False)


def oracle_parallel(counts, pt, genes, sf=None):
    """The oracle over a process pool (one chunk of genes per core)."""
    cores = min(os.cpu_count() or 1, 16, len(genes))
    chunks = [(counts[i::cores], pt, genes[i::cores], sf) for i in range(cores)]
    out = {}
    with ProcessPoolExecutor(max_workers=cores) as ex:
        for part in ex.map(_oracle_chunk, chunks):
            out.update(part)
    return {g: out[g] for g in genes}


def _oracle_chunk_opts(a):
    from threadpoolctl import threadpool_limits
    counts, pt, genes, sf, kw = a
    with threadpool_limits(limits=1):
        return oracle_test_dynamic(counts, pt, genes, sf, keep_preds=False, **kw)


def oracle_parallel_opts(counts, pt, genes, sf=None, **kw):
    cores = min(os.cpu_count() or 1, 16, len(genes))
    chunks = [(counts[i::cores], pt, genes[i::cores], sf, kw) for i in range(cores)]
    out = {}
    with ProcessPoolExecutor(max_workers=cores) as ex:
        for part in ex.map(_oracle_chunk_opts, chunks):
            out.update(part)
    return {g: out[g] for g in genes}


def test_c2_shape_208_genes_vs_oracle():
    """BASELINE configs[1] shape (10,000 cells, M = 5, no offset), 208 genes of the bench generator: default (X-compressed) route
    against the dense oracle -- term sets, LRT, coefficients, SEs, log-likelihoods, deviances, theta."""
    import sclane_b200 as sb
    from tools import synth
    counts, pt, _ = synth.make_counts(208, 10000, seed=312)
    genes = [f"g{i}" for i in range(counts.shape[0])]
    res = sb.testDynamic(counts, pt, genes=genes, n_gpus=1, preds="none")
    assert res.profile["ms"]["comp_forward"] > 0
    worst = _compare_with_oracle(res, oracle_parallel(counts, pt.reshape(-1), genes))
    print("C2 shape, worst relative differences:", worst)


def test_c4_shape_two_lineages_offsets_vs_oracle():
    """BASELINE configs[3] shape: 50,000 cells, 2 lineages (60 % of the cells each, NaN outside), createCellOffset() offsets;
    64 genes = 128 gene-lineage fits.  Exercises the offset quadrature (intercept-only fits) and the hybrid per-cell passes."""
    import bench
    import sclane_b200 as sb
    from tools import synth
    counts, pt, _ = synth.make_counts(64, 50000, seed=5)
    pt2 = bench.two_lineages(pt.reshape(-1))
    sf = sb.createCellOffset(counts)
    genes = [f"g{i}" for i in range(counts.shape[0])]
    res = sb.testDynamic(counts, pt2, genes=genes, size_factor_offset=sf, n_gpus=1, preds="none")
    assert res.profile["ms"]["off_null"] > 0 and res.profile["launches"]["off_aggregate"] == 2
    worst = _compare_with_oracle(res, oracle_parallel(counts, pt2, genes, sf))
    print("C4 shape, worst relative differences:", worst)


def _records_close(a, b, tol, what):
    assert len(a) == len(b)
    worst = 0.0
    for i, (ra, rb) in enumerate(zip(a, b)):
        assert ra.status == rb.status and ra.n_terms == rb.n_terms, (what, i)
        assert [ra.term_kind[j] for j in range(ra.n_terms)] == [rb.term_kind[j] for j in range(rb.n_terms)], (what, i)
        assert [ra.term_knot_idx[j] for j in range(ra.n_terms)] == [rb.term_knot_idx[j] for j in range(rb.n_terms)], (what, i)
        # near-Poisson fits (theta running away, th.warn): glm.nb's alternation count and theta itself depend on rounding-level
        # differences ("theta is schedule-dependent; l, beta, LRT are not") -- compared without theta there
        runaway = bool((ra.marge_notes | ra.null_notes | rb.marge_notes | rb.null_notes) & 7) or max(ra.null_theta, rb.null_theta) > 1e3
        fields = ["loglik", "null_loglik", "deviance", "null_deviance", "null_coef", "null_se"]
        if not runaway:
            assert (ra.alternations, ra.null_alternations, ra.irls_iters, ra.null_irls_iters) == (rb.alternations, rb.null_alternations, rb.irls_iters, rb.null_irls_iters), (what, i)
            fields += ["theta", "null_theta", "se_theta"]
        for f in fields:
            va, vb = getattr(ra, f), getattr(rb, f)
            if va == va or vb == vb:
                worst = max(worst, rel(va, vb) * (1e-4 if runaway else 1.0))          # runaway fits: 1e-6 instead of 1e-10
        if ra.test_stat == ra.test_stat:
            worst = max(worst, abs(ra.test_stat - rb.test_stat) / (abs(rb.test_stat) + 1e-3))
        for j in range(ra.n_terms):
            worst = max(worst, rel(ra.coef[j], rb.coef[j]), rel(ra.se[j], rb.se[j]))
    assert worst < tol, (what, worst)
    return worst


def test_offset_quadrature_equals_per_cell_fits():
    """The intercept-only glm.nb with an offset through the offset quadrature (K bins x 16 Chebyshev nodes) against the same
    fits per cell: identical iteration counts, every statistic within 1e-8 (measured 4e-10: different summation orders) -- incl.
    offsets spanning e^7."""
    import sclane_b200 as sb
    from sclane_b200 import api
    from tools import synth
    counts, pt, _ = synth.make_counts(96, 20000, seed=17)         # D = 10,001 abscissae <= 0.8 N: the X-compressed route
    rng = np.random.default_rng(3)
    for spread in (0.0, 1.0):
        sf = 1e4 / np.maximum(counts.sum(axis=0), 1) * np.exp(rng.normal(0.0, spread, counts.shape[1]))
        a, _ = api.run_records(counts, pt, sf, api._opts(gpu_ids=[0]))
        assert sb.get_profile()["ms"]["off_null"] > 0
        b, _ = api.run_records(counts, pt, sf, api._opts(gpu_ids=[0], offset_per_cell=True))
        assert sb.get_profile()["ms"]["off_null"] == 0
        w = _records_close(a, b, 1e-8, f"offset spread {spread}")
        print(f"offset quadrature vs per cell, log-offset range {np.ptp(np.log(sf)):.2f}: worst relative difference {w:.2e}")


def test_warp_per_gene_kernels_equal_cta_per_gene():
    """The opt-in warp-per-gene persistent kernels (work queue) give the records of the CTA-per-gene kernels."""
    from sclane_b200 import api
    from tools import synth
    counts, pt, _ = synth.make_counts(300, 20000, seed=23)
    a, _ = api.run_records(counts, pt, None, api._opts(gpu_ids=[0]))
    b, _ = api.run_records(counts, pt, None, api._opts(gpu_ids=[0], comp_warp_per_gene=True))
    _records_close(a, b, 1e-9, "warp per gene")


def _symmetric_case(n_genes, seed):
    """121 cells on a symmetric 4-dp grid with mirrored counts: the candidate knots come out symmetric (positions
    1, 6, ..., 121 of min_span; < 25 of them, so no seq()), and in the first forward iteration the pair candidate at knot t and at
    1 - t span mirror images of each other -> their scores tie after round(., 6) (marge2.R:537-549: which.max = FIRST maximum)."""
    rng = np.random.default_rng(seed)
    x = 0.02 + 0.008 * np.arange(121)
    half = np.empty((n_genes, 61), dtype=np.int64)
    for g in range(n_genes):
        mu = np.exp(1.0 + rng.normal(0, 0.5) + rng.normal(0, 3.0) * np.maximum(x[:61] - rng.uniform(0.1, 0.45), 0.0))
        half[g] = rng.negative_binomial(5.0, 5.0 / (5.0 + mu))
    counts = np.concatenate([half, half[:, :60][:, ::-1]], axis=1).astype(np.int32)
    assert np.array_equal(counts, counts[:, ::-1])
    return counts, x


def test_exact_score_ties_first_maximum_rule():
    """Engineered exact ties (north_star: "ties broken identically"): symmetric data make the two best pair candidates tie after
    round(score, 6); marge2.R's which.max() takes the first.  CUDA == oracle on every gene, and the tie is real."""
    import sclane_b200 as sb
    from sclane_b200 import api
    counts, x = _symmetric_case(40, 7)
    genes = [f"s{i}" for i in range(counts.shape[0])]
    res = sb.testDynamic(counts, x, genes=genes, n_gpus=1, preds="none")
    knots = np.array(res.lineage_info[0]["knots"])
    assert len(knots) < 25 and np.allclose(knots + knots[::-1], 1.0, atol=1e-12), knots
    _compare_with_oracle(res, oracle_test_dynamic(counts, x, genes, None, keep_preds=False))
    # the sweep's own scores: intercept-only basis, sigma / mu of the null fit
    beta, sigma, mu = api.stat_out_score_glm_null(counts, x, knots, [0], [0])
    sc = api.score_fun_glm_sweep(counts, mu, sigma, x, knots, [0], [0])
    n_tied = 0
    for g in range(counts.shape[0]):
        both = np.round(sc[g, 1], 6)
        k = int(np.nanargmax(both))
        mirror = len(knots) - 1 - k
        if mirror != k and both[mirror] == both[k]:
            n_tied += 1
            rec = res.records[g]
            if rec.fwd_n_terms >= 3 and rec.fwd_kind[2] == 2:            # iteration 1 took the pair: at the FIRST of the tied knots
                assert rec.fwd_knot_idx[1] == min(k, mirror), (g, k, mirror, rec.fwd_knot_idx[1])
    assert n_tied >= 10, n_tied


def test_heavy_pseudotime_ties_and_few_knots():
    """3,000 cells on only 40 distinct pseudotime values (D << 25 knots' worth after max_span), and a 120-cell lineage with
    fewer than 25 candidate knots: CUDA == oracle."""
    import sclane_b200 as sb
    from tools import synth
    rng = np.random.default_rng(11)
    counts, pt, _ = synth.make_counts(24, 3000, seed=9)
    pt40 = np.round(np.floor(pt.reshape(-1) * 40) / 40 + 0.0125, 4)
    genes = [f"g{i}" for i in range(24)]
    res = sb.testDynamic(counts, pt40, genes=genes, n_gpus=1, preds="none")
    assert len(np.unique(pt40)) == 40 and len(res.lineage_info[0]["knots"]) <= 25
    _compare_with_oracle(res, oracle_test_dynamic(counts, pt40, genes, None, keep_preds=False))
    c2, p2, _ = synth.make_counts(24, 120, seed=10)
    res2 = sb.testDynamic(c2, p2, genes=genes, n_gpus=1, preds="none")
    assert len(res2.lineage_info[0]["knots"]) < 25
    _compare_with_oracle(res2, oracle_test_dynamic(c2, p2.reshape(-1), genes, None, keep_preds=False))
    del rng


def test_pseudotime_outside_unit_interval_takes_per_cell_route():
    """Pseudotime on [0, 50]: 4-dp rounding leaves D ~ N abscissae (> 0.8 N), so the library routes every fit to the per-cell
    kernels on its own; with and without offsets, CUDA == oracle."""
    import sclane_b200 as sb
    from tools import synth
    counts, pt, _ = synth.make_counts(24, 3000, seed=13)
    pt50 = pt.reshape(-1) * 50.0
    genes = [f"g{i}" for i in range(24)]
    res = sb.testDynamic(counts, pt50, genes=genes, n_gpus=1, preds="none")
    assert res.profile["ms"]["comp_forward"] == 0 and res.profile["ms"]["fwd_fit"] > 0
    _compare_with_oracle(res, oracle_test_dynamic(counts, pt50, genes, None, keep_preds=False))
    sf = sb.createCellOffset(counts)
    res = sb.testDynamic(counts, pt50, genes=genes, size_factor_offset=sf, n_gpus=1, preds="none")
    assert res.profile["ms"]["off_null"] == 0
    _compare_with_oracle(res, oracle_test_dynamic(counts, pt50, genes, sf, keep_preds=False))


@pytest.mark.parametrize("kw", [{"minspan": 30}, {"n_knot_max": 10}, {"n_knot_max": 32}, {"M": 7, "n_knot_max": 12}])
def test_marge2_wrapper_options_vs_oracle(kw):
    """marge2() (marge2.R:56-72) with minspan given, n.knot.max in {10, 12, 32} and M = 7: coefficient names and values == oracle."""
    import sclane_b200 as sb
    from tools import synth
    counts, pt, _ = synth.make_counts(10, 2500, seed=31)
    x = pt.reshape(-1)
    n_ok = 0
    for g in range(counts.shape[0]):
        o = omarge.marge2(x, counts[g].astype(np.float64), None, **kw)
        try:
            r = sb.marge2(x, counts[g], **kw)
        except sb.SclaneError:
            assert not o.ok
            continue
        assert o.ok
        assert r["coef_names"] == o.coef_names, (g, kw)
        for a, b in zip(r["final_mod"]["coefficients"].values(), o.fit.coef):
            assert rel(a, b) < TOL
        assert rel(r["final_mod"]["logLik"], o.fit.loglik) < TOL
        lim = kw.get("n_knot_max", 25)
        assert len(r["knots"]) <= lim
        n_ok += 1
    assert n_ok >= 6


def test_product_mirror_vs_reference_golden(golden_inputs, golden_models):
    """The host mirror of the reference's R functions (sclane_b200/api.py) against the reference's own bundled output:
    Gene_Dynamics (summarizeModel), MARGE_Slope_Data / testSlope, marge2() and modelLRT()."""
    import sclane_b200 as sb
    res = sb.testDynamic(golden_inputs["counts"], golden_inputs["pt"], genes=golden_inputs["genes"],
                         size_factor_offset=golden_inputs["size_factor"], n_gpus=1, preds="none")
    checked = 0
    for g, gold in golden_models.items():
        if g in GOLDEN_EXCEPTIONS:
            continue
        r = res[g]["Lineage_A"]
        gd, gg = r["Gene_Dynamics"], gold["Gene_Dynamics"]
        assert set(gd) == set(gg), (g, sorted(gd), sorted(gg))
        for k in gg:
            a, b = gd[k][0], gg[k][0]
            if isinstance(b, str) or b is None or isinstance(a, str):
                assert a == b, (g, k, a, b)
            else:
                assert a is not None and (abs(a - b) <= 1e-6 * (abs(b) + 1e-9)), (g, k, a, b)
        sd, sg = r["MARGE_Slope_Data"], gold["MARGE_Slope_Data"]
        assert sd["Notes"] == sg["Notes"] and sd["Direction"] == sg["Direction"]
        for k in ("Breakpoint", "Rounded_Breakpoint", "Beta", "Test_Stat", "P_Val"):
            for a, b in zip(sd[k], sg[k]):
                assert (a is None and b is None) or rel(a, b, 1e-300) < 1e-5, (g, k, a, b)
        checked += 1
    assert checked == 89
    # testSlope(): BH over all breakpoints, flags per gene
    ts = sb.testSlope(res)
    assert {"Gene", "Breakpoint", "P_Val_Adj", "Gene_Dynamic_Lineage_Slope", "Gene_Dynamic_Lineage", "Gene_Dynamic_Overall"} <= set(ts.columns)
    pv = ts["P_Val"].to_numpy(dtype=float)
    ok = ~np.isnan(pv)
    from oracle.rmath import p_adjust_bh
    assert np.allclose(ts["P_Val_Adj"].to_numpy(dtype=float)[ok], p_adjust_bh(pv[ok]), rtol=1e-12)
    assert (ts.groupby("Gene")["Gene_Dynamic_Overall"].nunique() == 1).all()
    # marge2() + modelLRT() on single genes == the batch result and the golden LRT
    x, sf = golden_inputs["pt"].reshape(-1), golden_inputs["size_factor"]
    for g in [gg for gg in golden_inputs["genes"] if gg not in GOLDEN_EXCEPTIONS][:12]:
        i = golden_inputs["genes"].index(g)
        gold = golden_models[g]
        m1 = sb.marge2(x, golden_inputs["counts"][i], sf)
        assert m1["coef_names"] == gold["MARGE_Summary"]["term"]
        lrt = sb.modelLRT(m1, {"logLik": gold["LogLik_Null"], "df": 2})
        assert rel(lrt["LRT_Stat"], gold["Test_Stat"], 1e-6) < TOL and lrt["DF"] == gold["Degrees_Freedom"]
        assert rel(lrt["P_Val"], gold["P_Val"], 1e-300) < 1e-5
    assert sb.modelLRT(None, {"logLik": 0.0, "df": 2})["Notes"] == "No test performed due to model failure."


def test_in_library_multi_gpu_sharding_is_bit_identical():
    """opts.n_gpus > 1: one host thread + stream per GPU inside the library; records equal those of one GPU byte for byte."""
    import sclane_b200 as sb
    from sclane_b200 import _abi, api
    from tools import synth
    n = sb.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    counts, pt, _ = synth.make_counts(301, 20000, seed=5)
    a, _ = api.run_records(counts, pt, None, api._opts(gpu_ids=[0]))
    b, _ = api.run_records(counts, pt, None, api._opts(gpu_ids=list(range(n))))
    off, sz = _abi.SclaneRecord.gene_time.offset, 8
    cyc, csz = _abi.SclaneRecord.stage_cycles.offset, 24
    for ra, rb in zip(a, b):
        ba, bb = bytearray(bytes(ra)), bytearray(bytes(rb))
        for o_, s_ in ((off, sz), (cyc, csz)):
            ba[o_:o_ + s_] = bb[o_:o_ + s_] = bytes(s_)
        assert ba == bb


def _split_edge_knot_genes(res, ref, n_knots):
    """approx.knot = FALSE keeps candidates right up to 8 abscissae from either end of the lineage.  A gene whose FIRST forward
    iteration picks one of the two outermost candidates gets a hinge supported on a handful of cells: the NB fit of that column is
    ill-conditioned (no MLE when those cells are all zero), the second iteration's scores then differ at 1e-3 between any two
    implementations and the arg-max among near-equal candidates flips; an outermost knot picked in a LATER iteration leaves the
    forward pass identical but makes the backward pass's Wald statistics of that column just as arbitrary.  Those genes (decided
    from the ORACLE's trace alone) are compared on the first iteration only."""
    edge = set()
    for g in res:
        ft = ref[g]["Lineage_A"]["_trace"].forward_terms
        if any(t.knot_idx <= 1 or t.knot_idx >= n_knots - 2 for t in ft[1:]):
            edge.add(g)
    return edge


@pytest.mark.parametrize("n_cells,n_genes", [(3000, 56), (10000, 56)])
def test_approx_knot_false_vs_oracle(n_cells, n_genes):
    """approx.knot = FALSE (marge2.R:167-173): every abscissa kept by min_span() / max_span() is a candidate knot (hundreds to
    thousands instead of 25).  The prefix-moment sweep (sclane_exact_device.cuh) against the dense oracle, which rebuilds the two
    hinge columns per candidate like the reference does: term sets, knots, LRT, coefficients."""
    import sclane_b200 as sb
    from tools import synth
    counts, pt, _ = synth.make_counts(n_genes, n_cells, seed=41)
    genes = [f"g{i}" for i in range(n_genes)]
    res = sb.testDynamic(counts, pt, genes=genes, approx_knot=False, n_gpus=1, preds="none")
    n_knots = res.lineage_info[0]["n_knots"]
    assert n_knots > 100
    ref = oracle_parallel_opts(counts, pt.reshape(-1), genes, None, approx_knot=False)
    edge = _split_edge_knot_genes(res, ref, n_knots)
    assert len(edge) <= n_genes // 4, edge
    keep = {g: v for g, v in res.items() if g not in edge}
    worst = _compare_with_oracle(keep, ref)
    for i, g in enumerate(genes):                         # every gene, edge or not: the first iteration's score is the oracle's
        sc = ref[g]["Lineage_A"]["_trace"].forward_scores
        if sc:
            assert rel(res.records[i].fwd_score[0], float(sc[0])) < TOL, g
    print(f"approx.knot = FALSE, {n_cells} cells, {n_knots} candidate knots, {len(edge)} edge-knot genes set aside: worst relative differences", worst)


def test_approx_knot_false_with_offsets_vs_oracle():
    """approx.knot = FALSE with createCellOffset() offsets: per-gene buckets in the hybrid per-cell final fits + offset quadrature."""
    import sclane_b200 as sb
    from tools import synth
    counts, pt, _ = synth.make_counts(32, 4000, seed=43)
    sf = sb.createCellOffset(counts)
    genes = [f"g{i}" for i in range(32)]
    res = sb.testDynamic(counts, pt, genes=genes, size_factor_offset=sf, approx_knot=False, n_gpus=1, preds="none")
    ref = oracle_parallel_opts(counts, pt.reshape(-1), genes, sf, approx_knot=False)
    edge = _split_edge_knot_genes(res, ref, res.lineage_info[0]["n_knots"])
    assert len(edge) <= 16, edge                           # decided by the ORACLE's trace alone (offsets push more knots outwards)
    _compare_with_oracle({g: v for g, v in res.items() if g not in edge}, ref)
    for i, g in enumerate(genes):                          # every gene, edge or not: first forward score and the null model (offset quadrature)
        sc = ref[g]["Lineage_A"]["_trace"].forward_scores
        if sc:
            assert rel(res.records[i].fwd_score[0], float(sc[0])) < TOL, g
        o = ref[g]["Lineage_A"]
        if o["LogLik_Null"] == o["LogLik_Null"]:
            assert rel(res.records[i].null_loglik, o["LogLik_Null"]) < TOL and rel(res.records[i].null_deviance, o["Dev_Null"]) < TOL, g


def _plan_cases():
    rng = np.random.default_rng(20)
    cases = []
    cases.append(("uniform 200k", rng.uniform(0, 1, 200000), {}))
    cases.append(("uniform 10k", rng.uniform(0, 1, 10000), {}))
    x = rng.uniform(0, 1, 30000)
    x[rng.uniform(size=x.size) < 0.4] = np.nan
    cases.append(("40 % of the cells outside the lineage", x, {}))
    cases.append(("two-decimal pseudotime (101 abscissae)", np.round(rng.uniform(0, 1, 20000), 2), {}))
    cases.append(("41 abscissae, fewer candidates than n.knot.max", rng.integers(0, 41, 5000) / 40.0, {}))
    cases.append(("minspan given", rng.beta(2, 5, 15000), {"minspan": 5}))
    cases.append(("n.knot.max 10", rng.beta(0.5, 0.5, 15000), {"n_knot_max": 10}))
    cases.append(("n.knot.max 32", rng.uniform(0, 1, 4000), {"n_knot_max": 32}))
    cases.append(("pseudotime in [0, 37.5]: three radix passes", rng.uniform(0, 37.5, 120000), {}))
    cases.append(("negative pseudotime", rng.normal(0, 0.3, 25000), {}))
    cases.append(("35 cells", rng.uniform(0, 1, 35), {}))
    return cases


@pytest.mark.parametrize("case", _plan_cases(), ids=lambda c: c[0])
def test_device_plan_bit_identical(case):
    """K0 on the device (sclane_k0_device.cuh) against the host plan (sclane_host.hpp, itself == the oracle's knots in
    tests/test_host_logic.py): every table -- sort permutation, knots, bucket table, coordinates, unweighted moments, abscissa
    tables -- bit for bit (marge2.R:150-166, min_span.R:24-34, max_span.R:20-25)."""
    from sclane_b200 import api
    _, x, kw = case
    o = api._opts(**kw)
    host = api.lineage_plan(x, o, on_device=False)
    dev = api.lineage_plan(x, o, on_device=True)
    for k in ("N", "T", "D", "minspan"):
        assert host[k] == dev[k], k
    assert host["T"] >= 1
    for k in ("perm", "u", "knots", "bstart", "U", "pstart", "pn", "pu", "pb", "pbstart"):
        a, b = np.ascontiguousarray(host[k]), np.ascontiguousarray(dev[k])
        assert a.shape == b.shape and a.tobytes() == b.tobytes(), k


def test_device_plan_records_equal_host_plan_records():
    """The batch path with K0 on the device (default) and with opts.host_plan = 1: records byte for byte; inputs the device
    plan does not take (size factors, approx.knot = FALSE) keep the host plan."""
    import sclane_b200 as sb
    from sclane_b200 import api
    from test_gpu_parity import _rec_bytes
    from tools import synth
    counts, pt, _ = synth.make_counts(160, 20000, seed=9)
    pt2 = np.stack([pt.reshape(-1), np.where(np.arange(20000) % 3 == 0, np.nan, pt.reshape(-1) ** 2)], axis=1)
    a, ia = api.run_records(counts, pt2, None, api._opts(gpu_ids=[0]))
    assert sb.get_profile()["launches"]["plan"] > 0 and sb.get_profile()["ms"]["plan"] > 0
    b, ib = api.run_records(counts, pt2, None, api._opts(gpu_ids=[0], host_plan=True))
    assert sb.get_profile()["ms"]["plan"] == 0              # (launches["plan"] still counts k_bucket_nodes, one per lineage)
    for ra, rb in zip(a, b):
        assert _rec_bytes(ra) == _rec_bytes(rb)
    for la, lb in zip(ia, ib):
        assert bytes(la) == bytes(lb)
    sf = np.exp(np.random.default_rng(1).normal(0, 0.3, 20000))
    api.run_records(counts[:16], pt2, sf, api._opts(gpu_ids=[0]))
    assert sb.get_profile()["ms"]["plan"] == 0
    with pytest.raises(sb.SclaneError):
        api.lineage_plan(pt2[:, 0], api._opts(approx_knot=False), on_device=True)
    with pytest.raises(sb.SclaneError):                         # a lineage without candidate knots is an argument error on both builds
        api.run_records(counts[:4, :200], np.repeat([0.25, 0.5], 100), None, api._opts(gpu_ids=[0]))
    with pytest.raises(sb.SclaneError):
        api.run_records(counts[:4, :200], np.repeat([0.25, 0.5], 100), None, api._opts(gpu_ids=[0], host_plan=True))


def test_batch_streams_do_not_change_records():
    """Gene batches in flight on 1, 2 and 3 streams (sclane_opts.batch_streams), device-resident and host-fed, with and without
    offsets: records byte for byte (which stream fits which gene affects nothing)."""
    import torch
    import sclane_b200 as sb
    from sclane_b200 import api
    from test_gpu_parity import _rec_bytes
    from tools import synth
    counts, pt, _ = synth.make_counts(700, 12000, seed=21)
    sf = np.exp(np.random.default_rng(2).normal(0, 0.4, 12000))
    d_counts = torch.from_numpy(counts).cuda()
    for size_factor in (None, sf):
        ref, _ = api.run_records(counts, pt, size_factor, api._opts(gpu_ids=[0], batch_streams=1))
        for kw in ({"batch_streams": 2, "genes_per_batch": 100}, {"batch_streams": 3, "genes_per_batch": 64}, {"batch_streams": 0}):
            got, _ = api.run_records(counts, pt, size_factor, api._opts(gpu_ids=[0], **kw))
            assert all(_rec_bytes(a) == _rec_bytes(b) for a, b in zip(ref, got)), kw
            got, _ = api.run_records(counts.shape[0], pt, size_factor, api._opts(gpu_ids=[0], **kw), device_ptr=d_counts.data_ptr(), device=0)
            assert all(_rec_bytes(a) == _rec_bytes(b) for a, b in zip(ref, got)), kw


def test_hybrid_theta_nodes_vs_cell_walk():
    """Fits with a size-factor offset (hybrid final stage): theta.ml at fixed mu on the Chebyshev nodes of the (offset bin,
    bucket) groups against the same Newton steps walked over every cell -- identical schedules (alternations, IRLS iterations,
    passes) and statistics far below the 1e-6 parity bar; two lineages with NaN membership, wide and narrow offset ranges."""
    from sclane_b200 import api
    from tools import synth
    counts, pt, _ = synth.make_counts(240, 30000, seed=17)
    x = pt.reshape(-1)
    rng = np.random.default_rng(4)
    pt2 = np.stack([np.where(rng.uniform(size=x.size) < 0.7, x, np.nan), np.where(rng.uniform(size=x.size) < 0.5, 1.0 - x, np.nan)], axis=1)
    for spread in (0.25, 1.2):
        sf = np.exp(rng.normal(0, spread, x.size))
        try:
            os.environ["SCLANE_THETA_NODES"] = "1"
            a, _ = api.run_records(counts, pt2, sf, api._opts(gpu_ids=[0]))
            os.environ["SCLANE_THETA_NODES"] = "0"
            b, _ = api.run_records(counts, pt2, sf, api._opts(gpu_ids=[0]))
        finally:
            os.environ.pop("SCLANE_THETA_NODES", None)
        worst, n_fit, n_hinge = 0.0, 0, 0
        for ra, rb in zip(a, b):
            assert ra.status == rb.status and ra.n_terms == rb.n_terms
            if ra.status != 3:
                continue
            runaway = ra.theta > 1e3 or rb.theta > 1e3           # theta.ml's likelihood is flat there
            if not runaway:
                assert ra.alternations == rb.alternations and ra.irls_iters == rb.irls_iters and ra.fit_passes == rb.fit_passes
            n_fit += 1
            n_hinge += ra.n_terms > 1
            for u_, v_ in ((ra.loglik, rb.loglik), (ra.deviance, rb.deviance), (ra.test_stat, rb.test_stat)) + tuple(zip(ra.coef[:ra.n_terms], rb.coef[:rb.n_terms])):
                worst = max(worst, rel(u_, v_, 1e-3))
            if not runaway:
                worst = max(worst, rel(ra.theta, rb.theta))
        assert n_fit > 300 and n_hinge > 150 and worst < 1e-9, (n_fit, n_hinge, worst)


@pytest.mark.parametrize("n_genes,n_cells", [(400, 10000), (96, 200000)])
def test_bucket_nodes_vs_point_walk(n_genes, n_cells):
    """X-compressed route: every pass over a sloped knot bucket on the bucket's 16 Chebyshev nodes (sclane_comp.cuh: BucketNodes)
    against the same pass walked over the bucket's abscissae -- the same discrete choices (forward / backward term sets, glm.nb
    schedules) and statistics far below the 1e-6 parity bar."""
    from sclane_b200 import api
    from tools import synth
    counts, pt, _ = synth.make_counts(n_genes, n_cells, seed=33)
    try:
        os.environ["SCLANE_BUCKET_NODES"] = "1"
        a, _ = api.run_records(counts, pt, None, api._opts(gpu_ids=[0]))
        os.environ["SCLANE_BUCKET_NODES"] = "0"
        b, _ = api.run_records(counts, pt, None, api._opts(gpu_ids=[0]))
    finally:
        os.environ.pop("SCLANE_BUCKET_NODES", None)
    worst, n_fit, n_sched = 0.0, 0, 0
    for ra, rb in zip(a, b):
        assert ra.status == rb.status and ra.n_terms == rb.n_terms
        assert list(ra.term_kind[:ra.n_terms]) == list(rb.term_kind[:rb.n_terms]) and list(ra.term_knot_idx[:ra.n_terms]) == list(rb.term_knot_idx[:rb.n_terms])
        assert list(ra.fwd_kind[:ra.fwd_n_terms]) == list(rb.fwd_kind[:rb.fwd_n_terms]) and list(ra.fwd_knot_idx[:ra.fwd_n_terms]) == list(rb.fwd_knot_idx[:rb.fwd_n_terms])
        if ra.status != 3:
            continue
        n_fit += 1
        runaway = ra.theta > 1e3 or rb.theta > 1e3
        n_sched += (ra.alternations == rb.alternations and ra.irls_iters == rb.irls_iters)
        for u_, v_ in ((ra.loglik, rb.loglik), (ra.deviance, rb.deviance), (ra.test_stat, rb.test_stat), (ra.null_loglik, rb.null_loglik)) + tuple(zip(ra.coef[:ra.n_terms], rb.coef[:rb.n_terms])) + tuple(zip(ra.se[:ra.n_terms], rb.se[:rb.n_terms])):
            worst = max(worst, rel(u_, v_, 1e-3))
        if not runaway:
            worst = max(worst, rel(ra.theta, rb.theta))
    assert n_fit > 0.8 * n_genes and n_sched >= n_fit - 2 and worst < 1e-9, (n_fit, n_sched, worst)
