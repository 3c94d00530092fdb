"""GEE mode (testDynamic(is.gee = TRUE, gee.test = "wald")): the O(n) closed-form core against the dense oracle
(oracle/gee.py, which builds and inverts the n_i x n_i working covariances like the reference does), on the CPU through
the host emulation harness and on the GPU through the C ABI.

PARITY UNPINNED: the reference ships no GEE outputs and geeM::geem is third-party code outside the reference tree, so
these tests pin the CUDA path to the restated algorithm, not to numbers produced by R."""
import numpy as np
import pytest

from conftest import record_terms, rel

from oracle.gee import test_dynamic_gee as oracle_gee

TOL = 1e-6          # relative, fp64 statistics (north_star tolerance); discrete choices (term sets) must be identical


def _gee_data(seed=5, sizes=(150, 1, 120, 89), n_genes=6, sorted_within=False):
    from tools import synth
    n = int(sum(sizes))
    counts, pt, _ = synth.make_counts(n_genes=n_genes, n_cells=n, seed=seed)
    pt = np.asarray(pt, dtype=np.float64).reshape(-1)
    rng = np.random.default_rng(seed)
    subject = np.repeat(np.arange(len(sizes)), sizes)
    if sorted_within:
        start = 0
        order = []
        for s in sizes:
            idx = np.arange(start, start + s)
            order.append(idx[np.argsort(pt[idx], kind="stable")])
            start += s
        order = np.concatenate(order)
        counts, pt = counts[:, order], pt[order]
    sf = rng.uniform(0.6, 1.6, n)
    return np.ascontiguousarray(counts, dtype=np.int32), pt, subject, sf


def _compare(recs, ref, genes, tol=TOL):
    bad = []
    for g, name in enumerate(genes):
        r, b = recs[g], ref[name]["Lineage_A"]
        ok_ref = b["MARGE_Summary"] is not None
        null_ref = b["Null_Summary"] is not None
        if bool(r.status & 1) != ok_ref or bool(r.status & 2) != null_ref:
            bad.append((name, "status", r.status, b["Model_Status"]))
            continue
        if null_ref and (rel(r.null_coef, b["Null_Summary"]["estimate"][0], 1e-6) > tol or rel(r.null_se, b["Null_Summary"]["std.error"][0], 1e-6) > tol
                         or rel(r.null_gee_phi, b["Null_Summary"]["phi"], 1e-6) > tol):
            bad.append((name, "null", r.null_coef, b["Null_Summary"]["estimate"][0]))
        if not ok_ref:
            continue
        if record_terms(r) != b["MARGE_Summary"]["term"]:
            bad.append((name, "terms", record_terms(r), b["MARGE_Summary"]["term"]))
            continue
        for j in range(r.n_terms):
            if rel(r.coef[j], b["MARGE_Summary"]["estimate"][j], 1e-6) > tol or rel(r.se[j], b["MARGE_Summary"]["std.error"][j], 1e-6) > tol:
                bad.append((name, "coef/se", j, r.coef[j], b["MARGE_Summary"]["estimate"][j]))
        for j in range(r.n_terms):
            if r.p[j] != b["MARGE_Summary"]["p.value"][j] and abs(r.p[j] - b["MARGE_Summary"]["p.value"][j]) > 2e-8:
                bad.append((name, "coef p", j, r.p[j], b["MARGE_Summary"]["p.value"][j]))
        if np.isnan(b["Test_Stat"]):                 # score test without a null model: NA statistic on both sides
            if not (np.isnan(r.test_stat) and np.isnan(r.df) and np.isnan(r.p_val)):
                bad.append((name, "test should be NA", r.test_stat, r.df))
        else:
            if rel(r.test_stat, b["Test_Stat"], 1e-6) > tol or r.df != b["Degrees_Freedom"]:
                bad.append((name, "test stat", r.test_stat, b["Test_Stat"]))
            if abs(r.p_val - b["P_Val"]) > tol * max(b["P_Val"], 1e-12) + 1e-15:
                bad.append((name, "p", r.p_val, b["P_Val"]))
        if rel(r.gee_alpha, b["alpha"], 1e-4) > 1e-5 or rel(r.gee_phi, b["phi"], 1e-6) > tol or r.gee_iters != b["niter"]:
            bad.append((name, "alpha/phi/iters", r.gee_alpha, b["alpha"], r.gee_phi, b["phi"], r.gee_iters, b["niter"]))
    return bad


@pytest.mark.parametrize("cor", ["ar1", "exchangeable", "independence"])
def test_gee_core_vs_dense_oracle(emu, cor):
    counts, pt, subject, sf = _gee_data()
    genes = [f"G{i}" for i in range(counts.shape[0])]
    recs, _ = emu.run_gee(counts, pt, subject, sf, cor)
    ref = oracle_gee(counts, pt, subject, genes, sf, cor)
    assert sum(ref[g]["Lineage_A"]["MARGE_Summary"] is not None for g in genes) >= 3
    assert _compare(recs, ref, genes) == []


def test_gee_core_no_offset_and_all_zero_gene(emu):
    counts, pt, subject, _ = _gee_data(seed=9, sizes=(100, 100, 100), sorted_within=True)
    counts[2] = 0
    genes = [f"G{i}" for i in range(counts.shape[0])]
    recs, _ = emu.run_gee(counts, pt, subject, None, "ar1")
    ref = oracle_gee(counts, pt, subject, genes, None, "ar1")
    assert recs[2].status == 0 and ref["G2"]["Lineage_A"]["Model_Status"] == "MARGE model error, null model error"
    assert _compare(recs, ref, genes) == []


def test_gee_core_bundled_genes(emu, golden_inputs):
    """Two genes of the bundled data set with its real subject structure (3 subjects x ~400 cells), AR-1."""
    sel = [1, 4]
    counts = golden_inputs["counts"][sel]
    genes = [golden_inputs["genes"][i] for i in sel]
    recs, _ = emu.run_gee(counts, golden_inputs["pt"], golden_inputs["subject"], golden_inputs["size_factor"], "ar1")
    ref = oracle_gee(counts, golden_inputs["pt"].reshape(-1), golden_inputs["subject"], genes, golden_inputs["size_factor"], "ar1")
    assert _compare(recs, ref, genes) == []


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("cor", ["ar1", "exchangeable", "independence"])
def test_gee_gpu_vs_dense_oracle(cor):
    import sclane_b200 as sb
    from sclane_b200 import _abi, api
    counts, pt, subject, sf = _gee_data()
    genes = [f"G{i}" for i in range(counts.shape[0])]
    o = api._opts(gpu_ids=[0])
    o.mode = 1
    o.gee_cor = _abi.COR_CODES[cor]
    n0 = sb.launch_count()
    recs, _ = api.run_records(counts, pt, sf, o, subject_id=subject.astype(np.int32))
    assert sb.launch_count() > n0
    ref = oracle_gee(counts, pt, subject, genes, sf, cor)
    assert _compare(recs, ref, genes) == []


@pytest.mark.gpu
@pytest.mark.parametrize("cor", ["ar1", "exchangeable"])
def test_gee_gpu_vs_core_bundled(emu, golden_inputs, cor):
    """All 100 bundled genes: CUDA path vs the same algorithm run serially on the host (fast), which the CPU tests pin to
    the dense oracle.  Also exercises the public testDynamic(is_gee = True) wrapper."""
    import sclane_b200 as sb
    counts, pt, subject, sf = golden_inputs["counts"], golden_inputs["pt"], golden_inputs["subject"], golden_inputs["size_factor"]
    res = sb.testDynamic(counts, pt, genes=golden_inputs["genes"], size_factor_offset=sf, is_gee=True, id_vec=subject,
                         cor_structure=cor, n_gpus=1)
    recs, _ = emu.run_gee(counts, pt, subject, sf, cor)
    n_ok = 0
    for g, name in enumerate(golden_inputs["genes"]):
        a, b = res.records[g], recs[g]
        assert a.status == b.status, name
        assert record_terms(a) == record_terms(b), name
        if a.status & 1:
            n_ok += 1
            for j in range(a.n_terms):
                assert rel(a.coef[j], b.coef[j], 1e-6) < TOL and rel(a.se[j], b.se[j], 1e-6) < TOL, name
            assert rel(a.test_stat, b.test_stat, 1e-6) < TOL and a.df == b.df, name
            assert a.gee_iters == b.gee_iters, name
        rec = res[name]["Lineage_A"]
        assert rec["Test_Stat_Type"] == "Wald"
    assert n_ok >= 80
    de = sb.getResultsDE(res)
    assert len(de["Gene"]) == 100
    # marge2(is.gee = TRUE) stand-alone on one gene == the same gene inside testDynamic (the final fit carries the offset)
    g = golden_inputs["genes"][1]
    one = sb.marge2(pt.reshape(-1), counts[1], Y_offset=sf, is_gee=True, id_vec=subject, cor_structure=cor)
    assert one["model_type"] == "GEE" and one["coef_names"] == res[g]["Lineage_A"]["MARGE_Summary"]["term"]
    assert np.allclose(list(one["final_mod"]["coefficients"].values()), res[g]["Lineage_A"]["MARGE_Summary"]["estimate"], rtol=1e-12)


def test_gee_score_test_core_vs_dense_oracle(emu):
    """gee.test = "score": scoreTestGEE() as one estimating-equation pass at the null fit vs the dense restatement."""
    counts, pt, subject, sf = _gee_data(seed=11, sizes=(140, 90, 130))
    genes = [f"G{i}" for i in range(counts.shape[0])]
    for cor in ("ar1", "exchangeable"):
        recs, _ = emu.run_gee(counts, pt, subject, sf, cor, gee_test="score")
        ref = oracle_gee(counts, pt, subject, genes, sf, cor, gee_test="score")
        assert _compare(recs, ref, genes) == []


@pytest.mark.gpu
def test_gee_score_test_gpu_vs_dense_oracle():
    from sclane_b200 import _abi, api
    counts, pt, subject, sf = _gee_data(seed=11, sizes=(140, 90, 130))
    genes = [f"G{i}" for i in range(counts.shape[0])]
    o = api._opts(gpu_ids=[0])
    o.mode = 1
    o.gee_cor = _abi.COR_AR1
    o.gee_test = 1
    recs, _ = api.run_records(counts, pt, sf, o, subject_id=subject.astype(np.int32))
    ref = oracle_gee(counts, pt, subject, genes, sf, "ar1", gee_test="score")
    assert _compare(recs, ref, genes) == []


@pytest.mark.parametrize("method,cor", [("kc", "ar1"), ("df", "exchangeable")])
def test_gee_bias_corrected_sandwich_core_vs_dense_oracle(emu, method, cor):
    """gee.bias.correction.method: every geem fit of the path carries geeM's sandwich variance (backward Wald, summaries) and
    waldTestGEE() uses biasCorrectGEE()'s scaled sandwich."""
    counts, pt, subject, sf = _gee_data(seed=13, sizes=(120, 100, 90, 110, 80), n_genes=5)
    genes = [f"G{i}" for i in range(counts.shape[0])]
    recs, _ = emu.run_gee(counts, pt, subject, sf, cor, bias_correction=method)
    ref = oracle_gee(counts, pt, subject, genes, sf, cor, bias_correction=method)
    assert sum(ref[g]["Lineage_A"]["MARGE_Summary"] is not None for g in genes) >= 3
    assert _compare(recs, ref, genes) == []


@pytest.mark.gpu
def test_gee_bias_corrected_sandwich_gpu_vs_dense_oracle():
    from sclane_b200 import _abi, api
    counts, pt, subject, sf = _gee_data(seed=13, sizes=(120, 100, 90, 110, 80), n_genes=5)
    genes = [f"G{i}" for i in range(counts.shape[0])]
    for method, code, cor in (("kc", 1, "ar1"), ("df", 2, "exchangeable")):
        o = api._opts(gpu_ids=[0])
        o.mode = 1
        o.gee_cor = _abi.COR_CODES[cor]
        o.gee_bias_correction = code
        recs, _ = api.run_records(counts, pt, sf, o, subject_id=subject.astype(np.int32))
        ref = oracle_gee(counts, pt, subject, genes, sf, cor, bias_correction=method)
        assert _compare(recs, ref, genes) == []


@pytest.mark.gpu
def test_gee_gpu_vs_core_at_config3_subject_size(emu):
    """BASELINE configs[2]'s shape per subject (3 subjects x 10,000 cells, AR-1; pseudotime-sorted within subject for the
    first half of the genes' run, arbitrary order for the second): k_gee vs the serial run of the same core."""
    from sclane_b200 import _abi, api
    from tools import synth
    counts, pt, _ = synth.make_counts(8, 30000, seed=23)
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    pt = np.asarray(pt, dtype=np.float64).reshape(-1)
    subject = (np.arange(30000) * 3 // 30000).astype(np.int32)
    sf = np.random.default_rng(23).uniform(0.6, 1.6, 30000)
    for order in ("sorted", "random"):
        if order == "sorted":
            idx = np.lexsort((pt, subject))
            c, x, s = np.ascontiguousarray(counts[:, idx]), pt[idx], sf[idx]
        else:
            c, x, s = counts, pt, sf
        o = api._opts(gpu_ids=[0])
        o.mode = 1
        o.gee_cor = _abi.COR_AR1
        recs, _ = api.run_records(c, x, s, o, subject_id=subject)
        ref, _ = emu.run_gee(c, x, subject, s, "ar1")
        n_ok = 0
        for i in range(8):
            a, b = recs[i], ref[i]
            assert a.status == b.status and record_terms(a) == record_terms(b), (order, i)
            if a.status & 1:
                n_ok += 1
                for j in range(a.n_terms):
                    assert rel(a.coef[j], b.coef[j], 1e-6) < TOL and rel(a.se[j], b.se[j], 1e-6) < TOL, (order, i, j)
                assert rel(a.test_stat, b.test_stat, 1e-6) < TOL and rel(a.gee_phi, b.gee_phi) < TOL, (order, i)
        assert n_ok >= 6


# ---------------------------------------------------------------------------------------------------------------
# Round 2: the CUDA path against the INDEPENDENT dense oracle at subject sizes where the AR-1 / exchangeable conditioning
# (alpha -> 1, chains of > 1,000 cells) can bite.  The dense oracle builds and inverts n_i x n_i matrices like the reference
# (O(n_i^3)): ~75 s per gene at 3,300 cells, so the genes are spread over the box's cores.
# ---------------------------------------------------------------------------------------------------------------
def _oracle_gee_one(a):
    from threadpoolctl import threadpool_limits
    counts, pt, subject, gene, sf, cor = a
    with threadpool_limits(limits=1):
        return oracle_gee(counts, pt, subject, [gene], sf, cor)


def _correlated_counts(n_genes, sizes, seed, rho):
    """NB counts whose mean carries a slowly varying AR(1) log-factor per subject (lag-1 correlation rho): the working
    correlation the geem fits estimate comes out close to 1."""
    from tools import synth
    n = int(sum(sizes))
    rng = np.random.default_rng(seed)
    pt = np.concatenate([np.sort(rng.random(s)) for s in sizes])            # cells ordered by pseudotime inside each subject
    subject = np.repeat(np.arange(len(sizes)), sizes)
    counts = np.empty((n_genes, n), dtype=np.int32)
    for g in range(n_genes):
        z = np.empty(n)
        start = 0
        for s in sizes:
            e = rng.normal(0.0, 1.0, s)
            zz = np.empty(s)
            zz[0] = e[0]
            for j in range(1, s):
                zz[j] = rho * zz[j - 1] + np.sqrt(1.0 - rho * rho) * e[j]
            z[start:start + s] = zz
            start += s
        eta = rng.normal(3.0, 0.3) + rng.normal(0.0, 2.0) * np.maximum(pt - rng.uniform(0.2, 0.8), 0.0) + 1.0 * z
        counts[g] = rng.poisson(np.exp(np.clip(eta, -6, 7)))
    del synth
    return counts, pt, subject


@pytest.mark.gpu
@pytest.mark.parametrize("cor,rho", [("ar1", 0.0), ("ar1", 0.995), ("exchangeable", 0.0)])
def test_gee_gpu_vs_dense_oracle_large_subjects(cor, rho):
    import os
    from concurrent.futures import ProcessPoolExecutor
    from sclane_b200 import _abi, api
    sizes = (700, 1100, 1500)                                            # unequal subjects, 3,300 cells
    n_genes = min(12, max(4, (os.cpu_count() or 4)))
    if rho > 0:
        counts, pt, subject = _correlated_counts(n_genes, sizes, 41, rho)
        sf = None
    else:
        counts, pt, subject, sf = _gee_data(seed=43, sizes=sizes, n_genes=n_genes, sorted_within=True)
    genes = [f"G{i}" for i in range(n_genes)]
    o = api._opts(gpu_ids=[0])
    o.mode = 1
    o.gee_cor = _abi.COR_CODES[cor]
    recs, _ = api.run_records(counts, pt, sf, o, subject_id=subject.astype(np.int32))
    ref = {}
    with ProcessPoolExecutor(max_workers=min(os.cpu_count() or 1, n_genes)) as ex:
        for part in ex.map(_oracle_gee_one, [(counts[i:i + 1], pt, subject, genes[i], sf, cor) for i in range(n_genes)]):
            ref.update(part)
    assert _compare(recs, ref, genes) == []
    alphas = [r.gee_alpha for r in recs if r.status & 1]
    if rho > 0.9:
        assert max(alphas) > 0.9, alphas                                  # the case the test is about: alpha close to 1
    print(f"{cor} rho={rho}: fitted alpha range {min(alphas):.3f} .. {max(alphas):.3f}, genes with a fitted model {len(alphas)}")


def test_gee_tests_are_na_when_the_null_model_failed(emu):
    """waldTestGEE.R:31-37 / scoreTestGEE.R:34-39: Wald_Stat, DF and P_Val are NA when EITHER model is a try-error (the test
    must not reach getResultsDE's FDR adjustment with a real-looking p-value)."""
    import ctypes as C
    from sclane_b200 import _abi
    for gee_test in (0, 1):
        rec = _abi.SclaneRecord()
        emu.emu_finalize_gee_null_failed.argtypes = [C.c_int, C.POINTER(_abi.SclaneRecord)]
        emu.emu_finalize_gee_null_failed(gee_test, C.byref(rec))
        assert rec.status == 1 and rec.n_terms == 3                       # MARGE model OK, null model error
        assert np.isnan(rec.test_stat) and np.isnan(rec.df) and np.isnan(rec.p_val)
        assert rec.coef[1] == 0.5 and np.isnan(rec.null_coef)
