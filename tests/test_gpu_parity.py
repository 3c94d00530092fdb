"""Parity of the CUDA path (called through the C ABI) with the oracle and with the reference's golden run.
Everything here needs a B200 (pytest -m gpu)."""
import copy

import numpy as np
import pytest

from conftest import GOLDEN_EXCEPTIONS, record_terms, rel
from oracle import marge as omarge
from oracle.eigen_helpers import eigen_map_mat_mult, eigen_map_matrix_invert, eigen_map_pseudo_inverse
from oracle.knots import candidate_knots
from oracle.nbfit import nb_mle
from oracle.rmath import p_adjust_bh
from oracle.testdynamic import test_dynamic as oracle_test_dynamic

pytestmark = pytest.mark.gpu

TOL = 1e-6   # north_star: coefficients, theta and LRT within 1e-6 relative in fp64


def _compare_with_oracle(res, oracle, tol=TOL, theta=True):
    worst = {}
    for g, lins in res.items():
        for ln, r in lins.items():
            o = oracle[g][ln]
            assert r["Model_Status"] == o["Model_Status"], (g, ln)
            if o["MARGE_Summary"] is None:
                assert r["MARGE_Summary"] is None
                continue
            assert r["MARGE_Summary"]["term"] == o["MARGE_Summary"]["term"], (g, ln)
            assert r["Degrees_Freedom"] == o["Degrees_Freedom"]
            e = {}
            if abs(o["Test_Stat"]) > 1e-6:
                e["lrt"] = rel(r["Test_Stat"], o["Test_Stat"])
            else:
                assert abs(r["Test_Stat"]) < 1e-6
            e["coef"] = max(rel(a, b) for a, b in zip(r["MARGE_Summary"]["estimate"], o["MARGE_Summary"]["estimate"]))
            e["se"] = max(rel(a, b) for a, b in zip(r["MARGE_Summary"]["std.error"], o["MARGE_Summary"]["std.error"]))
            e["ll"] = rel(r["LogLik_MARGE"], o["LogLik_MARGE"])
            e["ll0"] = rel(r["LogLik_Null"], o["LogLik_Null"])
            e["dev"] = rel(r["Dev_MARGE"], o["Dev_MARGE"])
            e["dev0"] = rel(r["Dev_Null"], o["Dev_Null"])
            e["coef0"] = rel(r["Null_Summary"]["estimate"][0], o["Null_Summary"]["estimate"][0])
            rec = r["record"]
            if theta and not (rec.marge_notes & 3) and not (rec.null_notes & 3):   # theta only when glm.nb converged (no th.warn)
                e["theta"] = max(rel(r["Theta"][0], o["_theta"][0]), rel(r["Theta"][1], o["_theta"][1]))
            for k, v in e.items():
                if v > worst.get(k, (0.0, ""))[0]:
                    worst[k] = (v, g)
    bad = {k: v for k, v in worst.items() if v[0] > tol}
    assert not bad, bad
    return worst


@pytest.fixture(scope="module")
def gpu_golden(golden_inputs):
    import sclane_b200 as sb
    return sb.testDynamic(golden_inputs["counts"], golden_inputs["pt"], genes=golden_inputs["genes"],
                          size_factor_offset=golden_inputs["size_factor"], n_gpus=1)


@pytest.fixture(scope="module")
def oracle_golden(golden_inputs):
    return oracle_test_dynamic(golden_inputs["counts"], golden_inputs["pt"], golden_inputs["genes"], golden_inputs["size_factor"],
                               keep_preds=False)


def test_bundled_vs_oracle(gpu_golden, oracle_golden):
    """C1 (bundled sim_counts + sim_pseudotime, GLM + offset): CUDA == oracle on every gene, discrete and continuous."""
    _compare_with_oracle(gpu_golden, oracle_golden)
    import sclane_b200 as sb
    assert sb.launch_count() >= 7


def test_bundled_vs_reference_golden(gpu_golden, golden_models, golden_preds):
    """CUDA vs the reference's own bundled testDynamic() output: identical term sets outside the 11 stated
    platform-dependent genes, statistics within 1e-6, identical dynamic/static calls at FDR 0.01 for all 100 genes."""
    import sclane_b200 as sb
    mism = set()
    for g, gold in golden_models.items():
        r = gpu_golden[g]["Lineage_A"]
        assert r["Model_Status"] == gold["Model_Status"]
        if r["MARGE_Summary"]["term"] != gold["MARGE_Summary"]["term"]:
            mism.add(g)
            continue
        assert rel(r["Test_Stat"], gold["Test_Stat"], 1e-6) < TOL
        assert r["Degrees_Freedom"] == gold["Degrees_Freedom"]
        assert rel(r["P_Val"], gold["P_Val"], 1e-300) < 1e-5
        for k in ("estimate", "std.error", "statistic"):
            for a, b in zip(r["MARGE_Summary"][k], gold["MARGE_Summary"][k]):
                assert rel(a, b) < TOL, (g, k)
        assert rel(r["LogLik_MARGE"], gold["LogLik_MARGE"]) < 1e-8 and rel(r["LogLik_Null"], gold["LogLik_Null"]) < 1e-8
        assert rel(r["Dev_MARGE"], gold["Dev_MARGE"]) < TOL and rel(r["Dev_Null"], gold["Dev_Null"]) < TOL
        assert r["MARGE_Slope_Data"]["Direction"] == gold["MARGE_Slope_Data"]["Direction"]
        assert r["MARGE_Slope_Data"]["Rounded_Breakpoint"] == gold["MARGE_Slope_Data"]["Rounded_Breakpoint"]
    assert mism == GOLDEN_EXCEPTIONS
    de = sb.getResultsDE(gpu_golden)
    calls = dict(zip(de["Gene"], de["Gene_Dynamic_Overall"]))
    gadj = p_adjust_bh(np.array([golden_models[g]["P_Val"] for g in golden_models]))
    assert calls == {g: int(a < 0.01) for g, a in zip(golden_models, gadj)}
    # per-cell predictions (lazy MARGE_Preds / Null_Preds) vs the golden data.frames
    n = 0
    for g in sorted({k.split("/")[0] for k in golden_preds.files} - GOLDEN_EXCEPTIONS):
        r = gpu_golden[g]["Lineage_A"]
        for key, val in (("marge_link_fit", r["MARGE_Preds"]["marge_link_fit"]), ("marge_link_se", r["MARGE_Preds"]["marge_link_se"]),
                         ("null_link_fit", r["Null_Preds"]["null_link_fit"]), ("null_link_se", r["Null_Preds"]["null_link_se"])):
            gold = golden_preds[f"{g}/{key}"]
            assert np.max(np.abs(val - gold) / (np.abs(gold) + 1e-9)) < TOL, (g, key)
        n += 1
    assert n >= 8


def test_synthetic_vs_oracle_with_edge_cases():
    """Seeded synthetic NB counts incl. an all-zero gene, a nearly empty gene, a huge-count gene and a near-Poisson gene."""
    import sclane_b200 as sb
    from tools import synth
    counts, pt, _ = synth.make_counts(40, 3000, seed=11)
    rng = np.random.default_rng(5)
    counts[3] = 0
    counts[5] = (np.arange(3000) % 211 == 0).astype(np.int32)
    counts[7] = rng.poisson(3000.0 * np.exp(1.5 * pt[:, 0]))          # large counts, Poisson
    counts[9] = rng.poisson(2.0, 3000)                                  # pure Poisson, static
    genes = [f"g{i}" for i in range(40)]
    res = sb.testDynamic(counts, pt, genes=genes, n_gpus=1)
    oracle = oracle_test_dynamic(counts, pt, genes, None, keep_preds=False)
    assert res["g3"]["Lineage_A"]["Model_Status"] == "MARGE model error, null model error"
    _compare_with_oracle(res, oracle)


def test_two_lineages_with_offsets_vs_oracle():
    """C4-shaped: two lineages sharing a root (NaN = not in lineage), size-factor offsets from createCellOffset()."""
    import sclane_b200 as sb
    from tools import synth
    counts, pt, _ = synth.make_counts(24, 4000, seed=21, n_lineages=2)
    sf = sb.createCellOffset(counts)
    genes = [f"g{i}" for i in range(24)]
    res = sb.testDynamic(counts, pt, genes=genes, size_factor_offset=sf, n_gpus=1)
    oracle = oracle_test_dynamic(counts, pt, genes, sf, keep_preds=False)
    assert set(res["g0"].keys()) == {"Lineage_A", "Lineage_B"}
    _compare_with_oracle(res, oracle)
    assert res.lineage_info[0]["n_cells"] == int(np.sum(~np.isnan(pt[:, 0])))


def test_null_stats_and_score_sweep_vs_dense_oracle(golden_inputs):
    """K1 (stat_out_score_glm_null) and K2 (score_fun_glm over all knots x {one, both}) through their single-call
    entry points, against the dense numpy restatement -- all 50 scores per gene, not just the arg-max."""
    import sclane_b200 as sb
    X, knots = candidate_knots(golden_inputs["pt"])
    sel = [1, 3, 10, 17, 42, 63, 77, 90]
    counts = golden_inputs["counts"][sel]
    for kinds, kidx in (([0], [-1]), ([0, 1, 2], [-1, 12, 12]), ([0, 1], [-1, 5]), ([0, 1, 2, 1], [-1, 18, 18, 3])):
        beta, sigma, mu = sb.stat_out_score_glm_null(counts, golden_inputs["pt"], knots, kinds, kidx)
        terms = [omarge.Term(k, i) for k, i in zip(kinds, kidx)]
        B = omarge.build_design(terms, X, knots)
        for gi in range(len(sel)):
            y = counts[gi].astype(np.float64)
            fit = nb_mle(B, y)
            assert (sigma[gi] == 0.0) == fit.poisson
            assert rel(sigma[gi], fit.sigma, 1e-12) < 1e-6
            assert np.max(np.abs(beta[gi] - fit.beta) / (np.abs(fit.beta) + 1e-6)) < 1e-6
            assert np.max(np.abs(mu[gi] - fit.mu) / fit.mu) < 1e-6
        # sweep with the oracle's mu so that only the sweep itself is compared
        mus = np.stack([nb_mle(B, counts[gi].astype(np.float64)).mu for gi in range(len(sel))])
        sig = np.array([nb_mle(B, counts[gi].astype(np.float64)).sigma for gi in range(len(sel))])
        scores = sb.score_fun_glm_sweep(counts, mus, sig, golden_inputs["pt"], knots, kinds, kidx)
        rng = np.random.default_rng(0)
        for gi in range(len(sel)):
            y = counts[gi].astype(np.float64)
            ns = omarge.stat_out_score_glm_null(y, B)
            # conditioning probe: the same dense evaluation with mu perturbed by 1e-15 relative.  A candidate that is
            # nearly collinear with the basis amplifies rounding by up to 1e10 in the Schur complement D - C'AC, in the
            # dense and in the bucket-moment evaluation alike; its score is only defined to that spread.
            ns_p = copy.copy(ns)
            ns_p.mu = ns.mu * (1.0 + 1e-15 * rng.standard_normal(y.size))
            ns_p.V = ns_p.mu * (1.0 + ns_p.mu * ns.sigma)
            ns_p.VS = (y - ns_p.mu) / ns_p.V
            w_p = ns_p.mu * ns_p.mu / ns_p.V
            ns_p.B1 = B.T * w_p[None, :]
            Ai = ns_p.B1 @ B
            ns_p.A = np.linalg.inv(Ai) if Ai.shape[0] > 1 else 1.0 / Ai
            for t in range(knots.size):
                h1, h2 = omarge.tp1(X, knots[t]), omarge.tp2(X, knots[t])
                for row, XA, kinds_c in ((0, h1[:, None], (1,)), (1, np.stack([h1, h2], 1), (1, 2))):
                    na = omarge._collinear(terms, kinds_c, t)
                    ref = omarge.score_fun_glm(ns, XA, na)
                    mine = scores[gi, row, t]
                    if np.isnan(ref):
                        assert np.isnan(mine)
                        continue
                    spread = abs(omarge.score_fun_glm(ns_p, XA, na) - ref)
                    assert abs(mine - ref) < 1e-7 * max(1.0, abs(ref)) + 50.0 * spread, (gi, t, row, mine, ref, spread)


def test_eigen_dropins():
    """The three RcppEigen helpers (tests/testthat/test_scLANE.R:29-38 checks them on a random 25 x 25 PD matrix)."""
    import sclane_b200 as sb
    rng = np.random.default_rng(1)
    M = rng.normal(size=(25, 25))
    A = M @ M.T + 25 * np.eye(25)
    assert np.allclose(sb.eigenMapMatrixInvert(A), eigen_map_matrix_invert(A), rtol=1e-10, atol=1e-12)
    assert np.allclose(sb.eigenMapMatrixInvert(A), np.linalg.inv(A), rtol=1e-8)
    B = rng.normal(size=(25, 7))
    assert np.allclose(sb.eigenMapMatMult(A, B), eigen_map_mat_mult(A, B), rtol=1e-12, atol=1e-12)
    assert np.allclose(sb.eigenMapPseudoInverse(A), eigen_map_pseudo_inverse(A), rtol=1e-8, atol=1e-10)
    R = rng.normal(size=(30, 4)) @ rng.normal(size=(4, 12))        # rank 4, tall
    assert np.allclose(sb.eigenMapPseudoInverse(R), eigen_map_pseudo_inverse(R), rtol=1e-7, atol=1e-9)
    assert np.allclose(sb.eigenMapPseudoInverse(R.T), eigen_map_pseudo_inverse(R.T), rtol=1e-7, atol=1e-9)
    # non-PD input: Eigen's LLT stops early and solve() runs on the untouched remainder
    N = np.array([[4.0, 2.0, 1.0], [2.0, 0.5, 3.0], [1.0, 3.0, 2.0]])
    assert np.allclose(sb.eigenMapMatrixInvert(N), eigen_map_matrix_invert(N), rtol=1e-12, equal_nan=True)
    assert np.allclose(sb.eigenMapMatrixInvert(np.array([[-2.0]])), [[0.25]])


def _rec_bytes(rec) -> bytes:
    """Record bytes with the timing fields (gene_time: wall-clock surrogate; stage_cycles: clock64 profile) masked out."""
    from sclane_b200 import _abi
    b = bytearray(bytes(rec))
    for fld, n in ((_abi.SclaneRecord.gene_time, 8), (_abi.SclaneRecord.stage_cycles, 24)):
        b[fld.offset:fld.offset + n] = b"\0" * n
    return bytes(b)


def test_full_size_properties():
    """Size-independent properties at BASELINE config 2's cell count (10,000 cells; a gene subset keeps it quick):
    gene-order independence and batching independence are bit-exact; a permutation of the cells leaves every
    statistic unchanged to rounding; LRT >= 0 whenever the alternative nests the null; FDR calls are a function of P."""
    import sclane_b200 as sb
    from tools import synth
    counts, pt, _ = synth.make_counts(192, 10000, seed=312)
    base = sb.testDynamic(counts, pt, n_gpus=1)
    recs = base.records
    # (1) batching: 192 genes in one batch vs batches of 50
    small = sb.testDynamic(counts, pt, n_gpus=1, genes_per_batch=50)
    for i in range(192):
        assert _rec_bytes(recs[i]) == _rec_bytes(small.records[i])       # everything but gene_time
    # (2) gene order
    perm = np.random.default_rng(0).permutation(192)
    shuf = sb.testDynamic(counts[perm], pt, n_gpus=1)
    for j, i in enumerate(perm):
        assert _rec_bytes(shuf.records[j]) == _rec_bytes(recs[int(i)])
    # (3) cell permutation
    cp = np.random.default_rng(1).permutation(10000)
    pc = sb.testDynamic(np.ascontiguousarray(counts[:, cp]), pt[cp], n_gpus=1)
    for i in range(192):
        a, b = recs[i], pc.records[i]
        assert a.status == b.status and record_terms(a) == record_terms(b)
        if a.status == 3:
            assert rel(a.test_stat, b.test_stat, 1e-6) < 1e-7 and rel(a.loglik, b.loglik) < 1e-10
    # (4) sanity of the statistics
    for i in range(192):
        r = recs[i]
        if r.status == 3:
            assert r.test_stat > -1e-6 * abs(r.loglik)
            assert 0.0 <= r.p_val <= 1.0
            assert r.df == max(1, r.n_terms - 1)
            assert r.fwd_n_terms in (1, 3, 4) or r.fwd_n_terms == 2
    de = sb.getResultsDE(base)
    assert ((de["P_Val_Adj"] < 0.01).astype(int) == de["Gene_Dynamic_Lineage"]).all()
    assert 40 < int(de["Gene_Dynamic_Overall"].sum()) <= 192


def test_compressed_path_vs_per_cell_path():
    """The X-compressed kernels (default when 4-dp rounding leaves <= 0.8 N abscissae) and the per-cell kernels run the
    same algorithm: identical discrete choices, statistics within the fp64 tolerance.  Gene 6 carries a few counts >=
    kHistCap (4096; kept as a sorted list), gene 5 more than kBigCap (8192) of them, which routes it to the per-cell
    kernels inside the compressed run."""
    import sclane_b200 as sb
    from tools import synth
    counts, pt, _ = synth.make_counts(96, 20000, seed=21)
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    counts[5, :9000] = 4100 + (np.arange(9000) % 900)
    counts[6, 100:130] += 6000
    sf = sb.createCellOffset(counts)
    for offs in (None, sf):
        a = sb.testDynamic(counts, pt, size_factor_offset=offs, n_gpus=1, preds="none")
        b = sb.testDynamic(counts, pt, size_factor_offset=offs, n_gpus=1, preds="none", force_per_cell=True)
        assert a.profile["ms"]["comp_forward"] > 0 and b.profile["ms"]["comp_forward"] == 0
        assert a.records[5].fit_passes == b.records[5].fit_passes          # the routed gene ran the very same kernels
        for i in range(96):
            x, y = a.records[i], b.records[i]
            assert x.status == y.status and record_terms(x) == record_terms(y), i
            assert x.fwd_n_terms == y.fwd_n_terms and list(x.fwd_knot_idx) == list(y.fwd_knot_idx), i
            if x.status == 3:
                for j in range(x.n_terms):
                    assert rel(x.coef[j], y.coef[j], 1e-6) < 1e-6 and rel(x.se[j], y.se[j], 1e-6) < 1e-6, i
                assert rel(x.theta, y.theta) < 1e-6 and rel(x.loglik, y.loglik) < 1e-9, i
                assert abs(x.test_stat - y.test_stat) < 1e-6 * (abs(y.test_stat) + 1e-3), i


def test_negative_counts_fail_the_gene_only():
    """A negative count makes both models of that gene fail (glm.nb: 'negative values not allowed'), on every route, and
    leaves the other genes untouched."""
    import sclane_b200 as sb
    from sclane_b200 import _abi
    from tools import synth
    counts, pt, _ = synth.make_counts(12, 3000, seed=4)
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    ref = sb.testDynamic(counts, pt, n_gpus=1, preds="none")
    bad = counts.copy()
    bad[3, 77] = -2
    subject = np.arange(3000) * 3 // 3000
    for kw in ({}, {"force_per_cell": True}, {"is_gee": True, "id_vec": subject}):
        res = sb.testDynamic(bad, pt, n_gpus=1, preds="none", **kw)
        r = res.records[3]
        assert r.status == 0 and (r.marge_notes & _abi.NOTE_BAD_COUNTS) and (r.null_notes & _abi.NOTE_BAD_COUNTS)
        assert res.records[4].status == 3
    res = sb.testDynamic(bad, pt, n_gpus=1, preds="none")
    for i in range(12):
        if i != 3:
            assert _rec_bytes(res.records[i]) == _rec_bytes(ref.records[i])


def test_smoothed_counts_and_fitted_values_vs_golden(gpu_golden, golden_inputs, golden_preds):
    """smoothedCountsMatrix() / getFittedValues() (the per-cell consumers of the fitted models, smoothedCountsMatrix.R:27-74,
    getFittedValues.R:45-150) against the reference's bundled per-cell predictions: exp(marge_link_fit) * size factor."""
    import sclane_b200 as sb
    sf = golden_inputs["size_factor"]
    sm = sb.smoothedCountsMatrix(gpu_golden, size_factor_offset=sf, pt=golden_inputs["pt"])
    mat, names = sm["Lineage_A"]["matrix"], sm["Lineage_A"]["genes"]
    assert mat.shape == (golden_inputs["pt"].shape[0], len(names)) and len(names) == 100
    checked = sorted({k.split("/")[0] for k in golden_preds.files} - GOLDEN_EXCEPTIONS)
    for g in checked:
        gold = np.exp(golden_preds[f"{g}/marge_link_fit"]) * sf
        assert np.max(np.abs(mat[:, names.index(g)] - gold) / gold) < TOL, g
    lg = sb.smoothedCountsMatrix(gpu_golden, size_factor_offset=sf, pt=golden_inputs["pt"], genes=checked[:3], log1p_norm=True)
    assert np.allclose(lg["Lineage_A"]["matrix"], np.log1p(mat[:, [names.index(g) for g in checked[:3]]]), rtol=1e-12)
    fv = sb.getFittedValues(gpu_golden, genes=checked[:2], pt=golden_inputs["pt"], expr_mat=golden_inputs["counts"], size_factor_offset=sf)
    assert len(fv) == 2 * golden_inputs["pt"].shape[0]
    sub = fv[fv["gene"] == checked[0]]
    gold = np.exp(golden_preds[f"{checked[0]}/marge_link_fit"]) * sf
    assert np.max(np.abs(sub["scLANE_pred"].to_numpy() - gold) / gold) < TOL
    assert np.all(sub["scLANE_ci_ll"].to_numpy() <= sub["scLANE_pred"].to_numpy())


def test_sparse_csc_input_equals_dense():
    """sclane_test_dynamic_csc (genes x cells dgCMatrix layout, expanded on the device) gives byte-identical records to the
    dense entry point, with offsets and two lineages; bad indices are rejected."""
    import scipy.sparse as sp
    import sclane_b200 as sb
    from sclane_b200 import _abi, api
    from tools import synth
    counts, pt, _ = synth.make_counts(40, 6000, seed=8)
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    counts[:, ::3] = 0                                           # sparser than the generator makes it
    pt = np.asarray(pt, dtype=np.float64).reshape(-1)
    pt2 = np.stack([np.where(np.arange(6000) % 4 == 0, np.nan, pt), np.where(np.arange(6000) % 7 == 0, np.nan, 1.0 - pt)], 1)
    m = sp.csc_matrix(counts)
    assert np.allclose(sb.createCellOffset(m), sb.createCellOffset(counts))
    sf = sb.createCellOffset(counts)
    for offs in (None, sf):
        a = sb.testDynamic(counts, pt2, size_factor_offset=offs, n_gpus=1, preds="none")
        b = sb.testDynamic(m, pt2, size_factor_offset=offs, n_gpus=1, preds="none")
        assert len(b.records) == len(a.records) == 80
        for i in range(80):
            assert _rec_bytes(a.records[i]) == _rec_bytes(b.records[i]), i
    # a row index outside [0, n_genes) is a global argument error, reported through errbuf
    import ctypes as C
    from sclane_b200._lib import lib
    indptr = np.ascontiguousarray(m.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(m.indices, dtype=np.int32).copy()
    data = np.ascontiguousarray(m.data, dtype=np.float64)
    indices[5] = 40
    o = api._opts(gpu_ids=[0])
    recs = (_abi.SclaneRecord * 80)()
    err = C.create_string_buffer(512)
    ptc = np.ascontiguousarray(pt2.T)
    rc = lib().sclane_test_dynamic_csc(indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, 6000, 40, ptc.ctypes.data, 2, None,
                                       C.byref(o), recs, None, err, 512)
    assert rc == 1 and b"out of range" in err.value


def test_default_workload_cell_count_vs_oracle():
    """BASELINE configs[4]'s cell count (200,000 cells, 10,000 distinct abscissae -> X-compressed route): six genes of the
    bench generator against the dense oracle (about 7 s of numpy per gene)."""
    import sclane_b200 as sb
    from tools import synth
    counts, pt, _ = synth.make_counts(6, 200000, seed=312)
    res = sb.testDynamic(counts, pt, n_gpus=1, preds="none")
    assert res.profile["ms"]["comp_final"] > 0
    ref = oracle_test_dynamic(np.asarray(counts), np.asarray(pt, dtype=np.float64).reshape(-1), list(res.keys()), None, keep_preds=False)
    _compare_with_oracle(res, ref)


def test_staged_uint16_transfer_equals_direct():
    """h2d_mode = 2 (host threads narrow every gene batch to uint16 into pinned staging, double buffered) gives byte-identical
    records to the direct int32 copy, on both routes; a batch holding a count > 65534 or a negative count falls back to int32."""
    from sclane_b200 import api
    from tools import synth
    counts, pt, _ = synth.make_counts(90, 9000, seed=17)
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    counts[33, 100] = 70000            # does not fit uint16 -> its batch goes direct
    counts[61, 5] = -1                 # negative -> its batch goes direct, the gene fails
    for per_cell in (False, True):
        out = []
        for mode in (1, 2):
            o = api._opts(gpu_ids=[0], genes_per_batch=16, force_per_cell=per_cell, h2d_mode=mode)
            recs, _ = api.run_records(counts, pt, None, o)
            out.append(recs)
            p = api.get_profile()
            # 90 genes in equal batches of 15 (genes_per_batch = 16 -> 6 batches); genes 33 and 61 sit in two different batches
            assert p["h2d_bytes"] == (90 * 9000 * 4 if mode == 1 else (90 - 30) * 9000 * 2 + 30 * 9000 * 4)
        for i in range(90):
            assert _rec_bytes(out[0][i]) == _rec_bytes(out[1][i]), (per_cell, i)
        assert out[1][61].status == 0 and out[1][33].status == 3


@pytest.mark.parametrize("M", [3, 7])
def test_other_basis_budgets_vs_oracle(M):
    """n.potential.basis.fns = 3 (one forward iteration) and 7 (three iterations, up to seven columns -- SCLANE_PMAX), both
    routes, against the oracle."""
    import sclane_b200 as sb
    from tools import synth
    counts, pt, _ = synth.make_counts(20, 16000, seed=29)
    ref = oracle_test_dynamic(np.asarray(counts), np.asarray(pt, dtype=np.float64).reshape(-1), None, None, n_potential_basis_fns=M,
                              keep_preds=False)
    genes = list(ref.keys())
    for per_cell in (False, True):
        res = sb.testDynamic(counts, pt, genes=genes, n_potential_basis_fns=M, n_gpus=1, preds="none", force_per_cell=per_cell)
        _compare_with_oracle(res, ref)
    sizes = {res.records[i].n_terms for i in range(20)}
    assert max(sizes) <= M and (M == 3 or max(sizes) >= 4)
