"""Conformance of the R shim's C half (r/sclane_b200_shim.c): compiled against the stand-in for R's C API in tests/c/rstub/,
driven by tests/c/shim_conformance.c, its returned R object (dumped as JSON) must reproduce the 21-field structure of
testDynamic.R:320-342 that the reference's own bundled output (data/scLANE_models.rda -> tests/golden/sclane_models.json) holds."""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

import __graft_entry__ as ge
from conftest import GOLDEN_EXCEPTIONS, rel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CDIR = os.path.join(ROOT, "tests", "c")
EXE = os.path.join(CDIR, "shim_conformance")
FIELDS = ["Gene", "Lineage", "Test_Stat", "Test_Stat_Type", "Test_Stat_Note", "Degrees_Freedom", "P_Val", "LogLik_MARGE", "LogLik_Null",
          "Dev_MARGE", "Dev_Null", "Model_Status", "MARGE_Fit_Notes", "Null_Fit_Notes", "Gene_Time", "MARGE_Summary", "Null_Summary",
          "MARGE_Preds", "Null_Preds", "MARGE_Slope_Data", "Gene_Dynamics"]


def _build():
    ge.build()
    inc = ["-I", os.path.join(CDIR, "rstub"), "-I", os.path.join(ROOT, "include")]
    objs = []
    for src in (os.path.join(ROOT, "r", "sclane_b200_shim.c"), os.path.join(CDIR, "rstub", "rstub.c"), os.path.join(CDIR, "shim_conformance.c")):
        obj = os.path.join(tempfile.gettempdir(), os.path.basename(src) + ".o")
        r = subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", *inc, "-c", src, "-o", obj], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        objs.append(obj)
    r = subprocess.run(["gcc", *objs, "-o", EXE, "-L", ge.LIBDIR, "-lsclane_b200", "-lm", f"-Wl,-rpath,{ge.LIBDIR}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_shim_compiles_against_the_r_api_stub_and_registers_every_entry_point():
    """CPU: the shim builds with the R headers replaced by the stub, and every .Call entry point INTEGRATION.md names is registered."""
    _build()
    src = open(os.path.join(ROOT, "r", "sclane_b200_shim.c")).read()
    integ = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for name in ("C_sclane_test_dynamic", "C_sclane_test_dynamic_csc", "C_sclane_test_dynamic_gee", "C_sclane_marge_basis", "C_sclane_link_preds",
                 "C_sclane_smoothed_counts", "C_eigenMapMatMult", "C_eigenMapMatrixInvert", "C_eigenMapPseudoInverse"):
        assert f'{{"{name}", (DL_FUNC)&{name},' in src, name
        assert name in integ, name
    # every helper the R wrappers call exists in r/
    rsrc = "".join(open(os.path.join(ROOT, "r", f)).read() for f in os.listdir(os.path.join(ROOT, "r")) if f.endswith(".R"))
    import re
    called = set(re.findall(r"\.Call\(\"(C_[A-Za-z_]+)\"", rsrc))
    assert called and all(f'{{"{c}"' in src for c in called), called
    for fn in re.findall(r"(?<![.\w])(\.?sclane_[a-z_]+)\(", rsrc):
        assert re.search(rf"^{re.escape(fn)} <- function", rsrc, re.M), fn


@pytest.mark.gpu
def test_shim_c_half_reproduces_the_reference_golden_structure(golden_inputs, golden_models):
    _build()
    counts, pt, sf, genes = golden_inputs["counts"], golden_inputs["pt"], golden_inputs["size_factor"], golden_inputs["genes"]
    with tempfile.TemporaryDirectory() as d:
        pre = os.path.join(d, "c1")
        np.ascontiguousarray(counts, dtype=np.int32).tofile(pre + ".counts")           # gene-major == column-major cells x genes
        np.ascontiguousarray(pt.reshape(-1), dtype=np.float64).tofile(pre + ".pt")
        np.ascontiguousarray(sf, dtype=np.float64).tofile(pre + ".sf")
        with open(pre + ".genes", "w") as fh:
            fh.write("\n".join(genes) + "\n")
        r = subprocess.run([EXE, pre, str(counts.shape[1]), str(counts.shape[0]), "1", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout)
    assert res["_class"] == ["scLANE"] and [g for g in res if g != "_class"] == genes
    n = 0
    for g in genes:
        rec, gold = res[g]["Lineage_A"], golden_models[g]
        assert [k for k in rec if k != "_class"] == FIELDS, g                           # names AND order (getResultsDE takes fields 1-15)
        for k in ("Gene", "Lineage", "Test_Stat_Type", "Model_Status"):
            assert rec[k] == [gold[k]], (g, k)
        if g in GOLDEN_EXCEPTIONS:
            continue
        assert rec["MARGE_Summary"]["term"] == gold["MARGE_Summary"]["term"], g
        assert rec["MARGE_Summary"]["_class"] == ["data.frame"]
        for k in ("estimate", "std.error", "statistic"):
            for a, b in zip(rec["MARGE_Summary"][k], gold["MARGE_Summary"][k]):
                assert rel(a, b) < 1e-6, (g, k)
        for k in ("Test_Stat", "LogLik_MARGE", "LogLik_Null", "Dev_MARGE", "Dev_Null"):
            assert rel(rec[k][0], gold[k], 1e-6) < 1e-6, (g, k)
        assert rec["Degrees_Freedom"] == [gold["Degrees_Freedom"]] and rel(rec["P_Val"][0], gold["P_Val"], 1e-300) < 1e-5
        assert rec["Test_Stat_Note"] == [gold["Test_Stat_Note"]] and rec["MARGE_Fit_Notes"] == [gold["MARGE_Fit_Notes"]]
        assert rel(rec["Null_Summary"]["estimate"][0], gold["Null_Summary"]["estimate"][0]) < 1e-6
        sd, sg = rec["MARGE_Slope_Data"], gold["MARGE_Slope_Data"]
        assert [k for k in sd if k != "_class"] == list(sg.keys())
        for k in ("Gene", "Lineage", "Direction", "Notes"):
            assert sd[k] == sg[k], (g, k)
        for k in ("Breakpoint", "Rounded_Breakpoint", "Beta", "Test_Stat", "P_Val"):
            for a, b in zip(sd[k], sg[k]):
                assert (a is None and b is None) or rel(a, b, 1e-300) < 1e-5, (g, k, a, b)
        gd, gg = rec["Gene_Dynamics"], gold["Gene_Dynamics"]
        assert [k for k in gd if k != "_class"] == list(gg.keys()), (g, list(gd), list(gg))
        for k in gg:
            a, b = gd[k][0], gg[k][0]
            assert a == b if (isinstance(b, str) or b is None) else abs(a - b) <= 1e-6 * (abs(b) + 1e-9), (g, k, a, b)
        n += 1
    assert n == 89
