"""Export the reference's bundled datasets + golden results to portable fixtures.

Run HERE (the container with /root/reference mounted); the GPU box has no
reference checkout, so tests read only the files this script writes:

  tests/golden/sim_inputs.npz      counts (100 genes x 1198 cells, int32, gene-major),
                                   pseudotime (1198,), subject ids (1198, int32),
                                   gene names
  tests/golden/sclane_models.json  the 21-field testDynamic() record of every gene in
                                   data/scLANE_models.rda except the two per-cell
                                   data.frames (summarised, see below)
  tests/golden/sclane_preds.npz    MARGE_Preds / Null_Preds (link fit + se per cell)
                                   for a subset of genes, float64

Sources (all read-only, nothing is copied verbatim -- the .rda payloads are decoded
and re-encoded):
  /root/reference/data/sim_counts.rda       (R/data.R:1-8)
  /root/reference/data/sim_pseudotime.rda   (R/data.R:10-19)
  /root/reference/data/scLANE_models.rda    (R/data.R:21-27: "results from running
                                             testDynamic on sim_counts")
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from rdx3 import read_rda  # noqa: E402

REF = os.environ.get("SCLANE_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# genes whose per-cell predictions are kept (spread over model sizes)
PRED_GENES_N = 12


def _col(v):
    if v is None:
        return None
    x = v.value
    if isinstance(x, np.ndarray):
        if v.type == "integer" or v.type == "logical":
            return [None if int(a) == -2147483648 else int(a) for a in x]
        return [None if np.isnan(a) else float(a) for a in x]
    return list(x)


def _df(v):
    if v is None:
        return None
    return {nm: _col(c) for nm, c in zip(v.names(), v.value)}


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    # ---- inputs -------------------------------------------------------------
    sce = read_rda(os.path.join(REF, "data", "sim_counts.rda"))["sim_counts"]
    cm = sce["assays"]["data"]["listData"]["counts"]
    n_genes, n_cells = (int(a) for a in cm["Dim"].value)
    ii, pp, xx = cm["i"].value, cm["p"].value, cm["x"].value
    counts = np.zeros((n_genes, n_cells), dtype=np.int32)  # gene-major
    for c in range(n_cells):
        sl = slice(pp[c], pp[c + 1])
        counts[ii[sl], c] = np.round(xx[sl]).astype(np.int32)
    assert np.all(xx == np.round(xx))
    genes = list(cm["Dimnames"].value[0].value)
    cells = list(cm["Dimnames"].value[1].value)
    col = sce["colData"]["listData"]
    subj = list(col["subject"].value)
    levels = sorted(set(subj))
    subject = np.array([levels.index(s) for s in subj], dtype=np.int32)
    cell_time = col["cell_time_normed"].value.astype(np.float64)
    pt = read_rda(os.path.join(REF, "data", "sim_pseudotime.rda"))["sim_pseudotime"]
    pt = pt.value[0].value.astype(np.float64)
    assert pt.shape == (n_cells,)
    np.savez_compressed(os.path.join(OUT, "sim_inputs.npz"), counts=counts, pt=pt,
                        subject=subject, cell_time_normed=cell_time,
                        genes=np.array(genes), cells=np.array(cells),
                        subject_levels=np.array(levels))
    # ---- golden testDynamic output -----------------------------------------
    mods = read_rda(os.path.join(REF, "data", "scLANE_models.rda"))["scLANE_models"]
    assert list(mods.attr["class"].value) == ["scLANE"]
    out = {}
    preds = {}
    sizes = {}
    for g, gobj in zip(mods.names(), mods.value):
        lin = gobj.value[0]
        assert gobj.names() == ["Lineage_A"]
        rec = {}
        for nm, v in zip(lin.names(), lin.value):
            if nm in ("MARGE_Preds", "Null_Preds"):
                d = {k: c.value for k, c in zip(v.names(), v.value)}
                for k, arr in d.items():
                    preds[f"{g}/{k}"] = arr.astype(np.float64)
                # a summary that is cheap to compare for every gene
                rec[nm + "_digest"] = {k: [float(arr[0]), float(arr[len(arr) // 2]), float(arr[-1]),
                                           float(np.sum(arr)), float(np.sum(arr * arr))]
                                       for k, arr in d.items()}
            elif v is not None and v.type == "list":
                rec[nm] = _df(v)
            else:
                c = _col(v)
                rec[nm] = c[0] if c is not None and len(c) == 1 else c
        out[g] = rec
        sizes[g] = len(rec["MARGE_Summary"]["term"])
    with open(os.path.join(OUT, "sclane_models.json"), "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=False)
    # keep per-cell predictions for a few genes of every model size
    keep = []
    for sz in sorted(set(sizes.values())):
        gs = [g for g in out if sizes[g] == sz]
        keep += gs[: max(1, PRED_GENES_N // 4)]
    sub = {k: v for k, v in preds.items() if k.split("/")[0] in keep}
    np.savez_compressed(os.path.join(OUT, "sclane_preds.npz"), **sub)
    print(f"wrote fixtures for {len(out)} genes ({n_cells} cells) to {OUT}; preds for {keep}")


if __name__ == "__main__":
    main()
