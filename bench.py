#!/usr/bin/env python
"""bench.py -- genes/sec of testDynamic() (BASELINE.json metric) on synthetic NB counts, one JSON line per run.

Workloads (`--config`; the default is the one BASELINE.json quotes the metric on):
  c5 (default)  20,000 genes x 200,000 cells, 1 lineage, GLM, no offsets          (BASELINE configs[4]); genes sharded over the ranks
  c2            2,000 x 10,000, 1 lineage, GLM                                     (configs[1])
  c4            10,000 x 50,000, 2 lineages (60 % of the cells each), GLM with createCellOffset() offsets   (configs[3])
  c3            5,000 x 30,000, 3 subjects, GEE mode, AR-1 working correlation, Wald test                    (configs[2])
Genes are independent, so there is no collective on the data path: rank r fits genes [G r/N, G (r+1)/N) (strong scaling);
torch.distributed is used for the barrier and the max-over-ranks time only.

A "step" = one pass of the whole per-gene path over the rank's gene shard.  Reported:
  value     genes/s with the counts already resident in HBM -- WALL CLOCK around the library call (host plan, launches,
            record copy and finalisation included), synchronised on both sides, max over ranks.  `stage_ms_per_step` holds the
            library's CUDA-event time per stage, `device_event_ms_per_step` their sum (what round 1 reported as value).
            (c3: the GEE entry point takes host counts; value = the same call from pinned memory, i.e. value == e2e.)
  e2e       the same through the reference-facing call with HOST buffers (pinned): H2D of the counts and D2H of the records
            inside the timed region; `e2e_pageable` = the same from pageable memory (what an R integer matrix is)
  roofline  the HBM-bound kernel of the path, k_sweep_moments (streaming score sweep, 12 B per cell per swept gene-iteration:
            int32 count + fp64 mu), measured LIVE in this run on a per-cell leg (`--sweep-genes` genes of the same
            workload, per-cell kernels forced): algorithmic bytes / CUDA-event time of its launches vs the measured HBM peak
  fits      the FP64-bound kernels that dominate the default route (X-compressed fits, offset quadrature, per-cell fits with an
            offset): their share of the step, FP64-pipe / issue utilisation from the committed ncu capture, and -- only as an
            "equivalent" figure, not an achieved bandwidth -- the 12 N sweep bytes they make unnecessary divided by their time
  cpu_baseline  the numpy oracle ("port": R cannot run here) on this box's host cores: all cores and one core, bounded samples

`--impl reference` times the CPU restatement of the reference algorithm (the oracle; one process per core, the reference's own
parallel model) on a bounded sample of the same workload.  It does not import the product package.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "c5": {"kind": "glm", "genes": 20000, "cells": 200000, "lineages": 1, "offsets": False,
           "name": "synthetic NB 20,000 genes x 200,000 cells, 1 lineage, GLM, genes sharded across GPUs (BASELINE configs[4])"},
    "c2": {"kind": "glm", "genes": 2000, "cells": 10000, "lineages": 1, "offsets": False,
           "name": "synthetic NB 2,000 genes x 10,000 cells, 1 lineage, GLM (BASELINE configs[1])"},
    "c4": {"kind": "glm", "genes": 10000, "cells": 50000, "lineages": 2, "offsets": True,
           "name": "synthetic NB 10,000 genes x 50,000 cells, 2 lineages (60 % of the cells each, 20 % shared), GLM with createCellOffset() offsets (BASELINE configs[3])"},
    "c3": {"kind": "gee", "genes": 5000, "cells": 30000, "lineages": 1, "offsets": True, "subjects": 3,
           "name": "synthetic NB 5,000 genes x 30,000 cells, 3 subjects, GEE mode, AR-1, Wald test, offsets (BASELINE configs[2])"},
}

# ncu --set full captures (profiles/): DRAM bytes of one k_sweep_moments launch relative to its algorithmic bytes (12 B per
# cell per swept gene + 8 B per cell), and where the FP64 / issue utilisation of the fit kernels is recorded.
SWEEP_TRAFFIC_RATIO = {"c5": 1.0005, "c2": 1.044}            # profiles/r01_ncu_sweep_c5_full.txt, r01_ncu_sweep_v1.txt
FIT_NCU = {
    "comp_forward": {"kernel": "k_comp_stage<FORWARD> (X-compressed: NB null fits + score sweep + selection)", "source": "profiles/r02_ncu_final_c5.txt"},
    "comp_backward": {"kernel": "k_comp_stage<BACKWARD> (X-compressed WIC pass)", "source": "profiles/r02_ncu_final_c5.txt"},
    "comp_final": {"kernel": "k_comp_stage<FINAL> (X-compressed final + null glm.nb)", "source": "profiles/r02_ncu_final_c5.txt"},
    "final": {"kernel": "k_stage<FINAL, HYB> (hybrid per-cell glm.nb of the alternatives that carry an offset; theta.ml on nodes)", "source": "profiles/r02_ncu_final_c4.txt"},
    "off_null": {"kernel": "k_wg_stage<OFFNULL> (offset quadrature, warp per gene)", "source": "profiles/r02_ncu_final_c4.txt"},
    "gee": {"kernel": "k_gee", "source": "profiles/r01_ncu_gee_c3.txt"},
}


def _gene_block_gpu(c0, c1, x, d, seed, device, gen):
    """Counts of genes [c0, c1) (SURVEY.md 8d recipe): 50 % dynamic genes with up to two hinges, theta ~ logU(0.5, 20) with a
    5 % near-Poisson stratum, per-cell depth factors.  Parameters are seeded by the block's first gene index."""
    import torch
    n, n_cells = c1 - c0, x.numel()
    gr = np.random.default_rng([seed, c0])
    b0 = torch.from_numpy(gr.normal(0.0, 1.0, n)).to(device)
    dyn = torch.from_numpy((gr.random(n) < 0.5).astype(np.float64)).to(device)
    eta = b0[:, None].expand(n, n_cells).clone()
    for _ in range(2):
        tau = torch.from_numpy(gr.uniform(0.1, 0.9, n)).to(device)
        a = torch.from_numpy(gr.normal(0.0, 2.0, n)).to(device) * dyn
        right = torch.from_numpy((gr.random(n) < 0.5)).to(device)
        h = torch.where(right[:, None], (x[None, :] - tau[:, None]).clamp_min(0.0), (tau[:, None] - x[None, :]).clamp_min(0.0))
        eta += a[:, None] * h
    th = np.exp(gr.uniform(np.log(0.5), np.log(20.0), n))
    th[gr.random(n) < 0.05] = 1e3
    theta = torch.from_numpy(th).to(device)
    mu = torch.exp(eta.clamp(-8.0, 8.0)) * d[None, :]
    gen.manual_seed(seed * 1000003 + c0)
    lam = torch._standard_gamma(theta[:, None].expand(n, n_cells).contiguous(), generator=gen) * (mu / theta[:, None])
    return torch.poisson(lam, generator=gen).to(torch.int32)


def make_counts_gpu(g0: int, g1: int, n_cells: int, device, seed: int = 312, chunk: int = 128):
    """Synthetic NB counts for genes [g0, g1) generated ON the GPU; every gene block is seeded by its global index, so a shard
    is the same for any number of ranks.  Returns (int32 device tensor [g1 - g0, n_cells] gene-major, pt float64 numpy [n_cells])."""
    import torch
    rng = np.random.default_rng(seed)
    perm = rng.permutation(n_cells)
    pt = np.empty(n_cells)
    pt[perm] = (np.arange(n_cells) + 0.5) / n_cells
    depth = np.exp(rng.normal(0.0, 0.3, n_cells))
    x = torch.from_numpy(pt).to(device)
    d = torch.from_numpy(depth).to(device)
    out = torch.empty((g1 - g0, n_cells), dtype=torch.int32, device=device)
    gen = torch.Generator(device=device)
    for c0 in range(g0, g1, chunk):
        c1 = min(c0 + chunk, g1)
        out[c0 - g0:c1 - g0] = _gene_block_gpu(c0, c1, x, d, seed, device, gen)
    return out, pt


def two_lineages(pt: np.ndarray, seed: int = 312) -> np.ndarray:
    """C4 layout: v ~ U(0,1) per cell; lineage A = v < 0.6, lineage B = v > 0.4 (20 % shared); every lineage gets
    its own uniform pseudotime over its member cells (ranks of x, resp. of 1 - x); NaN = not in the lineage."""
    n = pt.size
    v = np.random.default_rng(seed).random(n)

    def rank01(x, mask):
        out = np.full(n, np.nan)
        idx = np.flatnonzero(mask)
        order = np.argsort(x[idx], kind="stable")
        r = np.empty(idx.size)
        r[order] = (np.arange(idx.size) + 0.5) / idx.size
        out[idx] = r
        return out

    return np.stack([rank01(pt, v < 0.6), rank01(1.0 - pt, v > 0.4)], 1)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh).get("hbm_gbs", 6546.2), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.samples = []
        self.reasons = set()
        self.max_mhz = None

    def _run_nvml(self) -> bool:
        """In-process NVML polling (no fork, no second NVML client): an nvidia-smi child every 200 ms costs the timed steps
        milliseconds of driver-lock and GIL stalls, which is a visible share of a 10 ms step."""
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        except Exception:
            return False
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for bit, name in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.05)
        return True

    def run(self):
        if self._run_nvml():
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(n)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def result(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle (test infrastructure) as the reported baseline / the --impl reference arm.  Nothing below imports
# the product package.
# ------------------------------------------------------------------------------------------------
def _oracle_inputs(cfg, n_genes: int, cells_cap=None):
    """Host-side sample of the workload for the oracle: (counts, pt, size_factor, subject)."""
    from tools import synth
    n_cells = cfg["cells"] if cells_cap is None else min(cfg["cells"], cells_cap)
    counts, pt, _ = synth.make_counts(n_genes, n_cells, seed=312)
    pt = pt.reshape(-1)
    sf = subject = None
    if cfg["kind"] == "gee":
        subject = (np.arange(n_cells) * cfg["subjects"] // n_cells).astype(np.int32)
        order = np.lexsort((pt, subject))
        counts, pt = np.ascontiguousarray(counts[:, order]), pt[order]
    if cfg["lineages"] == 2:
        pt = two_lineages(pt)
    if cfg["offsets"]:
        sf = 1e4 / np.maximum(counts.sum(axis=0), 1).astype(np.float64)
    return counts, pt, sf, subject


def _oracle_chunk(a):
    # one worker = one R PSOCK worker of the reference (testDynamic.R:136-142,170-179); BLAS single-threaded per process
    from threadpoolctl import threadpool_limits
    kind, counts, pt, sf, subject = a
    with threadpool_limits(limits=1):
        if kind == "gee":
            from oracle.gee import test_dynamic_gee
            test_dynamic_gee(counts, pt, subject, None, sf, "ar1")
        else:
            from oracle.testdynamic import test_dynamic
            test_dynamic(counts, pt, None, sf, keep_preds=False)
    return counts.shape[0]


def _oracle_rate(cfg, per_core: int, cores: int, pool=None, cells_cap=None):
    """(genes/s, seconds) of the oracle on per_core * cores genes spread over `cores` processes."""
    counts, pt, sf, subject = _oracle_inputs(cfg, per_core * cores, cells_cap)
    chunks = [(cfg["kind"], counts[i::cores], pt, sf, subject) for i in range(cores)]
    t0 = time.time()
    if cores == 1:
        _oracle_chunk(chunks[0])
    else:
        list(pool.map(_oracle_chunk, chunks))
    dt = time.time() - t0
    return per_core * cores / dt, dt


def _per_core_sample(cfg):
    """Genes per core for ~10-15 s of oracle time per step (measured rates of the numpy port with every core busy: ~0.07 genes/s/core
    at 200,000 cells, ~0.45 at 10,000; the dense GEE oracle is O(n_i^3) and only runs at a reduced subject size).
    Returns (genes per core, cells cap)."""
    if cfg["kind"] == "gee":
        return 1, 1200            # 3 subjects x 400 cells
    n = cfg["cells"] * (0.6 if cfg["lineages"] == 2 else 1.0) * cfg["lineages"]
    return max(1, min(16, int(round(4.0e4 / n)))), None


def cpu_baseline(cfg):
    from concurrent.futures import ProcessPoolExecutor
    cores = os.cpu_count() or 1
    per_core, cap = _per_core_sample(cfg)
    one, dt1 = _oracle_rate(cfg, per_core, 1, cells_cap=cap)
    with ProcessPoolExecutor(max_workers=cores) as ex:
        allc, dta = _oracle_rate(cfg, per_core, cores, ex, cells_cap=cap)
    out = {"value": allc, "unit": "genes/s", "cores": cores, "kind": "port",
           "sample": f"{per_core * cores} genes x {cap or cfg['cells']} cells of the same synthetic workload, numpy oracle, {cores} processes, {dta:.1f} s",
           "single_core": {"value": one, "unit": "genes/s", "cores": 1, "sample": f"{per_core} genes, 1 process, {dt1:.1f} s"},
           "note": "a reported baseline (CPU restatement of the reference algorithm; R is not installable here), not the optimisation target"}
    if cap:
        out["note"] += f"; the dense GEE oracle builds n_i x n_i matrices like the reference does and only runs at a reduced size ({cap} cells)"
    # context promised in BASELINE.md: the port on the reference's own bundled data next to the golden's Gene_Time
    try:
        from threadpoolctl import threadpool_limits
        from oracle.testdynamic import test_dynamic
        d = np.load(os.path.join(ROOT, "tests", "golden", "sim_inputs.npz"))
        counts = np.ascontiguousarray(d["counts"], dtype=np.int32)
        with open(os.path.join(ROOT, "tests", "golden", "sclane_models.json")) as fh:
            gold = json.load(fh)
        gt = [v["Gene_Time"] for v in gold.values() if isinstance(v.get("Gene_Time"), (int, float))]
        n_c1 = 12
        t0 = time.time()
        with threadpool_limits(limits=1):
            test_dynamic(counts[:n_c1], d["pt"], None, 1e4 / counts.sum(axis=0).astype(np.float64), keep_preds=False)
        out["c1_context"] = {"workload": "bundled sim_counts (1,198 cells, offsets; BASELINE configs[0])", "port_s_per_gene_1core": (time.time() - t0) / n_c1,
                             "reference_Gene_Time_s_per_gene": float(np.mean(gt)) if gt else None,
                             "reference_source": "data/scLANE_models.rda Gene_Time (the package authors' own R/Eigen run)"}
    except Exception as exc:      # the context line must never break the bench
        out["c1_context"] = {"error": str(exc)[:120]}
    return out


def run_reference(args, cfg):
    """--impl reference: the reference's algorithm on the host cores (oracle port; one process per core).  Each step fits a
    bounded sample of the workload's genes (same cell count, same generator)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor
    cores = os.cpu_count() or 1
    per_core, cap = _per_core_sample(cfg)
    per_step = per_core * cores
    times = []
    with ProcessPoolExecutor(max_workers=cores) as ex:
        for it in range(args.warmup + args.steps):
            _, dt = _oracle_rate(cfg, per_core, cores, ex, cells_cap=cap)
            if it >= args.warmup:
                times.append(dt)
    ms = 1000.0 * float(np.mean(times))
    val = per_step / (ms / 1000.0)
    cells = cap or cfg["cells"]
    line = {"impl": "reference", "metric": "genes/sec for testDynamic %s mode" % ("GEE" if cfg["kind"] == "gee" else "GLM"), "value": val, "unit": "genes/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "cells": cells, "genes_total": cfg["genes"], "genes_per_step": per_step,
                       "note": "bounded sample: genes are independent, throughput in genes/s does not depend on the gene count"
                               + ("; dense GEE oracle at a reduced subject size (the reference's n_i x n_i formulation does not run at 10,000 cells per subject)" if cap else "")},
            "cpu_baseline": {"value": val, "unit": "genes/s", "cores": cores, "kind": "port",
                             "sample": f"{per_step} genes x {cells} cells per step, numpy oracle, {cores} processes"},
            "e2e": {"value": val, "unit": "genes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c5", choices=sorted(CONFIGS))
    ap.add_argument("--genes", type=int, default=0, help="total genes (default: the config's)")
    ap.add_argument("--cpu-sample", type=int, default=-1, help="0 = skip the CPU baseline")
    ap.add_argument("--sweep-genes", type=int, default=2048, help="genes of the per-cell leg that times k_sweep_moments (0 = skip)")
    ap.add_argument("--no-pageable", action="store_true", help="skip the pageable-memory end-to-end leg")
    ap.add_argument("--h2d-mode", type=int, default=0, help="sclane_opts.h2d_mode for the e2e call: 0 auto, 1 direct int32, 2 staged uint16")
    ap.add_argument("--per-cell", action="store_true", help="force the per-cell kernels for the whole run (profiling)")
    ap.add_argument("--warp-per-gene", action="store_true", help="X-compressed fits on the warp-per-gene kernels (A/B against the CTA-per-gene ones)")
    ap.add_argument("--offset-per-cell", action="store_true", help="intercept-only fits with an offset per cell (A/B against the offset quadrature)")
    ap.add_argument("--genes-per-batch", type=int, default=0, help="sclane_opts.genes_per_batch of the end-to-end legs (0 = the library's choice)")
    ap.add_argument("--batch-streams", type=int, default=0, help="sclane_opts.batch_streams of the timed legs: 0 = the library's default (2 gene batches in flight), 1 = sequential batches")
    ap.add_argument("--approx-knot", type=int, default=1, help="0 = approx.knot = FALSE: every abscissa kept by min_span / max_span is a candidate knot")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.genes:
        cfg["genes"] = args.genes
    if args.impl == "reference":
        run_reference(args, cfg)
        return

    # one rank per GPU shares the host with the other ranks: each takes its share of the cores for the library's host threads
    # (plan / record finalisation / narrowing of the counts for the transfer) -- read by the library when it is loaded
    _local_world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    if _local_world > 1:
        os.environ.setdefault("SCLANE_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // _local_world)))
    import torch
    import torch.distributed as dist
    import sclane_b200 as sb
    from sclane_b200 import _abi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    gee = cfg["kind"] == "gee"

    G_total, N = cfg["genes"], cfg["cells"]
    g0, g1 = G_total * rank // world, G_total * (rank + 1) // world        # strong scaling: contiguous gene shard
    G = g1 - g0
    d_counts, pt = make_counts_gpu(g0, g1, N, dev)
    subject = None
    if gee:
        subject = (np.arange(N) * cfg["subjects"] // N).astype(np.int32)
        order = np.lexsort((pt, subject))                                   # cells ordered by subject, then pseudotime
        d_counts = d_counts[:, torch.from_numpy(order).to(dev)].contiguous()
        pt = pt[order]
    pt_in = two_lineages(pt) if cfg["lineages"] == 2 else pt
    sf = None
    if cfg["offsets"]:
        # createCellOffset(): 1e4 / colSums over ALL genes -> sum the shards' column sums over the ranks (data preparation)
        cs = d_counts.sum(dim=0, dtype=torch.int64).to(torch.float64)
        if world > 1:
            dist.all_reduce(cs)
        sf = (1e4 / cs.clamp_min(1.0)).cpu().numpy()
    torch.cuda.synchronize()

    def mk_opts(**kw):
        kw.setdefault("force_per_cell", args.per_cell)
        o = sb.api._opts(gpu_ids=[local], comp_warp_per_gene=args.warp_per_gene, offset_per_cell=args.offset_per_cell,
                         approx_knot=bool(args.approx_knot), **kw)
        if gee:
            o.mode = 1
            o.gee_cor = _abi.COR_CODES["ar1"]
        return o

    opts = mk_opts(batch_streams=args.batch_streams)
    opts_seq = mk_opts(batch_streams=1)        # sequential batches: per-kernel CUDA-event times that do not overlap (stage breakdown, roofline legs)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def acc_profile(acc, p):
        if acc is None:
            return p
        for k in p["ms"]:
            acc["ms"][k] += p["ms"][k]
            acc["launches"][k] += p["launches"][k]
        for k in ("sweep_bytes", "sweep_gene_iters", "h2d_bytes", "d2h_bytes", "wall_ms"):
            acc[k] += p[k]
        return acc

    # ---- host copy of the shard (e2e legs; GEE mode has no device-resident entry point) ------------
    e2e_note = None
    try:
        h_counts = torch.empty((G, N), dtype=torch.int32, pin_memory=True)
        h_counts.copy_(d_counts)
        Ge = G
    except RuntimeError as exc:       # pinned allocation refused: measure e2e on a bounded prefix of the shard
        Ge = min(G, 2048)
        h_counts = d_counts[:Ge].cpu()
        e2e_note = f"pinned host allocation of the full shard failed ({str(exc)[:60]}); e2e on the first {Ge} genes, pageable"
    torch.cuda.synchronize()
    h_np = h_counts.numpy()
    opts_host = mk_opts(h2d_mode=args.h2d_mode, genes_per_batch=args.genes_per_batch, batch_streams=args.batch_streams)

    def step_device():
        return sb.run_records(G, pt_in, sf, opts, device_ptr=d_counts.data_ptr(), device=local)

    def step_host(buf=None):
        return sb.run_records(h_np if buf is None else buf, pt_in, sf, opts_host, subject_id=subject)

    # ---- device-resident measurement (value) ---------------------------------------------------
    step_value = step_host if gee else step_device
    for _ in range(args.warmup):
        step_value()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    sb.reset_launch_count()
    prof_acc = None
    wall_steps = []
    last_recs = None
    for _ in range(args.steps):
        flush.zero_()                                   # evict the previous step's working set from L2
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        last_recs, _ = step_value()
        torch.cuda.synchronize()
        wall_steps.append(1000.0 * (time.perf_counter() - t0))
        prof_acc = acc_profile(prof_acc, sb.get_profile())
    barrier()
    launches = sb.launch_count()
    ms_step = float(np.mean(wall_steps))
    wall_steps_value = list(wall_steps)
    # stage breakdown: the same step with strictly sequential batches, so that the CUDA-event intervals of the stages add up
    # (with two batches in flight the intervals of the two streams overlap)
    seq_wall = []
    prof_acc = None
    for it_seq in range(3):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if gee:
            sb.run_records(h_np, pt_in, sf, opts_seq, subject_id=subject)
        else:
            sb.run_records(G, pt_in, sf, opts_seq, device_ptr=d_counts.data_ptr(), device=local)
        torch.cuda.synchronize()
        if it_seq == 0:
            continue                                    # the batch geometry differs from the timed legs: buffers are re-sized once
        seq_wall.append(1000.0 * (time.perf_counter() - t0))
        prof_acc = acc_profile(prof_acc, sb.get_profile())
    n_prof_steps = 2
    event_ms = sum(prof_acc["ms"].values()) / n_prof_steps

    # ---- end-to-end measurement (e2e): host buffers, copies inside the timed region ----------------
    def time_host(buf, steps):
        step_host(buf)
        barrier()
        t1 = time.perf_counter()
        acc = None
        for _ in range(steps):
            step_host(buf)
            acc = acc_profile(acc, sb.get_profile())
        barrier()
        return 1000.0 * (time.perf_counter() - t1) / steps, acc

    e2e_ms, e2e_prof = time_host(None, args.steps)
    pageable_ms = None
    if not args.no_pageable:
        try:
            h_page = np.array(h_np[:Ge], copy=True)      # plain malloc'ed memory, like an R integer matrix
            pageable_ms, _ = time_host(h_page, max(1, min(2, args.steps)))
            del h_page
        except MemoryError:
            pageable_ms = None
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    # ---- per-cell leg: the streaming score-sweep kernel, live -------------------------------------------
    sweep = None
    if args.sweep_genes > 0 and not gee and rank == 0 and args.approx_knot:
        Gs = min(G, args.sweep_genes)
        o_pc = mk_opts(force_per_cell=True, batch_streams=1)
        for _ in range(2):
            sb.run_records(Gs, pt_in, None, o_pc, device_ptr=d_counts.data_ptr(), device=local)
        acc = None
        for _ in range(3):
            flush.zero_()
            torch.cuda.synchronize()
            sb.run_records(Gs, pt_in, None, o_pc, device_ptr=d_counts.data_ptr(), device=local)
            acc = acc_profile(acc, sb.get_profile())
        sweep = {"genes": Gs, "ms": acc["ms"]["sweep"], "launches": acc["launches"]["sweep"], "bytes": acc["sweep_bytes"],
                 "select_ms": acc["ms"]["select"], "select_launches": acc["launches"]["select"]}

    ge_total = torch.tensor([float(Ge)], dtype=torch.float64, device=dev)
    if world > 1:
        t = torch.tensor([ms_step, e2e_ms, event_ms, pageable_ms or 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms, event_ms, pg = (float(v) for v in t.tolist())
        pageable_ms = pg if pageable_ms is not None else None
        dist.all_reduce(ge_total, op=dist.ReduceOp.SUM)
    value = G_total / (ms_step / 1000.0)
    e2e_val = float(ge_total.item()) / (e2e_ms / 1000.0)

    if rank == 0:
        peak, peak_kind = measured_peaks()
        stage_ms = {k: v / n_prof_steps for k, v in prof_acc["ms"].items()}
        fit_stages = ["fwd_fit", "backward", "final", "comp_forward", "comp_backward", "comp_final", "select", "gee", "off_null"]
        dom = max(fit_stages, key=lambda k: stage_ms.get(k, 0.0))
        off = type(last_recs[0]).fit_passes.offset
        passes = float(np.mean(np.frombuffer(last_recs, dtype=np.uint8).reshape(len(last_recs), -1)[:, off:off + 4].copy().view(np.int32)))
        n_lin_cells = N * (0.6 if cfg["lineages"] == 2 else 1.0)
        if sweep and sweep["ms"] > 0:
            achieved = sweep["bytes"] / 1e9 / (sweep["ms"] / 1000.0)
            ratio = SWEEP_TRAFFIC_RATIO.get(args.config)
            roofline = {"kernel": "k_sweep_moments (streaming score sweep, K2a: int32 count + fp64 mu per cell -> per-bucket moments)",
                        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_kind,
                        "traffic": (ratio * sweep["bytes"] / sweep["launches"]) if ratio else None,
                        "bytes_per_launch": sweep["bytes"] / sweep["launches"], "ms_per_launch": sweep["ms"] / sweep["launches"],
                        "measured": f"live in this run: per-cell leg, {sweep['genes']} genes x {N} cells of the same workload, 3 steps, CUDA events on the library's stream",
                        "select_kernel_ms_per_launch": sweep["select_ms"] / max(1, sweep["select_launches"]),
                        "note": "algorithmic bytes = 12 B/cell per swept gene-iteration + 8 B/cell per launch (SURVEY.md 8d); traffic = ncu dram bytes of a "
                                "launch of this shape scaled by launch size (profiles/r01_ncu_sweep_*.txt).  The DEFAULT route does not run this kernel: "
                                "it sweeps the <= 10,001 distinct abscissae inside the FP64-bound fit kernels (see `fits`)"}
        else:
            roofline = {"kernel": FIT_NCU.get(dom, {}).get("kernel", dom), "bound": "fp64", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None,
                        "traffic": None, "note": "no HBM-bound kernel was timed on this path (GEE mode, or the per-cell leg was skipped): see `fits`"}
        sweep_equiv = prof_acc["sweep_bytes"] / 1e9 / (prof_acc["ms"]["comp_forward"] / 1000.0) if prof_acc["ms"].get("comp_forward", 0) > 0 else None
        fits = {"dominant_stage": dom, "dominant_kernel": FIT_NCU.get(dom, {}).get("kernel", dom), "bound": "serial algebra + barriers between short passes (bucket nodes); fp64 pipe for the per-cell / GEE kernels",
                "dominant_share_of_step": stage_ms[dom] / max(event_ms, 1e-9), "passes_per_gene_lineage": passes,
                "ncu": FIT_NCU.get(dom, {}).get("source"),
                "forward_stage_equivalent_GBps": sweep_equiv,
                "note": "equivalent_GBps = the per-cell sweep bytes (12 N per swept gene-iteration) that the X-compressed forward stage makes unnecessary / its "
                        "time; NOT an achieved bandwidth (the kernel reads ~2 % of those bytes)"}
        line = {
            "metric": "genes/sec for testDynamic %s mode" % ("GEE" if gee else "GLM"), "value": value, "unit": "genes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "genes_total": G_total, "genes_per_gpu": G, "cells": N, "lineages": cfg["lineages"], "M": 5,
                       "offsets": cfg["offsets"], "mode": cfg["kind"], "cells_per_lineage": int(n_lin_cells), "approx_knot": bool(args.approx_knot),
                       "sharding": f"contiguous gene blocks x{world}, no collective",
                       "timing": "value = wall clock around the library call with the counts resident in HBM (host plan, launches, D2H of the records, record "
                                 "finalisation included); max over ranks" + ("; the GEE entry takes host counts: value == the e2e path from pinned memory" if gee else "")
                                 + "; stage_ms_per_step / device_event_ms_per_step come from two extra steps with sequential batches (batch_streams = 1), whose wall clock is sequential_wall_ms_per_step",
                       "batch_streams": args.batch_streams if args.batch_streams else 2,
                       "l2": "flushed between timed steps (256 MiB write); per-step working set >> 126 MB L2"},
            "e2e": {"value": e2e_val, "unit": "genes/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": e2e_prof["h2d_bytes"] / args.steps,
                    "d2h_bytes_per_step": e2e_prof["d2h_bytes"] / args.steps, "host_buffer": "pinned", "h2d_mode": args.h2d_mode,
                    "library_wall_ms_per_step": e2e_prof["wall_ms"] / args.steps, **({"note": e2e_note} if e2e_note else {})},
            "e2e_pageable": ({"value": float(ge_total.item()) / (pageable_ms / 1000.0), "unit": "genes/s", "ms_per_step": pageable_ms,
                              "host_buffer": "pageable (what an R integer matrix is); staged uint16 transfer"} if pageable_ms else None),
            "gpu_launches": int(launches),
            "device_event_ms_per_step": event_ms,
            "sequential_wall_ms_per_step": float(np.mean(seq_wall)),
            "wall_ms_of_each_step": [round(v, 3) for v in wall_steps_value],
            "stage_ms_per_step": stage_ms,
            "roofline": roofline,
            "fits": fits,
            "clocks": sampler.result(),
        }
        if args.cpu_sample != 0:
            line["cpu_baseline"] = cpu_baseline(cfg)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
