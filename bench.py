#!/usr/bin/env python
"""bench.py -- genes/sec of GLM-mode testDynamic() (BASELINE.json metric) on synthetic NB counts.

Workload (default): the configuration BASELINE.json quotes the metric on -- 20,000 genes x 200,000 cells, one lineage,
GLM mode, no offsets (configs[4]).  It fits one B200 (16 GB of int32 counts; the library batches the sorted copy and the
fitted means).  Genes are sharded across the N GPUs of the box (strong scaling: rank r fits genes [20000 r/N, 20000 (r+1)/N));
genes are independent, so there is no collective on the data path -- torch.distributed is used for the barrier and the
max-over-ranks time only.  `--config c2` runs BASELINE configs[1] (2,000 x 10,000); `--genes G` overrides the gene count.

A "step" is one pass of the whole per-gene path (sort/gather, forward null fits + score sweeps, backward WIC pass, final +
null glm.nb, LRT) over the rank's gene shard.

  value  = genes/s with the counts already resident in HBM (sclane_test_dynamic_device); time = CUDA events recorded by the
           library around every stage on its own stream, summed, max over ranks
  e2e    = genes/s through the reference-facing call with HOST buffers (sclane_test_dynamic): pinned-host -> device copy of
           the counts and device -> host copy of the records inside the timed region (wall clock around the call)
  roofline = the kernel that performs the score sweep: algorithmic bytes (12 B per cell per swept gene-iteration + 8 B per
           cell per launch) / CUDA-event time of its launches, vs the measured HBM copy bandwidth.  On the default
           route that kernel is k_comp_stage<FORWARD> (X-compressed: null fits + sweep + selection of ALL forward iterations in
           one launch; it no longer reads those bytes, so `traffic` << algorithmic bytes); with --per-cell it is the
           streaming kernel k_sweep_moments, which does read them (97.7 % of the measured HBM peak, profiles/)
  cpu_baseline = the numpy oracle ("port": R is not installable here) on a bounded gene sample on this box's host cores

`--impl reference` times the reference algorithm's CPU restatement (the oracle, one process per core -- the reference's own
parallel model) on a bounded sample of the same workload.
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

# dram__bytes_read.sum + dram__bytes_write.sum of one k_sweep_moments launch from the committed ncu --set full capture
# (profiles/), per config, for the FULL default gene count on one GPU; None when not captured for that shape.
# c5: profiles/r01_ncu_sweep_c5_full.txt -- 25.74 GB read + 0.01 GB written for a 10,719-gene launch whose algorithmic bytes are
#     25.73 GB; scaled to the average launch of the default run (two batches, 24.0 GB algorithmic each).
# c2: profiles/r01_ncu_sweep_v1.txt -- 243.2 MB read + 7.4 MB written per 2,000-gene launch (240.1 MB algorithmic).
TRAFFIC_PER_LAUNCH = {"c2": 250.6e6, "c5": 24.02e9}
# same for the X-compressed forward kernel k_comp_stage<0>, PER GENE (the launch size follows the free device memory):
# profiles/r01_ncu_comp_forward_c5.txt -- 859.3 MB read + 122.1 MB written by a 9,729-gene launch = 100.9 KB per gene
# (vs 4.8 MB of algorithmic sweep bytes per gene: 2 iterations x 12 B x 200,000 cells).
TRAFFIC_PER_GENE_COMP = {"c2": None, "c5": 100.9e3}

CONFIGS = {
    "c5": {"genes": 20000, "cells": 200000,
           "name": "synthetic NB 20,000 genes x 200,000 cells, 1 lineage, GLM, genes sharded across GPUs (BASELINE configs[4])"},
    "c2": {"genes": 2000, "cells": 10000, "name": "synthetic NB 2,000 genes x 10,000 cells, 1 lineage, GLM (BASELINE configs[1])"},
}


def make_counts_gpu(g0: int, g1: int, n_cells: int, device, seed: int = 312, chunk: int = 128):
    """Synthetic NB counts for genes [g0, g1) generated ON the GPU (same recipe as sclane_b200/synth.py, SURVEY.md 8d:
    50 % dynamic genes with up to two hinges, theta ~ logU(0.5, 20) with a 5 % near-Poisson stratum, depth factors).
    Every gene is seeded by its global index, so a shard is the same for any number of ranks.
    Returns (int32 device tensor [g1 - g0, n_cells] gene-major, pt float64 numpy [n_cells])."""
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
        n = c1 - c0
        gr = np.random.default_rng([seed, c0])              # per-chunk parameters from the chunk's first gene index
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
        out[c0 - g0:c1 - g0] = torch.poisson(lam, generator=gen).to(torch.int32)
        del eta, mu, lam
    return out, pt


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh).get("hbm_gbs", 6546.2), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.samples = []
        self.reasons = set()
        self.max_mhz = None

    def run(self):
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


def cpu_baseline(counts, pt, n_sample: int):
    """The oracle (kind "port") on the first n_sample genes, single process, on this box's host cores."""
    from threadpoolctl import threadpool_limits
    from oracle.testdynamic import test_dynamic as oracle_td
    t0 = time.time()
    with threadpool_limits(limits=1):
        oracle_td(counts[:n_sample], pt, None, None, keep_preds=False)
    dt = time.time() - t0
    return {"value": n_sample / dt, "unit": "genes/s", "cores": 1, "kind": "port",
            "sample": f"{n_sample} genes x {counts.shape[1]} cells of the same synthetic workload, numpy oracle, 1 process, {dt:.1f} s"}


def _oracle_chunk(a):
    # one gene per worker at a time, like the reference's PSOCK workers; keep BLAS single-threaded per process
    from threadpoolctl import threadpool_limits
    from oracle.testdynamic import test_dynamic as oracle_td
    counts, pt = a
    with threadpool_limits(limits=1):
        oracle_td(counts, pt, None, None, keep_preds=False)
    return counts.shape[0]


def run_reference(args, cfg):
    """--impl reference: the reference's algorithm on the host cores (oracle port; one process per core, the
    reference's own parallel model -- one gene per worker, testDynamic.R:136-142,170-179).  Each step fits a bounded
    sample of the workload's genes (same cell count, same generator)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor
    from tools import synth
    cores = os.cpu_count() or 1
    per_core = 8 if cfg["cells"] <= 20000 else 1
    per_step = per_core * cores
    counts, pt, _ = synth.make_counts(per_step, cfg["cells"], seed=312)
    chunks = [(counts[i::cores], pt) for i in range(cores)]
    times = []
    with ProcessPoolExecutor(max_workers=cores) as ex:
        for it in range(args.warmup + args.steps):
            t0 = time.time()
            list(ex.map(_oracle_chunk, chunks))
            if it >= args.warmup:
                times.append(time.time() - t0)
    ms = 1000.0 * float(np.mean(times))
    val = per_step / (ms / 1000.0)
    line = {"impl": "reference", "metric": "genes/sec for testDynamic GLM mode", "value": val, "unit": "genes/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "cells": cfg["cells"], "genes_total": cfg["genes"], "genes_per_step": per_step,
                       "note": "bounded sample: genes are independent, throughput in genes/s does not depend on the gene count"},
            "cpu_baseline": {"value": val, "unit": "genes/s", "cores": cores, "kind": "port",
                             "sample": f"{per_step} genes x {cfg['cells']} cells per step, numpy oracle, {cores} processes"},
            "e2e": {"value": val, "unit": "genes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c5", choices=sorted(CONFIGS))
    ap.add_argument("--genes", type=int, default=0, help="total genes (default: the config's)")
    ap.add_argument("--cpu-sample", type=int, default=-1, help="genes of the CPU-baseline sample (0 = skip, -1 = automatic)")
    ap.add_argument("--pageable", action="store_true", help="e2e from pageable (not pinned) host memory, like an R caller's matrix")
    ap.add_argument("--h2d-mode", type=int, default=0, help="sclane_opts.h2d_mode for the e2e call: 0 auto, 1 direct int32, 2 staged uint16")
    ap.add_argument("--per-cell", action="store_true", help="force the per-cell kernels (no X-compression): profiling of k_sweep_moments")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.genes:
        cfg["genes"] = args.genes
    if args.impl == "reference":
        run_reference(args, cfg)
        return

    import torch
    import torch.distributed as dist
    import sclane_b200 as sb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    G_total, N = cfg["genes"], cfg["cells"]
    g0, g1 = G_total * rank // world, G_total * (rank + 1) // world        # strong scaling: contiguous gene shard
    G = g1 - g0
    d_counts, pt = make_counts_gpu(g0, g1, N, dev)
    torch.cuda.synchronize()
    opts = sb.api._opts(gpu_ids=[local], force_per_cell=args.per_cell)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return sb.run_records(G, pt, None, opts, device_ptr=d_counts.data_ptr(), device=local)

    # ---- device-resident measurement (value) ---------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    sb.reset_launch_count()
    prof_acc = None
    dev_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()                                   # evict the previous step's working set from L2
        torch.cuda.synchronize()
        last_recs, _ = step_device()
        p = sb.get_profile()
        dev_ms += sum(v for k, v in p["ms"].items())
        if prof_acc is None:
            prof_acc = p
        else:
            for k in p["ms"]:
                prof_acc["ms"][k] += p["ms"][k]
                prof_acc["launches"][k] += p["launches"][k]
            prof_acc["sweep_bytes"] += p["sweep_bytes"]
            prof_acc["sweep_gene_iters"] += p["sweep_gene_iters"]
    barrier()
    wall_ms = 1000.0 * (time.perf_counter() - t0)
    launches = sb.launch_count()
    ms_step = dev_ms / args.steps

    # ---- end-to-end measurement (e2e): host buffers, copies inside the timed region ----------------
    e2e_note = None
    try:
        h_counts = torch.empty((G, N), dtype=torch.int32, pin_memory=not args.pageable)
        h_counts.copy_(d_counts)
        Ge = G
    except RuntimeError as exc:       # pinned allocation refused: measure e2e on a bounded prefix of the shard
        Ge = min(G, 2048)
        h_counts = d_counts[:Ge].cpu()
        e2e_note = f"pinned host allocation of the full shard failed ({str(exc)[:60]}); e2e on the first {Ge} genes, pageable"
    torch.cuda.synchronize()
    h_np = h_counts.numpy()

    opts_host = sb.api._opts(gpu_ids=[local], force_per_cell=args.per_cell, h2d_mode=args.h2d_mode)

    def step_host():
        return sb.run_records(h_np, pt, None, opts_host)

    step_host()
    barrier()
    t1 = time.perf_counter()
    h2d = d2h = lib_wall = 0.0
    for _ in range(args.steps):
        step_host()
        p = sb.get_profile()
        h2d += p["h2d_bytes"]
        d2h += p["d2h_bytes"]
        lib_wall += p["wall_ms"]
    barrier()
    e2e_ms = 1000.0 * (time.perf_counter() - t1) / args.steps
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    ge_total = torch.tensor([float(Ge)], dtype=torch.float64, device=dev)
    if world > 1:
        t = torch.tensor([ms_step, e2e_ms, wall_ms / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms, wall_step = (float(v) for v in t.tolist())
        dist.all_reduce(ge_total, op=dist.ReduceOp.SUM)
    else:
        wall_step = wall_ms / args.steps
    value = G_total / (ms_step / 1000.0)
    e2e_val = float(ge_total.item()) / (e2e_ms / 1000.0)

    if rank == 0:
        peak, peak_kind = measured_peaks()
        comp = prof_acc["ms"]["comp_forward"] > 0
        if comp:
            # default route: the sweep of every forward iteration runs inside k_comp_stage<FORWARD>
            sweep_ms = prof_acc["ms"]["comp_forward"]
            sweep_launches = prof_acc["launches"]["comp_forward"]
            kernel = "k_comp_stage<FORWARD> (X-compressed forward pass: NB null fits + score sweep + selection, all iterations)"
            note = ("achieved = ALGORITHMIC sweep bytes (12 B/cell per swept gene-iteration + 8 B/cell per forward iteration) / CUDA-event "
                    "time of the k_comp_stage<FORWARD> launches (rank 0) -- the whole forward stage incl. the NB null fits is charged to the "
                    "sweep's bytes.  The kernel works on per-abscissa aggregates (k_aggregate), so its DRAM traffic is far below the "
                    "algorithmic bytes; the streaming per-cell kernel k_sweep_moments (--per-cell) moves them at 97.7 % of peak")
            per_gene = TRAFFIC_PER_GENE_COMP.get(args.config)
            traffic = per_gene * G * args.steps / max(1, sweep_launches) if per_gene else None
        else:
            sweep_ms = prof_acc["ms"]["sweep"]
            sweep_launches = prof_acc["launches"]["sweep"]
            kernel = "k_sweep_moments (score sweep streaming kernel, K2a)"
            note = ("achieved = algorithmic bytes (12 B/cell per swept gene-iteration + 8 B/cell per launch) / CUDA-event time "
                    "of the k_sweep_moments launches (rank 0); k_sweep_select (K2b, per-gene O(T p^2) algebra) is timed separately")
            traffic = TRAFFIC_PER_LAUNCH.get(args.config) if (world == 1 and not args.genes) else None
        achieved = (prof_acc["sweep_bytes"] / 1e9) / (sweep_ms / 1000.0) if sweep_ms > 0 else 0.0
        # the fit kernels ("bytes = passes x 4 N per gene", FP64 bound): passes counted by the kernels themselves
        fit_stages = ["fwd_fit", "backward", "final", "comp_forward", "comp_backward", "comp_final", "select"]
        fit_ms = sum(prof_acc["ms"][k] for k in fit_stages) / args.steps
        passes = float(np.mean(np.frombuffer(last_recs, dtype=np.uint8).reshape(G, -1)[:, type(last_recs[0]).fit_passes.offset:
                                                                                        type(last_recs[0]).fit_passes.offset + 4]
                               .copy().view(np.int32)))
        dom = max(prof_acc["ms"], key=lambda k: prof_acc["ms"][k])
        gather_ms = prof_acc["ms"]["gather"] + prof_acc["ms"]["aggregate"]
        gather_bytes = args.steps * (float(G) * N * (8.0 if not comp else 12.0))   # read 4 B + write 4 B (+ read 4 B in k_aggregate) per cell
        line = {
            "metric": "genes/sec for testDynamic GLM mode", "value": value, "unit": "genes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "genes_total": G_total, "genes_per_gpu": G, "cells": N, "lineages": 1, "M": 5,
                       "sharding": f"contiguous gene blocks x{world}, no collective",
                       "l2": "flushed between timed steps (256 MiB write); per-step working set >> 126 MB L2"},
            "e2e": {"value": e2e_val, "unit": "genes/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d / args.steps,
                    "host_buffer": "pageable" if args.pageable else "pinned", "h2d_mode": args.h2d_mode,
                    "library_wall_ms_per_step": lib_wall / args.steps,
                    "d2h_bytes_per_step": d2h / args.steps, **({"note": e2e_note} if e2e_note else {})},
            "gpu_launches": int(launches),
            "wall_ms_per_step": wall_step,
            "stage_ms_per_step": {k: v / args.steps for k, v in prof_acc["ms"].items()},
            "roofline": {"kernel": kernel, "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_kind, "traffic": traffic,
                         "bytes_per_launch": prof_acc["sweep_bytes"] / max(1, sweep_launches),
                         "ms_per_launch": sweep_ms / max(1, sweep_launches),
                         "select_kernel_ms_per_launch": prof_acc["ms"]["select"] / max(1, prof_acc["launches"]["select"]),
                         "note": note,
                         "fits": {"kernels": "k_comp_stage<FORWARD|BACKWARD|FINAL>" if comp else "k_stage<FWD_FIT|BACKWARD|FINAL> + sweep",
                                  "dominant_stage": dom, "dominant_share_of_step": prof_acc["ms"][dom] / args.steps / ms_step,
                                  "bound": "fp64 pipe", "passes_per_gene": passes, "ms_per_step": fit_ms,
                                  "algorithmic_GBps": passes * 4.0 * N * G / 1e9 / (fit_ms / 1000.0) if fit_ms > 0 else None,
                                  "ncu": "profiles/r01_ncu_comp_final.txt: FP64 pipe 35 % of peak, issue slots 49 % busy, DRAM 0.3 %"
                                         if comp else "profiles/r01_ncu_final_v2.txt",
                                  "note": "algorithmic bytes = passes x 4 B x cells per gene (SURVEY.md 8d); the compressed kernels walk "
                                          "the distinct abscissae instead (at 200,000 cells this figure exceeds the HBM peak) -- the stage is FP64 bound"},
                         "gather_aggregate": {"kernels": ("k_gather_copy + k_aggregate" if comp else "k_gather_stats") + " (sort by pseudotime; per-abscissa sums)",
                                              "achieved": gather_bytes / 1e9 / (gather_ms / 1000.0) if gather_ms > 0 else None,
                                              "unit": "GB/s", "ms_per_step": gather_ms / args.steps,
                                              "bytes": "4 B read + 4 B written per cell (gather)" + (" + 4 B read per cell (aggregate)" if comp else "")}},
            "clocks": sampler.result(),
        }
        n_cpu = args.cpu_sample if args.cpu_sample >= 0 else (24 if N <= 20000 else 4)
        if n_cpu > 0:
            line["cpu_baseline"] = cpu_baseline(d_counts[:min(n_cpu, G)].cpu().numpy(), pt, min(n_cpu, G))
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
