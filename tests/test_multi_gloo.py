"""N > 1 host logic on CPU: world_size-2 gloo group, each rank fits its gene shard (through the test-only host
emulation of the device algorithm, since there is no GPU here), results are gathered and must equal the one-rank run
bit for bit -- the property the multi-GPU path relies on (genes are independent; no data-path collective)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_genes, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C
    import __graft_entry__ as ge
    from sclane_b200 import _abi, shard
    from tools import synth
    counts, pt, _ = synth.make_counts(n_genes, 1500, seed=99)
    g0, g1 = shard.gene_shard(rank, world, n_genes)
    lib = C.CDLL(ge.EMU)
    mine = np.ascontiguousarray(counts[g0:g1], dtype=np.int32)
    ptc = np.ascontiguousarray(pt.T)
    o = _abi.default_opts()
    recs = (_abi.SclaneRecord * (g1 - g0))()
    infos = (_abi.SclaneLineageInfo * 1)()
    rc = lib.emu_test_dynamic(C.c_void_p(mine.ctypes.data), C.c_int64(1500), C.c_int64(g1 - g0), C.c_void_p(ptc.ctypes.data), 1,
                              C.c_void_p(None), C.byref(o), recs, infos)
    assert rc == 0
    payload = torch.frombuffer(bytearray(bytes(recs)), dtype=torch.uint8)
    gathered = shard.gather_records(payload, rank, world, n_genes, C.sizeof(_abi.SclaneRecord))
    # timing reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), gathered.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_gene_shards_cover_and_are_contiguous():
    sys.path.insert(0, ROOT)
    from sclane_b200 import shard
    for n in (1, 7, 100, 2000, 20000):
        for w in (1, 2, 4, 8):
            spans = [shard.gene_shard(r, w, n) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_matches_single_rank(tmp_path, emu):
    from sclane_b200 import _abi
    from tools import synth
    import ctypes as C
    n_genes = 10
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, n_genes, str(tmp_path)), nprocs=2, join=True)
    gathered = np.load(os.path.join(str(tmp_path), "gathered.npy"))
    counts, pt, _ = synth.make_counts(n_genes, 1500, seed=99)
    recs, _ = emu.run(counts, pt)
    single = np.frombuffer(bytes(recs), dtype=np.uint8)
    rs = C.sizeof(_abi.SclaneRecord)
    assert gathered.size == single.size == n_genes * rs
    keep = np.ones(rs, dtype=bool)                                                      # all but the timing fields
    for fld, n in ((_abi.SclaneRecord.gene_time, 8), (_abi.SclaneRecord.stage_cycles, 24)):
        keep[fld.offset:fld.offset + n] = False
    for g in range(n_genes):
        a, b = gathered[g * rs:(g + 1) * rs][keep], single[g * rs:(g + 1) * rs][keep]
        assert np.array_equal(a, b), g
