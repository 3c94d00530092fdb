"""Gene sharding across ranks / GPUs: contiguous blocks of the gene-major count matrix, no collective
on the data path; records are gathered to rank 0 in gene order (the reference's `.inorder = TRUE`, testDynamic.R:177)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def gene_shard(rank: int, world: int, n_genes: int) -> tuple[int, int]:
    """Genes [g0, g1) of `rank` -- the same split sclane_test_dynamic() uses for its per-GPU threads."""
    return n_genes * rank // world, n_genes * (rank + 1) // world


def gather_records(payload: torch.Tensor, rank: int, world: int, n_genes: int, record_size: int) -> torch.Tensor:
    """All ranks pass their shard's records as a uint8 tensor; returns the concatenation in gene order (on every rank)."""
    if world == 1:
        return payload
    sizes = [(gene_shard(r, world, n_genes)[1] - gene_shard(r, world, n_genes)[0]) * record_size for r in range(world)]
    mx = max(sizes)
    buf = torch.zeros(mx, dtype=torch.uint8, device=payload.device)
    buf[: payload.numel()] = payload
    out = [torch.zeros(mx, dtype=torch.uint8, device=payload.device) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[:s] for o, s in zip(out, sizes)])
