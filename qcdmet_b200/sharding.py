"""Column sharding of the Jacobian over ranks (one process per GPU).

Perturbations are independent (lib/rhf_response.cc:78-79: an OpenMP `for` with no shared writes), so
rank g computes the contiguous column range shard(g) and the (Q x P/G) slices -- contiguous in the
column-major (Q, P) Jacobian -- are all-gathered.  Ranges are aligned to `align` columns (the
parameters of one fragment, 10 for 4-site impurities) so that no fragment is split between ranks.
"""
from __future__ import annotations


def column_shards(nterms: int, world: int, align: int = 1):
    """[(begin, end)] per rank; equal multiples of `align`, the tail ranks possibly shorter/empty."""
    assert nterms >= 0 and world >= 1 and align >= 1
    units = -(-nterms // align)                 # ceil
    per = -(-units // world) * align
    return [(min(r * per, nterms), min((r + 1) * per, nterms)) for r in range(world)]


def allgather_columns(local, nresid: int, shards, group=None):
    """All-gather column slices.  `local`: 1-D torch tensor holding this rank's (Q x ncols) slice in
    column-major order, padded to the widest shard.  Returns the full (Q*P,) column-major Jacobian."""
    import torch
    import torch.distributed as dist
    world = len(shards)
    width = max(e - b for b, e in shards)
    assert local.numel() == nresid * width
    gathered = torch.empty(world * nresid * width, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, local, group=group)
    if all(e - b == width for b, e in shards):
        return gathered
    parts = [gathered[r * nresid * width: r * nresid * width + nresid * (e - b)] for r, (b, e) in enumerate(shards)]
    return torch.cat(parts)
