"""world_size-2 gloo test (CPU) of the multi-rank host logic: column shards + all-gather reassemble
the full Jacobian.  The per-rank slice is produced by the oracle here (no GPU in this test)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import response_oracle as ro
from qcdmet_b200 import sharding, workloads as wl


def test_column_shards():
    assert sharding.column_shards(10000, 8, 10) == [(1250 * r, 1250 * (r + 1)) for r in range(8)]
    s = sharding.column_shards(101, 4, 10)
    assert s[0] == (0, 30) and s[-1] == (90, 101) and sum(e - b for b, e in s) == 101
    assert sharding.column_shards(5, 4, 10) == [(0, 5), (5, 5), (5, 5), (5, 5)]
    assert sharding.column_shards(0, 2) == [(0, 0), (0, 0)]
    for p, w, a in ((1440, 3, 10), (37, 5, 1), (60, 2, 10)):
        s = sharding.column_shards(p, w, a)
        assert s[0][0] == 0 and s[-1][1] == p and all(s[i][1] == s[i + 1][0] for i in range(w - 1))
        assert all(b % a == 0 for b, _ in s if b < p)


def _worker(rank, world, port, nterms, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = wl.sweep(24, nterms=nterms)
    e, c = ro.eigh_sorted(w.h0)
    shards = sharding.column_shards(w.nterms, world, 10)
    b, en = shards[rank]
    s = w.h1start
    jac = ro.jacobian_from_orbitals(c, e, w.num_pairs, s[b:en + 1] - s[b], w.h1row[s[b]:s[en]],
                                    w.h1col[s[b]:s[en]], w.row_start, w.rows, w.row_idx)
    width = max(x[1] - x[0] for x in shards)
    local = torch.zeros(w.nresid * width, dtype=torch.float64)
    local[:jac.size] = torch.from_numpy(jac.reshape(-1, order="F").copy())
    full = sharding.allgather_columns(local, w.nresid, shards)
    if rank == 0:
        np.save(out, full.numpy().reshape((w.nresid, w.nterms), order="F"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nterms", [60, 47])
def test_allgather_reassembles_jacobian(tmp_path, nterms):
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "jac.npy")
    mp.spawn(_worker, args=(2, port, nterms, out), nprocs=2, join=True)
    w = wl.sweep(24, nterms=nterms)
    e, c = ro.eigh_sorted(w.h0)
    ref = ro.jacobian_from_orbitals(c, e, w.num_pairs, w.h1start, w.h1row, w.h1col, w.row_start, w.rows, w.row_idx)
    assert np.array_equal(np.load(out), ref)
