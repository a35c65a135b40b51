"""One object's pair rows split over ranks (second axis): the partition helpers and the
claim the exchange rests on -- partial vote grids of row blocks sum to the full grid -- world_size 2 over gloo,
with the CPU oracle standing in for the vote kernel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cppf_b200 import rowsplit, synth
from oracle import clib


def test_row_blocks_partition_the_rows():
    for n in (1, 7, 96, 4096, 16384):
        for world in (1, 2, 3, 8):
            blocks = [rowsplit.row_block(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        rowsplit.row_block(10, 2, 2)


def test_block_pairs_are_slices_of_the_dense_enumeration():
    n = 37
    dense = synth.dense_pairs(n)
    for world in (1, 2, 5):
        got = [rowsplit.block_pairs(n, *rowsplit.row_block(n, world, r)).numpy() for r in range(world)]
        np.testing.assert_array_equal(np.concatenate(got), dense)
    lo, hi = rowsplit.row_block(n, 5, 3)
    np.testing.assert_array_equal(rowsplit.block_pairs(n, lo, hi).numpy(), dense[lo * n:hi * n])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = synth.BOTTLE["res"]
        pc, _ = synth.synth_bottle(n, 5)
        corner, dims = synth.vote_grid_geometry(pc, res)
        lo, hi = rowsplit.row_block(n, world, rank)
        idxs = rowsplit.block_pairs(n, lo, hi).numpy()
        mu_nu = synth.trained_like_tr(pc, idxs)
        part = clib.ppf_voting(pc, mu_nu, np.ones(n, np.float32), idxs, dims, corner, res, f64=True)
        t = torch.from_numpy(part)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)                   # exchange 1 of rowsplit._steps
        q.put((rank, (lo, hi), t.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_row_split_vote_grids_sum_to_the_full_grid_world2_gloo():
    world, n = 2, 96
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = synth.BOTTLE["res"]
    pc, _ = synth.synth_bottle(n, 5)
    corner, dims = synth.vote_grid_geometry(pc, res)
    idxs = synth.dense_pairs(n)
    full = clib.ppf_voting(pc, synth.trained_like_tr(pc, idxs), np.ones(n, np.float32), idxs, dims, corner, res, f64=True)
    assert full.max() > 10
    for _, _, grid in results:
        np.testing.assert_allclose(grid, full, rtol=1e-12, atol=1e-9)
        assert int(grid.argmax()) == int(full.argmax())
    np.testing.assert_array_equal(results[0][2], results[1][2])     # every rank holds the same reduced grid
