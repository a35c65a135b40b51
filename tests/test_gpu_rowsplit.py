"""One object over several GPUs by splitting the rows of its pair matrix (cppf_b200/rowsplit.py; SURVEY.md section
8e, second axis).  The exchange is three sums; the first is over the exact fixed-point vote grids, so the split
must reproduce the single-GPU grid, argmax and back-vote mask bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cppf_b200 import model, rowsplit, synth
from cppf_b200.pipeline import PoseConfig, PoseEstimator

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _estimator(dev=DEV, regress_right=False):
    torch.manual_seed(0)
    pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(dev).eval()
    ppf = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).to(dev).eval()
    cfg = PoseConfig.from_dict(dict(synth.BOTTLE, n_pairs=0, rot_subsample=0, regress_right=regress_right))
    return PoseEstimator(pe, ppf, cfg, dev)


def _inputs(n, dev=DEV):
    pc, nrm = synth.synth_bottle(n, 21)
    u = torch.rand(n * n, 4, generator=torch.Generator().manual_seed(6)).to(dev)
    inj = synth.trained_like_bins_dense_torch(torch.from_numpy(pc).to(dev), synth.BOTTLE)
    return pc, nrm, u, inj


@pytest.mark.parametrize("world,regress_right", [(2, False), (3, True)])
def test_row_split_in_lockstep_equals_the_whole_pair_list(world, regress_right):
    n = 333                                            # rows do not divide evenly
    est = _estimator(regress_right=regress_right)
    pc, nrm, u, inj = _inputs(n)
    # the whole pair matrix in ONE dense launch vs its row blocks (cppf_*_rows: the same row-aligned tiles, in-kernel)
    whole = est.estimate_fused(pc, nrm, seed=0, idxs=None, uniforms=u, inject_bins=inj, return_debug=True)
    split = rowsplit.estimate_rowsplit_local(est, pc, nrm, world, seed=0, uniforms=u, inject_bins=inj, return_debug=True)
    assert torch.equal(split["bins"], whole["bins"])
    assert torch.equal(split["grid"], whole["grid"])                    # exact integer sums, one rounding
    assert split["argmax_flat"] == whole["argmax_flat"]
    assert torch.equal(split["mask"], whole["mask"])
    assert split["n_survivors"] == whole["n_survivors"] > 0
    assert split["best_bins"] == whole["best_bins"]                     # no sub-sampling: integer histograms add up
    np.testing.assert_allclose(split["T_host"], whole["T_host"], rtol=0, atol=0)
    np.testing.assert_allclose(split["RT"], whole["RT"], rtol=1e-6, atol=1e-9)
    one = rowsplit.estimate_rowsplit(est, pc, nrm, seed=0, uniforms=u, inject_bins=inj)      # no process group: world 1
    assert one["argmax_flat"] == whole["argmax_flat"] and one["n_survivors"] == whole["n_survivors"]
    # no injected uniforms: the Philox stream is keyed by the pair's index in the WHOLE matrix, so the row blocks of any
    # world size draw exactly the bins of the single dense launch
    whole_s = est.estimate_fused(pc, nrm, seed=77, idxs=None, return_debug=True)
    split_s = rowsplit.estimate_rowsplit_local(est, pc, nrm, world, seed=77, return_debug=True)
    assert torch.equal(split_s["bins"], whole_s["bins"]) and torch.equal(split_s["grid"], whole_s["grid"])
    assert split_s["argmax_flat"] == whole_s["argmax_flat"] and torch.equal(split_s["mask"], whole_s["mask"])


def test_row_split_falls_back_to_a_pair_list_for_routed_grids():
    """A 64^3 grid does not fit the shared-memory vote kernel: the row blocks then run over an explicit pair list and the
    routed-slab vote, and still reproduce the single-GPU grid bit for bit."""
    n = 200
    est = _estimator()
    pc, nrm = synth.synth_cylinder_grid64(n, 3)
    u = torch.rand(n * n, 4, generator=torch.Generator().manual_seed(8)).to(DEV)
    whole = est.estimate_fused(pc, nrm, seed=0, idxs=rowsplit.block_pairs(n, 0, n, DEV), uniforms=u, return_debug=True)
    split = rowsplit.estimate_rowsplit_local(est, pc, nrm, 2, seed=0, uniforms=u, return_debug=True)
    assert torch.equal(split["grid"], whole["grid"]) and split["argmax_flat"] == whole["argmax_flat"]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        dev = f"cuda:{rank}"
        est = _estimator(dev)
        pc, nrm, u, inj = _inputs(n, dev)
        out = rowsplit.estimate_rowsplit(est, pc, nrm, seed=0, uniforms=u, inject_bins=inj, return_debug=True)
        q.put((rank, out["argmax_flat"], out["n_survivors"], out["best_bins"], out["RT"], out["grid"].cpu().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_row_split_over_two_gpus_nccl():
    world, n = 2, 333
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    est = _estimator()
    pc, nrm, u, inj = _inputs(n)
    whole = est.estimate_fused(pc, nrm, seed=0, idxs=None, uniforms=u, inject_bins=inj, return_debug=True)
    for _, flat, surv, bests, RT, grid in results:
        assert flat == whole["argmax_flat"] and surv == whole["n_survivors"] and bests == whole["best_bins"]
        np.testing.assert_array_equal(grid, whole["grid"].cpu().numpy())
        np.testing.assert_allclose(RT, whole["RT"], rtol=1e-6, atol=1e-9)
