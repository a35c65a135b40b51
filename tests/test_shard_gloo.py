"""Multi-rank host logic on CPU: object assignment and the single pose-record gather
(SURVEY.md section 8e), world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cppf_b200 import shard


def test_assign_objects_partitions_and_balances():
    costs = [4096 ** 2, 1024 ** 2, 2048 ** 2, 8192 ** 2, 512 ** 2, 4096 ** 2, 100.0]
    for mode in ("round_robin", "greedy"):
        for world in (1, 2, 3, 8):
            parts = shard.assign_objects(costs, world, mode)
            assert len(parts) == world
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(len(costs)))                      # a partition: every object exactly once
    g = shard.assign_objects(costs, 2, "greedy")
    load = [sum(costs[i] for i in p) for p in g]
    rr = shard.assign_objects(costs, 2, "round_robin")
    load_rr = [sum(costs[i] for i in p) for p in rr]
    assert max(load) <= max(load_rr)                                    # LPT never worse than dealing in order here
    assert shard.assign_objects([], 4) == [[], [], [], []]
    assert shard.assign_objects(costs, 2, "greedy") == g                 # deterministic
    with pytest.raises(ValueError):
        shard.assign_objects(costs, 2, "nope")


def _fake_record(i):
    r = np.zeros(shard.RECORD_FLOATS, np.float32)
    r[0] = i % 6
    r[1] = 1000 + i
    r[2:5] = (0.1 * i, 0.2 * i, 0.3 * i)
    r[5:14] = np.eye(3, dtype=np.float32).reshape(-1) * (i + 1)
    r[14:17] = (i, -i, 0.5 * i)
    return r


def _worker(rank, world, port, n_obj, mode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        costs = [(i % 5 + 1) ** 2 for i in range(n_obj)]
        called = []

        def est(i):
            called.append(i)
            return _fake_record(i)

        out = shard.estimate_sharded(est, costs, mode=mode)
        q.put((rank, called, out))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n_obj,mode", [(7, "greedy"), (1, "round_robin"), (6, "round_robin")])
def test_sharded_gather_world2_gloo(n_obj, mode):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_obj, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.stack([_fake_record(i) for i in range(n_obj)])
    all_called = sorted(i for _, called, _ in results for i in called)
    assert all_called == list(range(n_obj))                              # each object ran on exactly one rank
    for _, _, out in results:
        np.testing.assert_array_equal(out, want)                         # every rank holds the full, ordered result


def test_single_process_path_needs_no_process_group():
    out = shard.estimate_sharded(_fake_record, [1.0, 2.0, 3.0])
    np.testing.assert_array_equal(out, np.stack([_fake_record(i) for i in range(3)]))


def test_assignment_properties_hypothesis():
    """Any cost list, any world size: a partition; greedy never exceeds the list-scheduling bound
    mean load + largest object; round robin deals object i to rank i % world."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.floats(min_value=0.0, max_value=1e9, allow_nan=False), max_size=40), st.integers(1, 9))
    def check(costs, world):
        for mode in ("greedy", "round_robin"):
            parts = shard.assign_objects(costs, world, mode)
            assert sorted(i for p in parts for i in p) == list(range(len(costs)))
            assert all(p == sorted(p) for p in parts)
        g = shard.assign_objects(costs, world, "greedy")
        if costs:
            load = max(sum(costs[i] for i in p) for p in g)
            assert load <= sum(costs) / world + max(costs) + 1e-6 * (1 + sum(costs))
        rr = shard.assign_objects(costs, world, "round_robin")
        assert all(i % world == r for r, p in enumerate(rr) for i in p)

    check()
