"""Object sharding over ranks.

The reference handles objects one at a time in one process (``nocs/inference.py:120``,
``sunrgbd/inference.py:115``); objects are independent, so here they are dealt out to the ranks of
one ``torch.distributed`` job (one process per GPU, NCCL over NVLink on the box; ``gloo`` in the CPU
tests), each rank runs the whole per-object pipeline locally and ONE ``all_gather`` of fixed-size
pose records ends the batch.  Record layout = ``sunrgbd/inference.py:287``:
``[class_id, score, scale x3, R x9 (row-major), T x3]`` = 17 float32, prefixed here by the object's
global index so the gathered list can be put back in input order.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np
import torch
import torch.distributed as dist

RECORD_FLOATS = 17
_ROW = RECORD_FLOATS + 1            # + global object index


def assign_objects(costs: Sequence[float], world: int, mode: str = "greedy") -> List[List[int]]:
    """Deal objects to ranks.  costs[i] ~ work of object i (N_i^2 pairs for the dense path, P for the
    sampled one).  "round_robin": object i -> rank i % world.  "greedy": longest-processing-time
    first onto the least loaded rank (ties -> lowest rank), which is what matters when N varies by
    category.  Deterministic: every rank computes the same assignment without communicating."""
    n = len(costs)
    if world <= 0:
        raise ValueError("world must be positive")
    parts: List[List[int]] = [[] for _ in range(world)]
    if mode == "round_robin":
        for i in range(n):
            parts[i % world].append(i)
        return parts
    if mode != "greedy":
        raise ValueError(f"unknown assignment mode {mode!r}")
    load = [0.0] * world
    for i in sorted(range(n), key=lambda j: (-float(costs[j]), j)):
        r = min(range(world), key=lambda q: (load[q], q))
        parts[r].append(i)
        load[r] += float(costs[i])
    for p in parts:
        p.sort()
    return parts


def gather_records(local_ids: Sequence[int], local_records: np.ndarray, n_total: int, device=None, group=None) -> np.ndarray:
    """The one collective of the path: every rank contributes its [k_r, 17] records, every rank gets
    the full [n_total, 17] array in input order.  Ranks may hold different counts; each sends a
    fixed [cap, 18] block (cap = ceil(n_total / world) rounded up to the largest local count via one
    scalar all_reduce) with unused rows marked by index -1."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    local_records = np.asarray(local_records, dtype=np.float32).reshape(-1, RECORD_FLOATS)
    if len(local_ids) != local_records.shape[0]:
        raise ValueError("one record per local object expected")
    out = np.zeros((n_total, RECORD_FLOATS), np.float32)
    if world == 1:
        out[list(local_ids)] = local_records
        return out
    dev = torch.device(device) if device is not None else torch.device("cpu")
    cap = torch.tensor([len(local_ids)], dtype=torch.int64, device=dev)
    dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=group)
    cap = max(int(cap.item()), 1)
    block = torch.full((cap, _ROW), -1.0, dtype=torch.float32)
    if len(local_ids):
        block[:len(local_ids), 0] = torch.tensor(list(local_ids), dtype=torch.float32)
        block[:len(local_ids), 1:] = torch.from_numpy(local_records)
    block = block.to(dev)
    blocks = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(blocks, block, group=group)
    rows = torch.cat(blocks).cpu().numpy()                       # one device->host copy of the [world * cap, 18] block
    rows = rows[rows[:, 0] >= 0]
    ids = rows[:, 0].astype(np.int64)
    hits = np.bincount(ids, minlength=n_total)
    if (hits > 1).any():
        raise RuntimeError(f"objects {np.nonzero(hits > 1)[0].tolist()} were processed by two ranks")
    if (hits[:n_total] == 0).any() or len(hits) > n_total:
        raise RuntimeError(f"objects {np.nonzero(hits[:n_total] == 0)[0].tolist()} were processed by no rank")
    out[ids] = rows[:, 1:]
    return out


def estimate_sharded(estimate: Callable[[int], np.ndarray], costs: Sequence[float], device=None, mode: str = "greedy",
                     group=None) -> np.ndarray:
    """Run ``estimate(i) -> float32[17]`` for the objects assigned to this rank and gather.
    ``estimate`` is typically ``lambda i: est.estimate_fused(pcs[i], nrms[i], seed=i)["record"]``."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    mine = assign_objects(costs, world, mode)[rank]
    recs = np.stack([np.asarray(estimate(i), np.float32) for i in mine]) if mine else np.zeros((0, RECORD_FLOATS), np.float32)
    return gather_records(mine, recs, len(costs), device=device, group=group)
