"""Strong scaling of ONE dense object over the GPUs of a node by splitting its pair rows (cppf_b200/rowsplit.py).
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port P \
             tools/bench_rowsplit.py --n-points 16384 [--steps 5]
Rank 0 prints one JSON line: ms per object (CUDA events, max over ranks), pairs/s, and the single-GPU one-call time
of the same object measured on rank 0 for reference."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cppf_b200 import model, rowsplit, synth                      # noqa: E402
from cppf_b200.pipeline import PoseConfig, PoseEstimator          # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-points", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        saved = os.dup(1)
        os.dup2(2, 1)                                             # NCCL's banner goes to stderr, not into the JSON line
        dist.init_process_group("nccl", device_id=torch.device(dev))
        dist.barrier()
        os.dup2(saved, 1)
    torch.manual_seed(0)
    pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(dev).eval()
    ppf = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).to(dev).eval()
    est = PoseEstimator(pe, ppf, PoseConfig.from_dict(dict(synth.BOTTLE, n_pairs=0)), dev)
    n = args.n_points
    pc_np, nrm_np = synth.synth_bottle(n, 0)
    pc, nrm = torch.from_numpy(pc_np).to(dev), torch.from_numpy(nrm_np).to(dev)
    inj = synth.trained_like_bins_dense_torch(pc, synth.BOTTLE)

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out

    ms, pose = timed(lambda: rowsplit.estimate_rowsplit(est, pc, nrm, seed=0, inject_bins=inj))
    single = None
    if rank == 0:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(args.warmup + args.steps):
            if i == args.warmup:
                torch.cuda.synchronize()
                e0.record()
            ref = est.estimate_fused(pc, nrm, seed=0, inject_bins=inj)
        e1.record()
        torch.cuda.synchronize()
        single = e0.elapsed_time(e1) / args.steps
        line = {"bench": "rowsplit", "n_points": n, "pairs": n * n, "n_gpus": world, "ms_per_object": ms,
                "pairs_per_sec": n * n / (ms * 1e-3), "single_gpu_one_call_ms": single, "speedup_vs_single": single / ms,
                "same_argmax_as_single": bool(pose["argmax_flat"] == ref["argmax_flat"]),
                "T_split": [float(v) for v in pose["T_host"]], "T_single": [float(v) for v in ref["T_host"]],
                "same_survivors_as_single": bool(pose["n_survivors"] == ref["n_survivors"]),
                "note": "dense row-block mode (cppf_*_rows: in-kernel enumeration of the rank's rows, no pair list); exchanges: "
                        "u64 vote grid, orientation histogram, survivor statistics (3 all_reduce); the Philox stream is keyed "
                        "by the pair's index in the whole matrix, votes are injected trained-like bins"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
