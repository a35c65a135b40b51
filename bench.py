#!/usr/bin/env python
"""Benchmark of the CPPF hot path (point pairs -> pair MLP -> votes -> pose).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): synthetic NOCS-bottle clouds, N = 4096 points,
bottle constants, ALL N^2 = 16 777 216 ordered point pairs per object ("dense pairs"),
random-init weights of the reference architecture.  One step = one object per rank through
the whole per-object pipeline (point encoder -> pair MLP -> sampling -> centre vote ->
argmax -> back-vote -> second pass -> orientation vote -> pose).  Objects shard over
ranks (weak scaling); the 17-float pose records are gathered once over NCCL at the end of
the timed region.

  value : point-pairs / second, whole job, inputs resident in HBM.
  e2e   : same metric through the public API with HOST buffers: the pinned cloud is copied
          host->device and the pose record device->host inside every timed step.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "point_pairs_per_sec"
UNIT = "pairs/s"


def ncu_constants():
    """Counters that only a profiler can give (DRAM bytes per launch, pipe utilisations), from profiles/ncu_constants.json.
    Every entry names the capture it came from and the sha256 of the kernel's source file at capture time; an entry whose
    source has changed since is DROPPED (reported as null), never printed stale."""
    import hashlib
    path = os.path.join(ROOT, "profiles", "ncu_constants.json")
    out, stale = {}, []
    try:
        table = json.load(open(path))
    except Exception:
        return out, ["profiles/ncu_constants.json missing"]
    for name, ent in table.get("kernels", {}).items():
        try:
            cur = hashlib.sha256(open(os.path.join(ROOT, ent["source"]), "rb").read()).hexdigest()
        except Exception:
            cur = None
        if cur == ent.get("source_sha256"):
            out[name] = ent
        else:
            stale.append(f"{name}: {ent.get('source')} changed since {ent.get('capture')}")
    return out, stale


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-points", type=int, default=4096)
    ap.add_argument("--cpu-sample-pairs", type=int, default=400_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the extra legs (network votes, 100k sampled pairs): "
                                                                "for runs under a profiler")
    ap.add_argument("--votes", default="trained_like", choices=["trained_like", "network"],
                    help="trained_like: after sampling, the (mu,nu,up) bins are replaced by the geometric targets a trained "
                         "network would emit (SURVEY.md 8d i) so that voting runs under a realistic load; network: votes "
                         "come from the random-init network's own samples (cheap: most candidates fall outside the grid)")
    ap.add_argument("--encoder", default="tc", choices=["tc", "simt"],
                    help="fused path's pair encoder: tc = tcgen05 tensor cores (3xTF32), simt = fp32 FFMA warp tiles")
    ap.add_argument("--streams", type=int, default=3,
                    help="fused path: worker streams of the batched entry (cppf_pose_batch) the timed steps are enqueued through; "
                         "1 = one cppf_pose_fused call per object on one stream")
    ap.add_argument("--path", default="fused", choices=["fused", "twopass"],
                    help="fused: encode+sample / privatised vote kernels; twopass: materialised logits like the reference")
    return ap.parse_args()


def workload_config(args):
    return {"workload": f"synthetic NOCS bottle, N={args.n_points} points, dense N^2={args.n_points ** 2} pairs/object, "
                        "bottle constants (res 4e-3, 32 tr bins, 36 rot bins, 72 rots, adaptive), 1 object/rank/step",
            "n_points": args.n_points, "pairs_per_object": args.n_points ** 2, "objects_per_step_per_rank": 1,
            "out_dim": 141, "parallelism": f"objects sharded over {args.gpus} rank(s), one NCCL all_gather of pose records",
            "path": args.path, "votes": args.votes, "encoder": args.encoder,
            "entry": (f"cppf_pose_batch: the K timed objects in one call, {args.streams} worker streams" if args.path == "fused" and args.streams > 1
                      else "one cppf_pose_fused call per object, one stream"),
            "weights": "torch.manual_seed(0) default init of the reference architecture (no checkpoints exist offline)",
            "l2_policy": "per-step working set (bins + tail logits of 16.7M pairs = 403 MB, + 17 MB survivor mask) exceeds the "
                         "126 MB L2; each step is a different cloud"}


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi polled every 20 ms in the background; started before the warm-up steps (the tool needs ~100 ms to
    come up) and filtered to the samples whose timestamp falls inside the timed region [t0, t1]."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, t0=None, t1=None):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, all_sm = [], [], set(), []
        parsed = []
        for r in rows:
            try:
                ts = datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
                parsed.append((ts, float(r[1]), float(r[2]), r[5:9]))
            except Exception:
                continue
        inside = [q for q in parsed if t0 is None or t0 <= q[0] <= t1]
        if len(inside) < 2:                          # polling too coarse for the window: take the samples nearest to it
            mid = 0.5 * ((t0 or 0) + (t1 or 0))
            inside = sorted(parsed, key=lambda q: abs(q[0] - mid))[:4]
        all_sm = [q[1] for q in parsed]
        for ts, c, m, cells in inside:
            sm.append(c); mx.append(m)
            for name, cell in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), cells):
                if "Active" in cell and "Not" not in cell:
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                   "samples_total": len(all_sm)}
        return out


# ------------------------------------------------------------------------------ CPU reference path
def run_cpu_reference(args, steps, warmup, sample_pairs):
    """The reference's own per-object path on the host cores, timed on a bounded pair sample of the SAME object:
    oracle/ref_pipeline.estimate = nocs/inference.py:174-339 restated (torch CPU with all threads for the network -- a PORT
    of models/model.py, pinned to it by the golden fixtures -- and the reference's own CUDA-C voting strings compiled for the
    CPU with OpenMP, oracle/_ref/libref_voting_cpu.so, or the plain-C port if that library did not travel).

    The per-OBJECT cost (kNN + SPRIN point encoder, O(N k); the orientation vote on its 10 000-survivor sub-sample,
    nocs/inference.py:277-284) and the per-PAIR cost (pair MLP, sampling, votes, back-vote, second pass) are timed separately, so that the value can be stated like for like with the GPU arm, which amortises the per-object cost over
    all N^2 pairs of the object:  value = N^2 / (fixed + N^2 * per_pair).  The raw pairs/s of the sample is reported too."""
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core, both in torch and in
    # the OpenMP voting library (which reads the variable when it is first loaded)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from cppf_b200 import model, synth
    from oracle import clib, ref_model, ref_pipeline
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = dict(synth.BOTTLE)
    torch.manual_seed(0)
    sd_pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).state_dict()
    sd_ppf = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).state_dict()
    sphere = ref_model.fibonacci_sphere(480)
    n = args.n_points
    vote_impl = "ref_cpu" if clib.have_ref_cpu() else "oracle"
    tot, fixed, per_pair = [], [], []
    for s in range(warmup + steps):
        pc, nrm = synth.synth_bottle(n, s)
        idxs = synth.sample_pairs(n, sample_pairs, s)
        if args.votes == "trained_like":
            # the vote load of the GPU arm: the first-pass draws are replaced by the geometric targets (SURVEY.md 8d i)
            tl = synth.trained_like_tr(pc, idxs)
            lut_mu = (np.arange(32, dtype=np.float32) / np.float32(31) * np.float32(0.5) - np.float32(0.25))
            lut_nu = np.arange(32, dtype=np.float32) / np.float32(31) * np.float32(0.25)
            q_mu = np.full((sample_pairs, 32), 1e30, np.float32)
            q_nu = np.full((sample_pairs, 32), 1e30, np.float32)
            q_mu[np.arange(sample_pairs), np.argmin(np.abs(tl[:, :1] - lut_mu[None]), -1)] = 1e-30      # the race is decided
            q_nu[np.arange(sample_pairs), np.argmin(np.abs(tl[:, 1:] - lut_nu[None]), -1)] = 1e-30
            noise = {"q_mu": q_mu, "q_nu": q_nu}
        else:
            noise = None
        tm = {}
        t0 = time.perf_counter()
        ref_pipeline.estimate(pc, nrm, sd_pe, sd_ppf, idxs, cfg, noise=noise, seed=s, sphere=sphere, impl=vote_impl,
                              cdist_knn=True, timings=tm)
        dt = time.perf_counter() - t0
        if s >= warmup:
            tot.append(dt)
            if tm["orientation_is_fixed"]:          # the 10 000-survivor sub-sample costs the same for any P
                fixed.append(tm["point_encoder"] + tm["orientation"])
                per_pair.append(tm["pairs"] / sample_pairs)
            else:
                fixed.append(tm["point_encoder"])
                per_pair.append((tm["pairs"] + tm["orientation"]) / sample_pairs)
    fx, pp = float(np.mean(fixed)), float(np.mean(per_pair))
    n2 = float(n) * n
    return {"value": n2 / (fx + n2 * pp), "value_on_sample": sample_pairs * len(tot) / sum(tot),
            "ms_per_step": 1e3 * sum(tot) / len(tot), "vote_kind": "reference" if vote_impl == "ref_cpu" else "port",
            "mlp_kind": "port", "fixed_cost_ms": 1e3 * fx, "per_pair_ms": 1e3 * pp, "cores": torch.get_num_threads(),
            "sample": f"{sample_pairs} random pairs of the N={n} object per step ({len(tot)} steps), votes={args.votes}: "
                      "oracle/ref_pipeline.estimate = nocs/inference.py:174-339 restated -- cdist + topk + SPRIN and the pair "
                      "MLP as a torch-CPU port of models/model.py (all threads), multinomial, vote / back-vote / rot-vote "
                      f"({'the reference CUDA-C strings built for the CPU with OpenMP' if vote_impl == 'ref_cpu' else 'plain-C port'}). "
                      f"value = N^2 / (fixed_cost + N^2 * per_pair): the per-object stages (point encoder + orientation vote on "
                      f"its 10 000-survivor sub-sample, {1e3 * fx:.1f} ms) amortised over all N^2 pairs like the GPU arm; "
                      "value_on_sample = raw pairs/s of the sample"}


def cpu_baseline_block(r):
    # kind "port": the network half is a port of models/model.py (the reference modules cannot travel to the GPU box);
    # the voting half is the reference's own CUDA-C compiled for the CPU whenever oracle/_ref travelled (vote_kind)
    return {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "mlp_kind": r["mlp_kind"],
            "vote_kind": r["vote_kind"], "fixed_cost_ms": r["fixed_cost_ms"], "per_pair_ms": r["per_pair_ms"],
            "value_on_sample": r["value_on_sample"], "sample": r["sample"]}


# ------------------------------------------------------------------------------ GPU reference arm
def gpu_reference_arm(dev, clouds, inject, est, pe, ppf, stage_names, L, timing, args):
    """BASELINE config 2, "vs reference CuPy": the reference's OWN kernels on this GPU (models/voting.py's CUDA-C strings
    built for sm_100a, launched with the reference's launch shapes) + the reference's torch modules on torch-CUDA, in the
    order of nocs/inference.py:174-339 (oracle/ref_gpu_pipeline.py), next to this library on the same inputs:
      (i) the reference's regime, P = 100 000 sampled pairs;  (ii) the dense N^2 workload of this bench, chunked to fit.
    Per-stage CUDA-event times and the end-to-end time per object, both sides."""
    from cppf_b200 import synth
    from cppf_b200.pipeline import PoseConfig, PoseEstimator
    try:
        from oracle import ref_gpu, ref_gpu_pipeline
        if not ref_gpu.available():
            return {"unavailable": "oracle/_ref cubins or cuda-python not present on this box"}
    except Exception as e:                                   # pragma: no cover
        return {"unavailable": repr(e)}
    n = args.n_points
    cfg = dict(synth.BOTTLE)
    sd_pe = {k: v.detach() for k, v in pe.state_dict().items()}
    sd_ppf = {k: v.detach() for k, v in ppf.state_dict().items()}
    sphere = est.sphere
    pc = torch.from_numpy(clouds[0][0]).to(dev)
    nrm = torch.from_numpy(clouds[0][1]).to(dev)
    out = {}

    def ours_stages(e, **kw):
        L.cppf_timing_collect(timing, (C.c_float * len(stage_names))())
        e.timing = timing
        t = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e.enqueue_fused(pc, nrm, seed=0, **kw).result()
            t.append((time.perf_counter() - t0) * 1e3)
        acc = (C.c_float * len(stage_names))()
        calls = L.cppf_timing_collect(timing, acc)
        e.timing = None
        return {nm: acc[i] / max(calls, 1) for i, nm in enumerate(stage_names)}, min(t)

    def table(ref, ours, ours_wall):
        rs = ref["stage_ms"]
        rows = {"point_encoder": (rs.get("point_encoder", 0.0), ours["point_encoder"] + ours["geometry"]),
                "encode": (rs.get("encode", 0.0) + rs.get("encode2", 0.0), ours["preproject"] + ours["encode_sample"]),
                "vote": (rs.get("vote", 0.0), ours["vote"]),
                "argmax": (rs.get("argmax", 0.0), ours["argmax"]),
                "backvote": (rs.get("backvote", 0.0), ours["backvote"] + ours["compact"]),
                "rot_vote+sphere": (rs.get("rot_vote", 0.0), ours["rot_hist"]),
                "aux_sign+stats": (rs.get("aux_sign", 0.0), ours["stats"])}
        tab = {k: {"reference_ms": a, "ours_ms": b, "speedup": (a / b if b > 0 else None)} for k, (a, b) in rows.items()}
        r_tot, o_tot = sum(a for a, _ in rows.values()), sum(b for _, b in rows.values())
        return {"stages": tab, "reference_gpu_ms_per_object": r_tot, "ours_gpu_ms_per_object": o_tot,
                "gpu_speedup": r_tot / o_tot, "reference_wall_ms_per_object": ref["wall_ms"], "ours_wall_ms_per_object": ours_wall,
                "reference_objects_per_sec": 1e3 / ref["wall_ms"], "ours_objects_per_sec": 1e3 / ours_wall,
                "reference_survivors": ref["n_survivors"]}

    # (i) P = 100 000 sampled pairs, the network's own votes on both sides
    p = 100000
    idx = torch.from_numpy(synth.sample_pairs(n, p, 0).astype(np.int32)).to(dev)
    ref = None
    for _ in range(3):                                        # first call loads the cubins / warms cuBLAS
        r = ref_gpu_pipeline.estimate(pc, nrm, sd_pe, sd_ppf, idx, cfg, sphere, seed=0)
        if ref is None or r["wall_ms"] < ref["wall_ms"]:
            ref = r
    est_s = PoseEstimator(pe, ppf, PoseConfig.from_dict(dict(synth.BOTTLE, n_pairs=p)), dev)
    o_st, o_wall = ours_stages(est_s, idxs=idx)
    out["sampled_100k"] = table(ref, o_st, o_wall)
    # (ii) dense N^2 pairs with the trained-like vote load of the headline number
    inj = inject[0]
    if inj is not None:
        lut = est.lut
        b = inj.long()
        tr = torch.stack([lut[b[:, 0]], lut[32 + b[:, 1]]], -1).contiguous()
        ii = torch.arange(n, device=dev, dtype=torch.int32)
        idx_d = torch.stack([ii[:, None].expand(n, n), ii[None, :].expand(n, n)], -1).reshape(-1, 2).contiguous()
        ref = None
        for _ in range(2):
            r = ref_gpu_pipeline.estimate(pc, nrm, sd_pe, sd_ppf, idx_d, cfg, sphere, seed=0, inject_tr=tr)
            if ref is None or r["wall_ms"] < ref["wall_ms"]:
                ref = r
        del idx_d, tr
        o_st, o_wall = ours_stages(est, inject_bins=inj)
        out["dense_n2_trained_like"] = table(ref, o_st, o_wall)
    out["note"] = ("reference = models/voting.py kernel strings compiled for sm_100a (oracle/_ref/*.cubin) with the reference's "
                   "launch shapes + the reference modules restated on torch-CUDA (cuBLAS SGEMM, ATen), host round trips as in "
                   "nocs/inference.py; ours = cppf_pose_fused.  Same cloud, same weights, same pairs.")
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------ main
def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    cfgj = workload_config(args)

    if args.impl == "reference":
        if rank != 0:
            return
        r = run_cpu_reference(args, max(1, args.steps), max(0, min(args.warmup, 1)), args.cpu_sample_pairs)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfgj,
                "cpu_baseline": cpu_baseline_block(r),
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    from cppf_b200 import _lib, model, shard, synth
    from cppf_b200.pipeline import PoseConfig, PoseEstimator, enqueue_batch

    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout when the communicator comes up; stdout carries exactly one JSON line,
        # so file descriptor 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    # nvidia-smi needs up to a second to deliver its first sample on a fresh box: start it before the inputs are built
    sampler = ClockSampler(local) if rank == 0 else None
    torch.manual_seed(0)
    pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(dev).eval()
    ppf = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).to(dev).eval()
    pcfg = PoseConfig.from_dict(dict(synth.BOTTLE, n_pairs=0))        # 0 = all N^2 ordered pairs
    est = PoseEstimator(pe, ppf, pcfg, dev)
    est.encoder_impl = args.encoder
    n = args.n_points
    pairs_per_obj = n * n
    n_obj = args.warmup + args.steps
    clouds = [synth.synth_bottle(n, 1000 * rank + s) for s in range(n_obj)]
    pinned = [(torch.from_numpy(p).pin_memory(), torch.from_numpy(q).pin_memory()) for p, q in clouds]
    h2d = 2 * n * 3 * 4             # xyz + normals of the cloud (float32)
    d2h = 16 * 8                    # the pose record (16 doubles)

    def barrier():
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()

    cells = [int(np.prod(synth.vote_grid_geometry(p, synth.BOTTLE["res"])[1])) for p, _ in clouds]
    inject = [None] * n_obj
    if args.votes == "trained_like" and args.path == "fused":
        inject = [synth.trained_like_bins_dense_torch(torch.from_numpy(p).to(dev), synth.BOTTLE) for p, _ in clouds]

    L = _lib.lib()
    timing = L.cppf_timing_create() if args.path == "fused" else None
    if timing:
        L.cppf_timing_reserve(timing, max(args.steps, 200))
    stage_names = [L.cppf_timing_stage_name(i).decode() for i in range(L.cppf_timing_stages())]

    def run(leg, votes_injected=True, batched=False):
        """leg 'hbm': clouds already on the device; leg 'e2e': pinned host buffers in, pose record out.
        The fused path enqueues every step with ONE library call (cppf_pose_fused) and never waits for the GPU
        inside the loop: the records come back through pinned buffers and are read after the last enqueue."""
        records = []
        resident = [(p.to(dev), q.to(dev)) for p, q in pinned] if leg == "hbm" else None
        timers = {}
        if args.path == "fused":
            def step(a, b, seed):
                return est.enqueue_fused(a, b, seed=seed, inject_bins=inject[seed] if votes_injected else None,
                                         max_cells=cells[seed])
        else:
            est.timers = timers
            step = lambda a, b, seed: est.estimate(a, b, seed=seed)
        for s in range(args.warmup):
            src = resident[s] if leg == "hbm" else pinned[s]
            r = step(src[0], src[1], seed=s)
            if args.path == "fused":
                r.result()
        if batched:                     # warm the worker streams / workspaces of the batch entry too
            w_items = [(est, (resident[s] if leg == "hbm" else pinned[s])[0], (resident[s] if leg == "hbm" else pinned[s])[1], s)
                       for s in range(min(args.warmup, 3))]
            enqueue_batch(w_items, n_streams=args.streams, n_threads=args.streams,
                          inject_bins=[inject[s] if votes_injected else None for s in range(len(w_items))],
                          capacities=[(cells[s], 0) for s in range(len(w_items))]).results(on_error="none")
        timers.clear()
        if timing:
            L.cppf_timing_collect(timing, (C.c_float * len(stage_names))())      # drop the warm-up marks
        est.timing = None if batched else timing
        barrier()
        l0 = _lib.launch_count()
        t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if batched:
            # the K timed steps as ONE library call (cppf_pose_batch): the objects are fanned out over a few worker streams,
            # so the dozen small kernels of one object run beside the large kernels of the next
            items = [(est, (resident[s] if leg == "hbm" else pinned[s])[0], (resident[s] if leg == "hbm" else pinned[s])[1], s)
                     for s in range(args.warmup, n_obj)]
            pb = enqueue_batch(items, n_streams=args.streams, n_threads=args.streams,
                               inject_bins=[inject[s] if votes_injected else None for s in range(args.warmup, n_obj)],
                               capacities=[(cells[s], 0) for s in range(args.warmup, n_obj)])
            records = list(pb.records17())
        else:
            pend = []
            for s in range(args.warmup, n_obj):
                src = resident[s] if leg == "hbm" else pinned[s]
                pend.append(step(src[0], src[1], seed=s))
            records = [(p.result() if args.path == "fused" else p)["record"] for p in pend]
        rec = np.stack(records)
        ids = [rank * args.steps + i for i in range(args.steps)]
        shard.gather_records(ids, rec, world * args.steps, device=dev)    # the one collective: pose hypotheses
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = (t0, time.time())           # window of this leg; the samples are filtered when the sampler stops
        launches = _lib.launch_count() - l0
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if dist_on:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        est.timers = None
        est.timing = None
        stage_ms = {}
        if timing:
            acc = (C.c_float * len(stage_names))()
            calls = L.cppf_timing_collect(timing, acc)
            if calls > 0:
                stage_ms = {nm: {"avg_ms": acc[i] / calls, "launches": calls} for i, nm in enumerate(stage_names)}
        else:
            for name, evs in timers.items():
                durs = [a.elapsed_time(b) for a, b in evs]
                if durs:
                    stage_ms[name] = {"avg_ms": sum(durs) / len(durs), "launches": len(durs)}
        return float(t.item()), launches, clocks, stage_ms

    use_batch = args.path == "fused" and args.streams > 1
    ms_hbm, launches, window, timers = run("hbm", batched=use_batch)
    clocks = sampler.stop(*window) if sampler else None
    ms_e2e, _, _, _ = run("e2e", batched=use_batch)
    ms_inst = ms_hbm
    if use_batch:       # per-kernel times: a separate pass over the same objects, one stream, CUDA events around every stage
        ms_inst, _, _, timers = run("hbm", batched=False)
    ms_net = None
    if args.votes == "trained_like" and args.path == "fused" and not args.no_variants:
        ms_net, _, _, _ = run("hbm", votes_injected=False, batched=use_batch)
    sampled = None
    if args.path == "fused" and not args.no_variants:
        # The reference's own inference regime (nocs/inference.py:120-129,177): 100 000 random pairs per object, clouds in
        # host memory.  The whole object loop is ONE library call (cppf_pose_batch through pipeline.enqueue_batch): one
        # pinned block + one H2D for all clouds, pairs drawn on the device, objects fanned out over worker streams (one
        # object fills ~1.3 waves per kernel), launches issued by several host threads, one D2H for all records.
        from cppf_b200.pipeline import enqueue_batch
        est_s = PoseEstimator(pe, ppf, PoseConfig.from_dict(dict(synth.BOTTLE, n_pairs=100000)), dev)
        n_s = 240
        items = [(est_s, clouds[s_ % n_obj][0], clouds[s_ % n_obj][1], s_) for s_ in range(n_s)]
        enqueue_batch(items[:16]).results(on_error="none")
        barrier()
        sweep = {}
        best = None
        for n_str, n_thr in ((1, 1), (2, 2), (4, 4), (8, 4), (8, 8), (6, 3)):
            reps, host = [], []
            for _ in range(3):
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0 = time.perf_counter()
                a0.record()
                pend = enqueue_batch(items, n_streams=n_str, n_threads=n_thr)
                host.append((time.perf_counter() - t0) * 1e6 / n_s)
                pend.records17()                       # the host tail of every object, column-wise
                a1.record()
                barrier()
                reps.append(a0.elapsed_time(a1))
            sweep[f"streams{n_str}_threads{n_thr}"] = {"objects_per_sec": n_s / (statistics.median(reps) * 1e-3),
                                                        "host_enqueue_us_per_object": statistics.median(host)}
            if best is None or statistics.median(reps) < best[0]:
                best = (statistics.median(reps), reps, statistics.median(host), n_str, n_thr)
        ms_1, reps, host_us, n_str, n_thr = best
        # host cost of the enqueue alone, on an idle GPU with batches small enough (8 and 32 objects, < 1000 launches) that
        # the launch queue never fills and the call returns long before the kernels finish: the slope is the cost per
        # object (packing + ~22 driver calls spread over the host threads), the intercept the cost per call
        free = {}
        for k_ in (8, 32):
            t_best = None
            for _ in range(7):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                pend = enqueue_batch(items[:k_], n_streams=n_str, n_threads=n_thr)
                dt = (time.perf_counter() - t0) * 1e6
                pend.records17()
                t_best = dt if t_best is None else min(t_best, dt)
            free[k_] = t_best
        host_free_us = (free[32] - free[8]) / 24.0
        host_call_us = free[8] - 8 * host_free_us
        # per-object calls on one stream (round 1's path) for comparison, and the stage breakdown of one object
        def enq(s_):
            return est_s.enqueue_fused(pinned[s_ % n_obj][0], pinned[s_ % n_obj][1], seed=s_, max_cells=cells[s_ % n_obj])
        for s_ in range(3):
            enq(s_).result()
        barrier()
        ms_loop = None
        for _ in range(3):              # the first pass grows torch's caching allocator (240 objects' pair lists alive at once)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            pend = [enq(s_) for s_ in range(n_s)]
            for q in pend:
                q.result()
            a1.record()
            barrier()
            ms_loop = a0.elapsed_time(a1) if ms_loop is None else min(ms_loop, a0.elapsed_time(a1))
        stage_s = {}
        if timing:                      # stage breakdown from a separate short run (event records slow short objects down)
            L.cppf_timing_collect(timing, (C.c_float * len(stage_names))())
            est_s.timing = timing
            for s_ in range(20):
                enq(s_).result()
            acc = (C.c_float * len(stage_names))()
            calls = L.cppf_timing_collect(timing, acc)
            est_s.timing = None
            stage_s = {nm: round(acc[i] / max(calls, 1), 4) for i, nm in enumerate(stage_names)}
        sampled = {"objects_per_sec_per_gpu": n_s / (ms_1 * 1e-3), "pairs_per_sec_per_gpu": n_s * 100000 / (ms_1 * 1e-3),
                   "ms_per_object": ms_1 / n_s, "ms_per_object_repeats": [r / n_s for r in reps],
                   "host_enqueue_us_per_object": host_free_us, "host_enqueue_us_per_call": host_call_us,
                   "host_enqueue_us_per_object_queue_full": host_us,
                   "n_streams": n_str, "n_threads": n_thr, "sweep": sweep,
                   "per_object_calls_one_stream": {"objects_per_sec": n_s / (ms_loop * 1e-3), "ms_per_object": ms_loop / n_s},
                   "stage_ms": stage_s, "pairs_per_object": 100000, "objects_per_batch": n_s,
                   "note": "reference regime: P = 100 000 sampled pairs (nocs/inference.py:177), host clouds in, pose records "
                           "out, the object loop (nocs/inference.py:120-129) as ONE cppf_pose_batch call; timed region = "
                           "packing + H2D + all kernels + D2H + host tail of all objects; host_enqueue_us_per_object = slope of "
                           "the enqueue call's host time between 8 and 32 objects on an idle GPU (launch queue never full), "
                           "host_enqueue_us_per_call its intercept, ..._queue_full = the same call's time / 240 objects when "
                           "the GPU is the bottleneck and the launch queue pushes back"}
    trained = None
    ckpt = os.path.join(ROOT, "tests", "golden", "trained_bottle.npz")
    if args.path == "fused" and not args.no_variants and os.path.exists(ckpt):
        # the same dense workload with the TRAINED checkpoint (oracle/train_synth_bottle.py: the reference modules trained on
        # this synthetic bottle) and NO bin injection: the votes are the network's own
        d = np.load(ckpt)
        pe_t = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32)
        ppf_t = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141)
        pe_t.load_state_dict({k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("pe/")})
        ppf_t.load_state_dict({k[4:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("ppf/")})
        est_t = PoseEstimator(pe_t.to(dev).eval(), ppf_t.to(dev).eval(), pcfg, dev)
        res_t = [(torch.from_numpy(p).to(dev), torch.from_numpy(q).to(dev)) for p, q in clouds[:8]]
        for s_ in range(2):
            est_t.enqueue_fused(res_t[s_][0], res_t[s_][1], seed=s_, max_cells=cells[s_]).result()
        if timing:
            L.cppf_timing_collect(timing, (C.c_float * len(stage_names))())
            est_t.timing = timing
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        pend = [est_t.enqueue_fused(res_t[s_][0], res_t[s_][1], seed=s_, max_cells=cells[s_]) for s_ in range(8)]
        poses = [q.result() for q in pend]
        a1.record()
        barrier()
        ms_t = a0.elapsed_time(a1) / 8
        stage_t = {}
        if timing:
            acc = (C.c_float * len(stage_names))()
            calls = L.cppf_timing_collect(timing, acc)
            est_t.timing = None
            stage_t = {nm: round(acc[i] / max(calls, 1), 4) for i, nm in enumerate(stage_names)}
        err_T = float(np.mean([np.linalg.norm(q["T_host"]) for q in poses]))          # the bottles are centred at the origin
        trained = {"value": pairs_per_obj / (ms_t * 1e-3), "unit": UNIT, "ms_per_step": ms_t, "stage_ms": stage_t,
                   "mean_survivors": float(np.mean([q["n_survivors"] for q in poses])),
                   "mean_centre_error_m": err_T, "mean_abs_up_y": float(np.mean([abs(q["up"][1]) for q in poses])),
                   "note": "dense N^2 pairs, weights = tests/golden/trained_bottle.npz (the reference modules trained on the "
                           "synthetic bottle), votes = the network's own draws (no injection); the poses are checked: centre "
                           "error in metres (res = 4e-3), |up . y| (the bottles stand along y)"}
    # ---- algorithmic work of the vote launch and the same-run peaks, measured (untimed)
    vote_work = None
    if rank == 0 and args.path == "fused":
        pcd = torch.from_numpy(clouds[args.warmup][0]).to(dev)
        corner_np, dims = synth.vote_grid_geometry(clouds[args.warmup][0], synth.BOTTLE["res"])
        cnt = torch.zeros(3, dtype=torch.int64, device=dev)
        sp = torch.cuda.current_stream(dev).cuda_stream
        inj = inject[args.warmup]
        if inj is not None:
            bins4 = torch.zeros((pairs_per_obj, 4), dtype=torch.uint8, device=dev)
            bins4[:, :inj.shape[1]] = inj
            _lib.check(L.cppf_vote_count(pcd.data_ptr(), None, bins4.data_ptr(), est.lut.data_ptr(), None, 0,
                                         torch.from_numpy(corner_np).to(dev).data_ptr(), float(synth.BOTTLE["res"]), n,
                                         pairs_per_obj, 72, dims[0], dims[1], dims[2], 1, cnt.data_ptr(), sp), "cppf_vote_count")
            c = cnt.cpu().numpy()
            pk = (C.c_double * 3)()
            _lib.check(L.cppf_peak_shared_atomics(dims[0], dims[1], dims[2], 0, 5, C.byref(pk, 0), sp), "peak")
            _lib.check(L.cppf_peak_shared_atomics(dims[0], dims[1], dims[2], 1, 5, C.byref(pk, 8), sp), "peak")
            _lib.check(L.cppf_peak_global_red(dims[0], dims[1], dims[2], 3, C.byref(pk, 16), sp), "peak")
            vote_work = {"rotation_steps": int(c[0]), "in_bounds_candidates": int(c[1]), "live_pairs": int(c[2]),
                         "atomics": int(c[1]) * 8, "grid_dims": list(dims),
                         "peak_shared_random_bank_gatoms": pk[0], "peak_shared_conflict_free_gatoms": pk[1],
                         "peak_global_fp32_red_gatoms": pk[2]}
            del bins4
    gpu_ref = None
    if rank == 0 and args.path == "fused" and not args.no_variants:
        gpu_ref = gpu_reference_arm(dev, clouds, inject, est, pe, ppf, stage_names, L, timing, args)
    total_pairs = world * args.steps * pairs_per_obj
    value = total_pairs / (ms_hbm * 1e-3)
    e2e = total_pairs / (ms_e2e * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        tensor_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))      # kernels are timed inside a long step
        peak_src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
        ncu, ncu_stale = ncu_constants()
        kern = dict(timers)
        # Algorithmic work per launch.  Dense pairs are enumerated in-kernel (no index read).
        #   encode_sample : 24 B/pair written (4 bin bytes + 5 tail floats); 23 968 FLOP/pair canonical pair MLP
        #                   (models/model.py:12-23,87; 13.9 k executed after the per-point pre-projection of layer 0)
        #   vote          : 4 B/pair read (bins); atomics = 8 x in-bounds candidates, COUNTED in this run (cppf_vote_count)
        #   twopass       : first-pass encode writes 64 fp32 logits/pair, vote reads 8 B (mu,nu)/pair
        algo_bytes = {"encode_sample": pairs_per_obj * 24, "vote": pairs_per_obj * 4, "backvote": pairs_per_obj * 5,
                      "stats": pairs_per_obj * 24,
                      "ppf_encode_pass1": pairs_per_obj * 64 * 4, "ppf_vote": pairs_per_obj * 8}
        algo_flops = {"encode_sample": pairs_per_obj * 23968.0, "ppf_encode_pass1": pairs_per_obj * (23968.0 - 2 * 16 * 77)}
        binding = {"encode_sample": "SIMT epilogues between the tcgen05 MMA steps (3xTF32 chain)" if args.encoder == "tc" else "fp32 FMA pipe",
                   "vote": "shared-memory atomic pipe (random-bank replays) and issue slots",
                   "backvote": "fp32 / issue", "ppf_vote": "L2 atomic throughput", "stats": "HBM gather",
                   "ppf_encode_pass1": "fp32 FMA pipe", "point_encoder": "kNN: issue / LSU (two-sweep select); SPRIN: SIMT LayerNorm epilogues between the tcgen05 MMA steps (3xTF32, two points per 128-row tile)"}
        detail = {}
        for k, v in kern.items():
            t = v["avg_ms"] * 1e-3
            d = {"avg_ms": v["avg_ms"], "binding_resource": binding.get(k)}
            if k in ncu:                 # profiler-only counters, with the capture and source hash they belong to
                d["ncu"] = {kk: vv for kk, vv in ncu[k].items() if kk != "source_sha256"}
            if k in algo_bytes:
                d["hbm"] = {"achieved": algo_bytes[k] / t / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": algo_bytes[k] / t / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": algo_bytes[k]}
                v["achieved_gbs"] = d["hbm"]["achieved"]
                v["pairs_per_s"] = pairs_per_obj / t
            if k in algo_flops:
                d["tensor"] = {"achieved": algo_flops[k] / t / 1e12, "peak": tensor_peak, "unit": "TFLOP/s",
                               "frac": algo_flops[k] / t / 1e12 / tensor_peak,
                               "algorithmic_flops_per_launch": algo_flops[k],
                               "note": "canonical fp32 pair-MLP FLOPs against the measured dense bf16 peak; the kernel runs "
                                       "tf32 (half the bf16 rate) in 3 passes for fp32-grade logits"}
            if k == "vote" and vote_work is not None:
                # the vote phase in atomics/s.  Count and peaks are measured in THIS run: cppf_vote_count
                # walks the same bins with the reference's acceptance test; cppf_peak_* time the bare patterns.
                n_atom = float(vote_work["atomics"])
                d["atomics"] = {"per_launch": n_atom, "per_pair": n_atom / pairs_per_obj, "achieved": n_atom / t / 1e9,
                                "unit": "G atomics/s", "peak_random_bank": vote_work["peak_shared_random_bank_gatoms"],
                                "frac": n_atom / t / 1e9 / vote_work["peak_shared_random_bank_gatoms"],
                                "peak_conflict_free": vote_work["peak_shared_conflict_free_gatoms"],
                                "frac_of_conflict_free": n_atom / t / 1e9 / vote_work["peak_shared_conflict_free_gatoms"],
                                "global_fp32_red_same_grid": vote_work["peak_global_fp32_red_gatoms"],
                                "work": vote_work,
                                "note": "peak_random_bank: 32 lanes x the 8-corner splat of a random cell each, nothing else in "
                                        "the loop (what any unsorted scatter into shared memory can reach); "
                                        "peak_conflict_free: lane l in bank l (the hardware roof); all measured in this process"}
            detail[k] = d
        roof = None
        if "vote" in detail and "atomics" in detail["vote"] and kern["vote"]["avg_ms"] >= max(v["avg_ms"] for v in kern.values()):
            at = detail["vote"]["atomics"]
            roof = {"kernel": "vote", "bound": "smem_atomics", "achieved": at["achieved"], "peak": at["peak_random_bank"],
                    "unit": "G atomics/s", "frac": at["frac"], "traffic": (ncu.get("vote") or {}).get("dram_bytes_per_launch"),
                    "peak_source": "cppf_peak_shared_atomics, same process (random-bank 8-corner pattern on the same grid)",
                    "peak_conflict_free": at["peak_conflict_free"], "frac_of_conflict_free": at["frac_of_conflict_free"],
                    "algorithmic_atomics_per_launch": at["per_launch"], "avg_launch_ms": kern["vote"]["avg_ms"],
                    "hbm_frac_for_the_record": detail["vote"]["hbm"]["frac"],
                    "note": "dominant kernel by live CUDA-event time.  It accumulates into a grid privatised in shared memory: "
                            "its HBM traffic is 4 B/pair (hbm_frac_for_the_record) and says nothing; the resource it uses is "
                            "the shared-memory atomic pipe, so achieved / peak are in atomics/s (SURVEY.md 8d), count and "
                            "peak both measured in this run"}
        else:
            timed = [k for k in kern if k in algo_bytes]
            if timed:
                dom = max(timed, key=lambda k: kern[k]["avg_ms"])
                use_tensor = dom in algo_flops and args.encoder == "tc"
                src = detail[dom]["tensor" if use_tensor else "hbm"]
                roof = {"kernel": dom, "bound": "tensor" if use_tensor else "hbm", "achieved": src["achieved"],
                        "peak": src["peak"], "unit": src["unit"], "frac": src["frac"],
                        "traffic": (ncu.get(dom) or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                        "avg_launch_ms": kern[dom]["avg_ms"], "binding_resource": binding.get(dom)}
        cpu = None
        if not args.no_cpu_baseline:
            cpu = cpu_baseline_block(run_cpu_reference(args, 2, 1, args.cpu_sample_pairs))
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_hbm / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": cfgj,
                "objects_per_sec": world * args.steps / (ms_hbm * 1e-3),
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "objects_per_sec": world * args.steps / (ms_e2e * 1e-3), "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "roofline_detail": detail, "kernels": kern,
                "cpu_baseline": cpu, "ncu_constants_dropped_as_stale": ncu_stale,
                "kernels_pass": {"ms_per_step": ms_inst / args.steps, "streams": 1,
                                 "note": "`kernels`, `roofline` and `roofline_detail` come from a pass over the same K objects "
                                         "enqueued one cppf_pose_fused call at a time on ONE stream with CUDA events around every "
                                         "stage (under the multi-stream batch entry the kernels of neighbouring objects overlap "
                                         "and per-kernel event times mean nothing); a kernel's share of the step is its avg_ms "
                                         "over this pass's ms_per_step"} if use_batch else None}
        if ms_net is not None:
            line["variant_network_votes"] = {"value": total_pairs / (ms_net * 1e-3), "unit": UNIT,
                                             "ms_per_step": ms_net / args.steps,
                                             "note": "no bin injection: votes from the random-init network's own samples"}
        if trained is not None:
            line["variant_trained_network"] = trained
        if sampled is not None:
            line["variant_sampled_100k"] = sampled
        if gpu_ref is not None:
            line["vs_gpu_reference"] = gpu_ref
        print(json.dumps(line))
    if dist_on:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
