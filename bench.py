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
# dram__bytes_read.sum + dram__bytes_write.sum per launch and the utilisation of the resource that binds each kernel,
# from the committed `ncu --set full` capture of this pipeline (profiles/r1f_kernels_ncu.md), N=4096 dense
TRAFFIC_NCU = {"encode_sample": 3.3728e6 + 345.649152e6, "vote": 67.36128e6 + 0.84736e6,
               "backvote": 67.252736e6 + 3.7312e6, "stats": 285.441792e6 + 7.912192e6}
BINDING_NCU = {"encode_sample": {"issue_slots_busy": 0.456, "tensor_pipe_active": 0.438, "warps_per_sm": 16},
               "vote": {"shared_memory_wavefronts_of_peak": 0.738, "issue_slots_busy": 0.772,
                        "wavefronts_per_ATOMS": 3.64},
               "backvote": {"issue_slots_busy": 0.778}, "stats": {"dram_read_tbs": 2.88}}


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
    ap.add_argument("--path", default="fused", choices=["fused", "twopass"],
                    help="fused: encode+sample / privatised vote kernels; twopass: materialised logits like the reference")
    return ap.parse_args()


def workload_config(args):
    return {"workload": f"synthetic NOCS bottle, N={args.n_points} points, dense N^2={args.n_points ** 2} pairs/object, "
                        "bottle constants (res 4e-3, 32 tr bins, 36 rot bins, 72 rots, adaptive), 1 object/rank/step",
            "n_points": args.n_points, "pairs_per_object": args.n_points ** 2, "objects_per_step_per_rank": 1,
            "out_dim": 141, "parallelism": f"objects sharded over {args.gpus} rank(s), one NCCL all_gather of pose records",
            "path": args.path, "votes": args.votes, "encoder": args.encoder,
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
def cpu_reference_step(pc, nrm, sd_pe, sd_ppf, idxs, cfg, sphere, seed):
    """The reference's own per-object path on the CPU for a bounded pair sample: oracle
    restatement of the torch modules (reference models/model.py, run by torch CPU with all
    threads) + the reference's CUDA-C voting strings compiled for the CPU with OpenMP
    (oracle/_ref/libref_voting_cpu.so; falls back to the plain-C port if it did not travel)."""
    from oracle import clib, ref_model
    impl = "ref_cpu" if clib.have_ref_cpu() else "oracle"
    n = pc.shape[0]
    tpc, tn = torch.from_numpy(pc), torch.from_numpy(nrm)
    dist = torch.cdist(tpc[None], tpc[None])[0]
    feat = ref_model.point_encode(tpc, tn, dist, sd_pe, cfg["knn"])
    logits = ref_model.ppf_encode_idx(tpc, tn, feat, idxs, sd_ppf)
    B = cfg["tr_num_bins"]
    g = torch.Generator().manual_seed(seed)
    pr = torch.softmax(logits[:, :2 * B].reshape(-1, 2, B), -1)
    bins = torch.cat([torch.multinomial(pr[:, 0], 1, generator=g), torch.multinomial(pr[:, 1], 1, generator=g)], -1)
    tr = ref_model.decode_tr(bins[:, 0], bins[:, 1], B, cfg["vote_range"]).numpy()
    lo, hi = pc.min(0), pc.max(0)
    dims = ((hi - lo) / cfg["res"]).astype(np.int32) + 1
    idx32 = idxs.astype(np.int32)
    grid = clib.ppf_voting(pc, tr, np.ones(n, np.float32), idx32, dims, lo, cfg["res"], 72, True, impl=impl)
    flat, centre = ref_model.centre_from_grid(grid, lo, cfg["res"])
    oc = clib.backvote(pc, tr, idx32, dims, lo, cfg["res"], centre.astype(np.float32), 3 * cfg["res"], 72, impl=impl)
    kept = idxs[np.any(oc != 0, -1)]
    if len(kept):
        l2 = ref_model.ppf_encode_idx(tpc, tn, feat, kept, sd_ppf)
        up = torch.multinomial(torch.softmax(l2[:, 2 * B:2 * B + cfg["rot_num_bins"]], -1), 1, generator=g)[:, 0]
        rot = ref_model.decode_rot(up, cfg["rot_num_bins"]).numpy().astype(np.float32)
        sub = np.random.default_rng(seed).permutation(len(kept))[:10000]
        cand = clib.rot_voting(pc, rot[sub], kept[sub].astype(np.int32), 72, impl=impl)
        counts = ((torch.from_numpy(cand.reshape(-1, 3)) @ torch.from_numpy(sphere.T.astype(np.float32))) >
                  float(np.cos(1.5 / 180 * np.pi))).sum(0)
        best = sphere[int(torch.argmax(counts))]
        ref_model.aux_sign(pc, nrm, kept, best, l2[:, -5].numpy())
    return int(flat), impl


def run_cpu_reference(args, steps, warmup, sample_pairs):
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core, both in torch and in
    # the OpenMP voting library (which reads the variable when it is first loaded)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from cppf_b200 import model, synth
    from oracle import ref_model
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = synth.BOTTLE
    torch.manual_seed(0)
    sd_pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).state_dict()
    sd_ppf = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).state_dict()
    sphere = ref_model.fibonacci_sphere(480)
    n = args.n_points
    times, impl = [], "oracle"
    for s in range(warmup + steps):
        pc, nrm = synth.synth_bottle(n, s)
        idxs = synth.sample_pairs(n, sample_pairs, s)
        t0 = time.perf_counter()
        _, impl = cpu_reference_step(pc, nrm, sd_pe, sd_ppf, idxs, cfg, sphere, s)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    total = sum(times)
    return {"value": sample_pairs * len(times) / total, "ms_per_step": 1e3 * total / len(times), "impl": impl,
            "cores": torch.get_num_threads(),
            "sample": f"{sample_pairs} random pairs of the N={n} object per step ({len(times)} steps): point encoder + "
                      "pair MLP (torch CPU, all threads) + multinomial + vote/back-vote/rot-vote "
                      f"({'reference CUDA-C strings built for CPU, OpenMP' if impl == 'ref_cpu' else 'plain-C oracle port'})"}


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
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"],
                                 "kind": "reference" if r["impl"] == "ref_cpu" else "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    from cppf_b200 import _lib, model, shard, synth
    from cppf_b200.pipeline import PoseConfig, PoseEstimator

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

    def run(leg, votes_injected=True):
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
        timers.clear()
        if timing:
            L.cppf_timing_collect(timing, (C.c_float * len(stage_names))())      # drop the warm-up marks
        est.timing = timing
        barrier()
        l0 = _lib.launch_count()
        t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
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

    ms_hbm, launches, window, timers = run("hbm")
    clocks = sampler.stop(*window) if sampler else None
    ms_e2e, _, _, _ = run("e2e")
    ms_net = None
    if args.votes == "trained_like" and args.path == "fused" and not args.no_variants:
        ms_net, _, _, _ = run("hbm", votes_injected=False)
    sampled = None
    if args.path == "fused" and not args.no_variants:
        # the reference's own inference regime (nocs/inference.py:177): 100 000 random pairs per object, from pinned host
        # clouds; one cppf_pose_fused call per object, all objects of the run enqueued back to back.  The kernels of one
        # such object are short (0.5 ms in ~25 launches), so objects are also dealt round-robin to a few CUDA streams.
        est_s = PoseEstimator(pe, ppf, PoseConfig.from_dict(dict(synth.BOTTLE, n_pairs=100000)), dev)
        n_s = 240

        def enq(s_):
            return est_s.enqueue_fused(pinned[s_ % n_obj][0], pinned[s_ % n_obj][1], seed=s_, max_cells=cells[s_ % n_obj])
        for s_ in range(3):
            enq(s_).result()
        barrier()
        host_us = []
        for s_ in range(20):            # host cost of one enqueue with an idle GPU
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            q = enq(s_)
            host_us.append((time.perf_counter() - t0) * 1e6)
            q.result()
        barrier()
        reps = []
        for _ in range(3):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            pend = [enq(s_) for s_ in range(n_s)]
            for q in pend:
                q.result()
            a1.record()
            barrier()
            reps.append(a0.elapsed_time(a1))
        ms_1 = statistics.median(reps)
        stage_s = {}
        if timing:                      # stage breakdown from a separate short run (event records slow short objects down)
            L.cppf_timing_collect(timing, (C.c_float * len(stage_names))())
            est_s.timing = timing
            for s_ in range(20):
                est_s.enqueue_fused(pinned[s_ % n_obj][0], pinned[s_ % n_obj][1], seed=s_, max_cells=cells[s_ % n_obj]).result()
            acc = (C.c_float * len(stage_names))()
            calls = L.cppf_timing_collect(timing, acc)
            est_s.timing = None
            stage_s = {nm: round(acc[i] / max(calls, 1), 4) for i, nm in enumerate(stage_names)}
        sampled = {"objects_per_sec_per_gpu": n_s / (ms_1 * 1e-3), "pairs_per_sec_per_gpu": n_s * 100000 / (ms_1 * 1e-3),
                   "ms_per_object": ms_1 / n_s, "ms_per_object_repeats": [r / n_s for r in reps], "host_enqueue_us_per_object": statistics.median(host_us), "stage_ms": stage_s,
                   "pairs_per_object": 100000,
                   "note": "reference regime: P = 100 000 sampled pairs (nocs/inference.py:177), host clouds in, pose records out"}
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
        kern = dict(timers)
        # Algorithmic work per launch.  Dense pairs are enumerated in-kernel (no index read).
        #   encode_sample : 24 B/pair written (4 bin bytes + 5 tail floats); 23 968 FLOP/pair canonical pair MLP
        #                   (models/model.py:12-23,87; 13.9 k executed after the per-point pre-projection of layer 0)
        #   vote          : 4 B/pair read (bins); ~180 shared-memory atomics/pair under the trained-like load
        #   twopass       : first-pass encode writes 64 fp32 logits/pair, vote reads 8 B (mu,nu)/pair
        algo_bytes = {"encode_sample": pairs_per_obj * 24, "vote": pairs_per_obj * 4, "backvote": pairs_per_obj * 5,
                      "stats": pairs_per_obj * 24,
                      "ppf_encode_pass1": pairs_per_obj * 64 * 4, "ppf_vote": pairs_per_obj * 8}
        algo_flops = {"encode_sample": pairs_per_obj * 23968.0, "ppf_encode_pass1": pairs_per_obj * (23968.0 - 2 * 16 * 77)}
        binding = {"encode_sample": "SIMT epilogues between the tcgen05 MMA steps (3xTF32 chain; issue slots 46 %, tensor pipe 44 %)" if args.encoder == "tc"
                                    else "fp32 FMA pipe",
                   "vote": "shared-memory pipe (74 % of peak wavefronts, 3.6 bank/same-cell replays per ATOMS) and issue slots (77 %)",
                   "backvote": "fp32 / issue", "ppf_vote": "L2 atomic throughput", "stats": "HBM gather (2.9 TB/s)",
                   "ppf_encode_pass1": "fp32 FMA pipe", "point_encoder": "FP32/issue (two-sweep kNN select + one-warp-per-point SPRIN MLP)"}
        detail = {}
        for k, v in kern.items():
            t = v["avg_ms"] * 1e-3
            d = {"avg_ms": v["avg_ms"], "binding_resource": binding.get(k)}
            if k in BINDING_NCU:
                d["binding_utilisation_ncu"] = BINDING_NCU[k]
            if k in TRAFFIC_NCU:
                d["dram_bytes_per_launch_ncu"] = TRAFFIC_NCU[k]
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
            if k == "vote" and args.votes == "trained_like" and args.path == "fused" and args.n_points == 4096:
                # the vote phase is measured in atomics/s against a same-box microbenchmark
                # (tools/atomics_bench.cu -> profiles/r1_atomics_microbench.json); the count per launch is ncu's
                # predicated-on thread count of the ATOMS instructions of this workload (profiles/r1f_kernels_ncu.md)
                n_atom = 3.0246e9
                d["atomics"] = {"per_launch_ncu": n_atom, "achieved": n_atom / t / 1e9, "unit": "G atomics/s",
                                "peak_shared_u32_trilinear_pattern_microbench": 2813.4, "frac": n_atom / t / 1e9 / 2813.4,
                                "global_fp32_red_microbench": {"uniform_71KB": 93.9, "concentrated_71KB": 16.4},
                                "note": "shared-memory u32 atomics of the privatised grid; the microbenchmark peak is the bare "
                                        "8-corner splat pattern with nothing else in the loop"}
            detail[k] = d
        roof = None
        timed = [k for k in kern if k in algo_bytes]
        if timed:
            dom = max(timed, key=lambda k: kern[k]["avg_ms"])
            use_tensor = dom in algo_flops and args.encoder == "tc"
            src = detail[dom]["tensor" if use_tensor else "hbm"]
            roof = {"kernel": dom, "bound": "tensor" if use_tensor else "hbm", "achieved": src["achieved"], "peak": src["peak"],
                    "unit": src["unit"], "frac": src["frac"], "traffic": TRAFFIC_NCU.get(dom), "peak_source": peak_src,
                    "avg_launch_ms": kern[dom]["avg_ms"], "binding_resource": binding.get(dom),
                    "note": "dominant kernel by live CUDA-event time; none of this path's kernels is HBM-bound "
                            "(logits never reach HBM) -- see roofline_detail for every kernel against the resource "
                            "that binds it"}
        cpu = None
        if not args.no_cpu_baseline:
            r = run_cpu_reference(args, 2, 1, args.cpu_sample_pairs)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"],
                   "kind": "reference" if r["impl"] == "ref_cpu" else "port", "sample": r["sample"]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_hbm / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": cfgj,
                "objects_per_sec": world * args.steps / (ms_hbm * 1e-3),
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "objects_per_sec": world * args.steps / (ms_e2e * 1e-3), "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "roofline_detail": detail, "kernels": kern,
                "cpu_baseline": cpu}
        if ms_net is not None:
            line["variant_network_votes"] = {"value": total_pairs / (ms_net * 1e-3), "unit": UNIT,
                                             "ms_per_step": ms_net / args.steps,
                                             "note": "no bin injection: votes from the random-init network's own samples"}
        if sampled is not None:
            line["variant_sampled_100k"] = sampled
        print(json.dumps(line))
    if dist_on:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
