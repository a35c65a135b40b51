"""Host cost of enqueue_batch on an idle GPU (the launch queue never fills): fixed cost per call and slope per object."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from cppf_b200 import model, synth
from cppf_b200.pipeline import PoseConfig, PoseEstimator, enqueue_batch
dev = torch.device("cuda")
torch.manual_seed(0)
pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(dev).eval()
ppf = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).to(dev).eval()
est = PoseEstimator(pe, ppf, PoseConfig.from_dict(dict(synth.BOTTLE, n_pairs=100000)), dev)
n = 4096
clouds = [synth.synth_bottle(n, 1000 + s) for s in range(24)]
items = [(est, clouds[s % 24][0], clouds[s % 24][1], s) for s in range(64)]
caps = [est.grid_capacity(it[1]) for it in items]
enqueue_batch(items[:32]).results(on_error="none")
torch.cuda.synchronize()
for ns, nt in ((1, 1), (4, 1), (4, 4), (8, 8)):
    res = {}
    for k in (8, 32):
        best = 1e9
        for rep in range(7):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            p = enqueue_batch(items[:k], n_streams=ns, n_threads=nt, capacities=caps[:k])
            t1 = time.perf_counter()
            p.records17()
            best = min(best, t1 - t0)
        res[k] = best * 1e6
    slope = (res[32] - res[8]) / 24
    print(f"streams {ns} threads {nt}: 8 objects {res[8]:.0f} us, 32 objects {res[32]:.0f} us -> {slope:.1f} us/object + {res[8] - 8 * slope:.0f} us/call")

# throughput of a long batch by (streams, threads)
big = [(est, clouds[s % 24][0], clouds[s % 24][1], s) for s in range(240)]
bcaps = [caps[s % 24] for s in range(240)]
for ns, nt in ((4, 1), (4, 2), (4, 4), (8, 1), (8, 2), (8, 8), (6, 1), (6, 3)):
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        enqueue_batch(big, n_streams=ns, n_threads=nt).records17()
        best = min(best, time.perf_counter() - t0)
    print(f"streams {ns} threads {nt}: {240 / best:.0f} objects/s")

# python-only cost: the library call stubbed out
from cppf_b200 import _lib
L = _lib.lib()
real = L.cppf_pose_batch
L.cppf_pose_batch = lambda *a: 0
for capmode in ("given", "derived"):
    res = {}
    for k in (8, 32):
        best = 1e9
        for rep in range(7):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            p = enqueue_batch(items[:k], n_streams=4, n_threads=4, capacities=caps[:k] if capmode == "given" else None)
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            best = min(best, t1 - t0)
        res[k] = best * 1e6
    slope = (res[32] - res[8]) / 24
    print(f"python only, caps {capmode}: {slope:.1f} us/object + {res[8] - 8 * slope:.0f} us/call")
L.cppf_pose_batch = real
