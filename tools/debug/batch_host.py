"""Where does the host time of enqueue_batch go?"""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from cppf_b200 import model, synth
from cppf_b200.pipeline import PoseConfig, PoseEstimator, enqueue_batch
dev = torch.device("cuda")
torch.manual_seed(0)
pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(dev).eval()
ppf = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).to(dev).eval()
est = PoseEstimator(pe, ppf, PoseConfig.from_dict(dict(synth.BOTTLE, n_pairs=100000)), dev)
n, k = 4096, 240
clouds = [synth.synth_bottle(n, 1000 + s) for s in range(24)]
items = [(est, clouds[s % 24][0], clouds[s % 24][1], s) for s in range(k)]
caps = [est.grid_capacity(it[1]) for it in items]
enqueue_batch(items[:16]).results(on_error="none")
torch.cuda.synchronize()
for ns, nt, cap in ((1, 1, None), (1, 1, caps), (4, 4, caps), (8, 8, caps)):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        p = enqueue_batch(items, n_streams=ns, n_threads=nt, capacities=cap)
        t1 = time.perf_counter()
        p.records17()
        t2 = time.perf_counter()
    print(f"streams {ns} threads {nt} caps={'given' if cap else 'derived'}: enqueue {1e6*(t1-t0)/k:.1f} us/object, total {1e3*(t2-t0)/k:.4f} ms/object")
# python loop of per-object calls
pend=[est.enqueue_fused(it[1], it[2], seed=it[3], device_pairs=True, max_cells=caps[i][0]) for i, it in enumerate(items[:8])]
[q.result() for q in pend]
torch.cuda.synchronize(); t0=time.perf_counter()
pend=[est.enqueue_fused(it[1], it[2], seed=it[3], device_pairs=True, max_cells=caps[i][0]) for i, it in enumerate(items)]
t1=time.perf_counter(); [q.result() for q in pend]; t2=time.perf_counter()
print(f"python loop: enqueue {1e6*(t1-t0)/k:.1f} us/object, total {1e3*(t2-t0)/k:.4f} ms/object")
pr = cProfile.Profile(); torch.cuda.synchronize(); pr.enable()
p = enqueue_batch(items[:48], n_streams=4, n_threads=4, capacities=caps[:48])
pr.disable(); p.records17()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
