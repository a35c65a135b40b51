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
n, k = 4096, 120
clouds = [synth.synth_bottle(n, 1000 + s) for s in range(24)]
pinned = [(torch.from_numpy(p).pin_memory(), torch.from_numpy(q).pin_memory()) for p, q in clouds]
cells = [int(np.prod(synth.vote_grid_geometry(p, 4e-3)[1])) for p, _ in clouds]
def loop(tag):
    for s in range(3): est.enqueue_fused(pinned[s][0], pinned[s][1], seed=s, max_cells=cells[s]).result()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pend = [est.enqueue_fused(pinned[s % 24][0], pinned[s % 24][1], seed=s, max_cells=cells[s % 24]) for s in range(k)]
    t1 = time.perf_counter(); [q.result() for q in pend]; t2 = time.perf_counter()
    print(f"{tag}: enqueue {1e6*(t1-t0)/k:.1f} us/object, total {1e3*(t2-t0)/k:.4f} ms/object", flush=True)
loop("before any batch")
items = [(est, clouds[s % 24][0], clouds[s % 24][1], s) for s in range(k)]
enqueue_batch(items, n_streams=1, n_threads=1).records17(); loop("after batch(1,1)")
enqueue_batch(items, n_streams=4, n_threads=1).records17(); loop("after batch(4,1)")
enqueue_batch(items, n_streams=4, n_threads=4).records17(); loop("after batch(4,4)")
pr = cProfile.Profile(); torch.cuda.synchronize(); pr.enable()
pend = [est.enqueue_fused(pinned[s % 24][0], pinned[s % 24][1], seed=s, max_cells=cells[s % 24]) for s in range(k)]
pr.disable(); [q.result() for q in pend]
pstats.Stats(pr).sort_stats("tottime").print_stats(6)
