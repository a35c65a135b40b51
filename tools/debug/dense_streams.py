"""Does fanning dense N=4096 objects over 2-3 worker streams (cppf_pose_batch) hide the small kernels of one object behind
the big kernels of the next?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from cppf_b200 import model, synth
from cppf_b200.pipeline import PoseConfig, PoseEstimator, enqueue_batch
dev = torch.device("cuda")
torch.manual_seed(0)
pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(dev).eval()
ppf = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).to(dev).eval()
est = PoseEstimator(pe, ppf, PoseConfig.from_dict(dict(synth.BOTTLE, n_pairs=0)), dev)
n, k = 4096, 24
clouds = [synth.synth_bottle(n, 1000 + s) for s in range(k)]
dcl = [(torch.from_numpy(p).to(dev), torch.from_numpy(q).to(dev)) for p, q in clouds]
inj = [synth.trained_like_bins_dense_torch(p, synth.BOTTLE) for p, _ in dcl]
items = [(est, p, q, s) for s, (p, q) in enumerate(dcl)]
for ns, nt in ((1, 1), (2, 2), (3, 3), (4, 4)):
    enqueue_batch(items[:ns * 2], n_streams=ns, n_threads=nt, inject_bins=inj[:ns * 2]).results()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        enqueue_batch(items, n_streams=ns, n_threads=nt, inject_bins=inj).results()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / k)
    print(f"streams {ns} threads {nt}: {best:.4f} ms/object  {n*n/best*1e-6:.3f} Gpairs/s")
