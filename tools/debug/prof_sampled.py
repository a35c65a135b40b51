"""Profile aid: a few objects of the reference's regime (P = 100 000 sampled pairs, N = 4096) through cppf_pose_fused."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cppf_b200 import model, synth
from cppf_b200.pipeline import PoseConfig, PoseEstimator
dev = torch.device("cuda")
torch.manual_seed(0)
pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(dev).eval()
ppf = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).to(dev).eval()
est = PoseEstimator(pe, ppf, PoseConfig.from_dict(dict(synth.BOTTLE, n_pairs=100000)), dev)
for s in range(4):
    pc, nrm = synth.synth_bottle(4096, 1000 + s)
    try:
        est.enqueue_fused(torch.from_numpy(pc).pin_memory(), torch.from_numpy(nrm).pin_memory(), seed=s, device_pairs=True).result()
    except RuntimeError as e:
        print("object", s, e)
torch.cuda.synchronize()
print("done")
