"""Profile aid: one routed 64^3 vote (trained-like) for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cppf_b200 import fast, synth
dev='cuda'; res=4e-3; n=4096
pc,_=synth.synth_cylinder_grid64(n,0,res=res)
corner,dims=synth.vote_grid_geometry(pc,res)
pcd=torch.from_numpy(pc).to(dev); cd=torch.from_numpy(corner).to(dev)
lut=fast.decode_lut(synth.BOTTLE['vote_range']).to(dev)
bins=torch.zeros(n*n,4,dtype=torch.uint8,device=dev)
bins[:,:3]=synth.trained_like_bins_dense_torch(pcd,synth.BOTTLE)
for _ in range(2):
    g=torch.zeros(dims,device=dev)
    fast.vote_routed(pcd,None,g,cd,res,bins=bins,lut=lut)
torch.cuda.synchronize()
a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
a.record(); g.zero_(); fast.vote_routed(pcd,None,g,cd,res,bins=bins,lut=lut); b.record(); torch.cuda.synchronize()
print("routed ms", a.elapsed_time(b), "sum", float(g.sum()))
