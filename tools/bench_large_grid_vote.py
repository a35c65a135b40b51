#!/usr/bin/env python
"""Centre vote on a 64^3 grid (BASELINE config 3), N=4096 dense pairs: the three large-grid implementations side by side --
slab passes (cppf_vote_slabs), routed x-slabs (cppf_vote_routed), global fp32 reductions (cppf_ppf_vote) -- under the
trained-like and the random vote load.  The two shared-memory variants must produce bit-identical grids."""
import sys, json, numpy as np, torch
sys.path.insert(0,'/root/repo')
from cppf_b200 import fast, synth, voting, _lib
dev='cuda'
res=4e-3; n=4096
pc,_=synth.synth_cylinder_grid64(n,0,res=res)
corner,dims=synth.vote_grid_geometry(pc,res)
pcd=torch.from_numpy(pc).to(dev); cd=torch.from_numpy(corner).to(dev)
lut=fast.decode_lut(synth.BOTTLE['vote_range']).to(dev)
out={}
for tag in ('trained','random'):
    bins=torch.zeros(n*n,4,dtype=torch.uint8,device=dev)
    if tag=='trained': bins[:,:3]=synth.trained_like_bins_dense_torch(pcd,synth.BOTTLE)
    else: bins[:,:2]=torch.randint(0,32,(n*n,2),device=dev,dtype=torch.uint8)
    grids={}
    for name,fn in (('slabs',fast.vote_slabs),('routed',fast.vote_routed)):
        g=torch.zeros(dims,device=dev)
        fn(pcd,None,g,cd,res,bins=bins,lut=lut); torch.cuda.synchronize()
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            g.zero_(); fn(pcd,None,g,cd,res,bins=bins,lut=lut)
        b.record(); torch.cuda.synchronize()
        out[f'{tag}_{name}_ms']=a.elapsed_time(b)/3
        grids[name]=g
    b=bins.long()
    mu_nu=torch.stack([lut[b[:,0]],lut[32+b[:,1]]],-1).contiguous()
    g=torch.zeros(dims,device=dev)
    voting.ppf_vote(pcd,mu_nu,None,g,cd,res,72,True); torch.cuda.synchronize()
    a,b2=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); g.zero_(); voting.ppf_vote(pcd,mu_nu,None,g,cd,res,72,True); b2.record(); torch.cuda.synchronize()
    out[f'{tag}_global_fp32_ms']=a.elapsed_time(b2)
    out[f'{tag}_global_argmax_equal']=bool(int(voting.grid_argmax(g).item())==int(voting.grid_argmax(grids['slabs']).item()))
    out[f'{tag}_equal']=bool(torch.equal(grids['slabs'],grids['routed']))
    out[f'{tag}_sum']=float(grids['slabs'].sum())
print(json.dumps(out))
