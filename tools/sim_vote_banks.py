#!/usr/bin/env python
"""Offline simulation of the shared-memory bank behaviour of the centre-vote splat (csrc/vote_private.cu): the
candidate stream of a CTA batch is regenerated in numpy in the kernel's order (pairs sorted by rotation count, 32-pair
chunks, in-bounds candidates compacted into groups of 32) and the wavefronts per ATOMS instruction are counted as the
maximum number of lanes per bank (same-address lanes serialise too).  It reproduces the ncu figure (4.9 wavefronts per
ATOMS, profiles/r1c_vote_sorted_ncu.md) and was used to evaluate alternatives before spending GPU time: padded strides
(no gain: the replays are same-CELL collisions, a quarter of the lanes of a splat share their base cell), a hot box
around the peak (4.4-4.6), two / four grid replicas by lane parity (4.0 / 3.7; two fit in shared memory and were
adopted: vote 2.92 -> 2.65 ms).  `--tile A B` regenerates the batches as A x B tiles of the dense pair matrix, the
layout the kernel uses in dense mode (128 x 16: 3.65 with two replicas, ncu 3.64).  Pure numpy, no GPU."""
import numpy as np, sys
TILE = None
if '--tile' in sys.argv:
    k = sys.argv.index('--tile')
    TILE = (int(sys.argv[k + 1]), int(sys.argv[k + 2]))
sys.path.insert(0,'/root/repo')
from cppf_b200 import synth
n=1024
pc,_=synth.synth_bottle(n,0)
res=np.float32(4e-3)
corner,dims=synth.vote_grid_geometry(pc,float(res))
gx,gy,gz=dims
rng=np.random.default_rng(0)
# emulate one CTA batch: 2048 consecutive dense pairs (same a mostly), sorted by n desc, chunks of 32
def batch_candidates(a, b0):
    if TILE is None:
        b=np.arange(b0,b0+2048)%n
        idx=np.stack([np.full(2048,a),b],-1)
    else:                                   # local index = ai * B + bi, as in vote_private_kernel's pair_index
        A,B=TILE
        ai=(a+np.arange(A))%n; bi=(b0+np.arange(B))%n
        idx=np.stack([np.repeat(ai,B),np.tile(bi,A)],-1)
    tr=synth.trained_like_tr(pc,idx)
    mu,nu=tr[:,0],tr[:,1]
    nrot=np.minimum((nu.astype(np.float64)/res*2*np.pi).astype(int),72)
    order=np.argsort(-nrot,kind='stable')
    return idx[order],mu[order],nu[order],nrot[order]
def frames(idx,mu,nu):
    A=pc[idx[:,0]].astype(np.float64); B=pc[idx[:,1]].astype(np.float64)
    ab=A-B; ln=np.linalg.norm(ab,axis=1); ok=ln>1e-7
    ab=ab/(ln[:,None]+1e-7)
    co=np.stack([np.zeros(len(ab)),-ab[:,2],ab[:,1]],-1)
    lc=np.linalg.norm(co,axis=1)
    ex=co/(lc[:,None]+1e-7)
    c=A-ab*mu[:,None]; x=ex*nu[:,None]; y=np.cross(x,ab)
    return c,x,y,ok
def wavefronts(flat_groups, strides):
    # flat_groups: list of arrays of base coords (fx,fy,fz) per 32-group
    sx,sy=strides
    tot=0; cnt=0
    for g in flat_groups:
        fx,fy,fz=g[:,0],g[:,1],g[:,2]
        base=fx*sx+fy*sy+fz
        for off in (0,1,sy,sy+1,sx,sx+1,sx+sy,sx+sy+1):
            addr=base+off
            bank=addr%32
            # wavefronts = max over banks of number of lanes (same address also serialises)
            tot+=np.bincount(bank,minlength=32).max(); cnt+=1
    return tot/cnt
groups=[]
for a in range(0,12):
    idx,mu,nu,nrot=batch_candidates(a*37%n, (a*211)%n)
    c,x,y,ok=frames(idx,mu,nu)
    for w in range(0,len(idx),32):
        sl=slice(w,w+32)
        nm=nrot[sl].max()
        q=[]
        for i in range(nm):
            nn=np.maximum(nrot[sl],1)
            ang=(i*2*np.pi/nn).astype(np.float32)
            cand=c[sl]+np.cos(ang)[:,None]*x[sl]+np.sin(ang)[:,None]*y[sl]
            g=(cand-corner)/res
            inb=(i<nrot[sl])&ok[sl]&np.all(g>=0.01,1)&(g[:,0]<gx-1.01)&(g[:,1]<gy-1.01)&(g[:,2]<gz-1.01)
            q.extend(list(np.floor(g[inb]).astype(int)))
            while len(q)>=32:
                groups.append(np.array(q[:32])); q=q[32:]
print(len(groups),'groups')
print('current strides',(gy*gz,gz), wavefronts(groups,(gy*gz,gz)))
for sy in (17,19,21,23,33):
    for sx in (gy*sy, gy*sy+1, gy*sy+3):
        print('sy',sy,'sx',sx,'mod32',(sx%32,sy%32), round(wavefronts(groups,(sx,sy)),3))
# random baseline
rg=[np.stack([rng.integers(0,gx-1,32),rng.integers(0,gy-1,32),rng.integers(0,gz-1,32)],-1) for _ in range(2000)]
print('uniform random cells', wavefronts(rg,(gy*gz,gz)))
# decompose: same-address multiplicity
import collections
sx,sy=gy*gz,gz
mx_addr=[];mx_bank=[]; dup_frac=[]
for g in groups[:6000]:
    base=g[:,0]*sx+g[:,1]*sy+g[:,2]
    u,c=np.unique(base,return_counts=True)
    mx_addr.append(c.max()); dup_frac.append(1-len(u)/32)
    mx_bank.append(np.bincount(base%32,minlength=32).max())
print('mean max same-address', np.mean(mx_addr), 'mean max bank', np.mean(mx_bank), 'dup frac', np.mean(dup_frac))
print('hist max same-addr', np.bincount(mx_addr)[:12])
# if same-address lanes were combined (counted once): wavefronts
def wf_combined(groups):
    tot=0;cnt=0
    for g in groups[:6000]:
        base=np.unique(g[:,0]*sx+g[:,1]*sy+g[:,2])
        for off in (0,1,sy,sy+1,sx,sx+1,sx+sy,sx+sy+1):
            tot+=np.bincount((base+off)%32,minlength=32).max(); cnt+=1
    return tot/cnt
print('wavefronts if identical base cells were merged', wf_combined(groups))
# distance of candidates from centre cell
cen=np.array([( -corner[0])/res, (-corner[1])/res, (-corner[2])/res])
allc=np.concatenate(groups[:6000])
d=np.abs(allc-np.floor(cen)).max(1)
print('frac within 1 cell of centre', np.mean(d<=1), 'within 2', np.mean(d<=2))
def wf_replicas(groups, R, off_mod):
    tot=0;cnt=0
    cells=gx*gy*gz
    rep_off=cells + ((off_mod - cells) % 32)
    for g in groups[:6000]:
        lane=np.arange(len(g))
        base=g[:,0]*sx+g[:,1]*sy+g[:,2] + (lane%R)*rep_off
        for off in (0,1,sy,sy+1,sx,sx+1,sx+sy,sx+sy+1):
            addr=base+off
            # wavefront count: lanes with same bank serialize (same address too)
            tot+=np.bincount(addr%32,minlength=32).max(); cnt+=1
    return tot/cnt
for R in (1,2,4):
    for om in (0,8,16,5):
        print('replicas',R,'offset mod32',om, round(wf_replicas(groups,R,om),3))
def wf_hotbox(groups, R, half, box):
    tot=0;cnt=0
    cells=gx*gy*gz
    c0=np.floor(cen).astype(int)-half
    hot_base=cells+ (0 - cells)%32
    bs=box
    for g in groups[:6000]:
        lane=np.arange(len(g))
        loc=g-c0
        hot=np.all((loc>=0)&(loc<bs-1),1)
        for (dx,dy,dz) in [(a,b,c) for a in (0,1) for b in (0,1) for c in (0,1)]:
            main=(g[:,0]+dx)*sx+(g[:,1]+dy)*sy+g[:,2]+dz
            hb=hot_base+(lane%R)*(bs**3 + 1) + (loc[:,0]+dx)*bs*bs+(loc[:,1]+dy)*bs+loc[:,2]+dz
            addr=np.where(hot,hb,main)
            tot+=np.bincount(addr%32,minlength=32).max(); cnt+=1
    return tot/cnt, 
for R,half,box in ((4,2,6),(8,2,6),(8,3,8),(16,3,8),(32,2,6),(32,3,8)):
    print('hotbox R',R,'half',half,'box',box, wf_hotbox(groups,R,half,box), 'smem KB', R*(box**3+1)*4/1024)
