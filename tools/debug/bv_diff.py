"""Debug aid: find the pairs on which backvote_bins differs from the reference cubin (dense N=1024 case of the test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
from cppf_b200 import fast, synth, voting
from oracle import ref_gpu, clib
from parity_util import backvote_decision_margin
DEV = "cuda"
def _t(a, dt=torch.float32): return torch.from_numpy(np.ascontiguousarray(a)).to(DEV, dt)
cfg = synth.BOTTLE
pc, _ = synth.synth_bottle(1024, 8)
idxs = synth.dense_pairs(1024)
tr = synth.trained_like_tr(pc, idxs)
corner, dims = synth.vote_grid_geometry(pc, cfg["res"])
lut = fast.decode_lut(cfg["vote_range"])
b_mu = torch.argmin((torch.from_numpy(tr[:, 0:1]) - lut[None, :32]).abs(), -1)
b_nu = torch.argmin((torch.from_numpy(tr[:, 1:2]) - lut[None, 32:64]).abs(), -1)
bins = torch.stack([b_mu, b_nu, b_mu * 0, b_mu * 0], -1).to(torch.uint8).to(DEV)
grid = torch.zeros(dims, device=DEV)
fast.vote_fast(_t(pc), None, grid, _t(corner), cfg["res"], bins=bins, lut=lut.to(DEV))
flat = voting.grid_argmax(grid)
gyz = dims[1] * dims[2]
for shift in (0, 4 * gyz + 3 * dims[2] + 2, -(3 * gyz) - 5 * dims[2], 2 * gyz - 7 * dims[2] + 1, 5):
    f2 = (flat + shift).clamp(0, dims[0] * gyz - 1)
    mask = fast.backvote_bins(_t(pc), bins, lut.to(DEV), None, dims, _t(corner), f2, cfg["res"], 3 * cfg["res"])
    cell = np.array(np.unravel_index(int(f2.item()), dims))
    centre = (np.asarray(corner, np.float64) + cell * cfg["res"]).astype(np.float32)
    ref_off = ref_gpu.backvote(_t(pc), _t(tr), torch.zeros(idxs.shape[0], 3, device=DEV), _t(idxs, torch.int32), _t(corner),
                               cfg["res"], 72, dims, _t(centre), 3 * cfg["res"])
    ref_mask = (ref_off != 0).any(-1)
    # the literal (mode M) kernel of this library on the same (mu, nu) floats
    off2, mask2 = voting.backvote(_t(pc), _t(tr), _t(idxs, torch.int32), dims, _t(corner), cfg["res"], _t(centre), 3 * cfg["res"])
    d = torch.nonzero(mask.bool() != ref_mask)[:, 0].cpu().numpy()
    d2 = torch.nonzero(mask2.bool() != ref_mask)[:, 0].cpu().numpy()
    print("cell", cell.tolist(), "survivors", int(ref_mask.sum()), "bins-kernel diffs", d.tolist(), "literal-kernel diffs", d2.tolist())
    for p in d:
        a, b = idxs[p]
        m = backvote_decision_margin(pc, tr, idxs, dims, corner, cfg["res"], centre, np.float32(3 * cfg["res"]), rows=[p])
        ab = pc[a].astype(np.float64) - pc[b].astype(np.float64)
        abn = ab / np.linalg.norm(ab)
        nu = float(tr[p, 1]); mu = float(tr[p, 0])
        n = min(int(np.float32(nu) / np.float32(cfg["res"]) * (2 * np.pi)), 72)
        print("  pair", p, "a,b", a, b, "mu,nu", mu, nu, "n", n, "ab", abn.tolist(), "ours", int(mask[p]), "ref", int(ref_mask[p]),
              "margin_m", m.tolist(), "ref_off", ref_off[p].cpu().numpy().tolist())
        # float64 candidate distances
        c = pc[a].astype(np.float64) - abn * mu
        co = np.array([0.0, -abn[2], abn[1]])
        x = co / np.linalg.norm(co) * nu
        y = np.cross(x, abn)
        ang = (np.arange(n) * 2 * np.pi / n)
        cand = c[None] + np.cos(ang)[:, None] * x[None] + np.sin(ang)[:, None] * y[None]
        dist = np.linalg.norm(cand - centre.astype(np.float64)[None], axis=1)
        order = np.argsort(dist)[:4]
        print("   closest candidates i", order.tolist(), "dist - tol", (dist[order] - np.float32(3 * cfg["res"])).tolist(),
              "unit frame ex.ex", float(np.dot(co / (np.linalg.norm(co) + 1e-7), co / (np.linalg.norm(co) + 1e-7))))
