"""Indexed tcgen05 encoder: time against the number of pairs (prologue + rounds of 592 tiles)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cppf_b200 import model, synth, fast
dev = torch.device("cuda")
torch.manual_seed(0)
n = 4096
ppf = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).to(dev).eval()
pc, nrm = synth.synth_bottle(n, 1)
pc, nrm = torch.from_numpy(pc).to(dev), torch.from_numpy(nrm).to(dev)
feat = torch.randn(n, 40, device=dev)
table = ppf.tc_preproject(feat)
heads = 1 | 2 | 8
for P in (128, 128 * 148, 128 * 592, 100000, 128 * 592 * 2, 128 * 592 * 3, 128 * 592 * 4, 128 * 592 * 8, 128 * 592 * 16):
    idx = torch.randint(0, n, (P, 2), device=dev, dtype=torch.int32)
    bins = torch.empty((P, 4), dtype=torch.uint8, device=dev)
    tail = torch.empty((5, P), dtype=torch.float32, device=dev)
    best = 1e9
    for _ in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fast.encode_sample(ppf, pc, nrm, table, idx, heads=heads, seed=3, bins=bins, tail=tail)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"P {P:8d} tiles {-(-P // 128):6d} rounds {-(-P // 128) / 592:.2f}: {best * 1e3:.1f} us")
