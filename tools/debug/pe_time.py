"""Point encoder (kNN + SPRIN) timing, N = 4096, k = 60; CPPF_PE_IMPL=simt for the FFMA kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cppf_b200 import model, synth
dev = torch.device("cuda")
torch.set_grad_enabled(False)
torch.manual_seed(0)
pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(dev).eval()
for n in (4096, 1024, 16384):
    pc, nrm = synth.synth_bottle(n, 1)
    pc, nrm = torch.from_numpy(pc).to(dev), torch.from_numpy(nrm).to(dev)
    nbrs = pe.knn(pc)
    f = pe.encode_fused(pc, nrm, nbrs)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f = pe.encode_fused(pc, nrm, nbrs); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"impl {os.environ.get('CPPF_PE_IMPL', 'tc')} n {n}: encode {best * 1e3:.1f} us  checksum {float(f.double().abs().sum()):.6f}")
