"""GPU parity of the fused point encoder (csrc/point_encoder.cu): exact kNN selection and the SPRIN
convolution of models/model.py:46-77 + models/sprin.py:40-107, against fixtures minted from the
reference modules (tests/golden/encoder_bottle.npz) and against the oracle's restatement (oracle/ref_model.py)."""
import os

import numpy as np
import pytest
import torch

from cppf_b200 import model, synth
from oracle import ref_model

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(__file__), "golden", "encoder_bottle.npz")


def _golden_encoder():
    d = np.load(GOLD)
    pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32)
    pe.load_state_dict({k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("pe/")})
    return d, pe.to(DEV).eval()


def test_fused_forward_nbrs_matches_reference_fixture():
    d, pe = _golden_encoder()
    pc, nrm = torch.from_numpy(d["pc"]).to(DEV), torch.from_numpy(d["nrm"]).to(DEV)
    nbrs = torch.from_numpy(d["nbrs"]).to(DEV)
    with torch.no_grad():
        got = pe.forward_nbrs(pc[None], nrm[None], nbrs[None])[0]
    np.testing.assert_allclose(got.cpu().numpy(), d["feat_nbrs"], rtol=1e-4, atol=2e-5)
    # the drop-in forward(pc, normal, dist): torch.topk of the caller's matrix picks the neighbours (models/model.py:47)
    dist = torch.cdist(pc[None], pc[None])
    with torch.no_grad():
        got2 = pe(pc[None], nrm[None], dist)[0]
    np.testing.assert_allclose(got2.cpu().numpy(), d["feat"], rtol=1e-4, atol=5e-5)


@pytest.mark.parametrize("n,k,seed", [(2048, 60, 0), (333, 20, 1), (100, 64, 2), (70, 33, 3)])
def test_fused_encoder_equals_oracle_restatement(n, k, seed):
    torch.manual_seed(seed)
    pe = model.PointEncoder(k=k, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(DEV).eval()
    with torch.no_grad():
        for p in pe.parameters():
            p.mul_(1.5)
        for m in pe.modules():
            if isinstance(m, torch.nn.LayerNorm):
                m.weight.add_(0.3 * torch.randn_like(m.weight))
                m.bias.add_(0.3 * torch.randn_like(m.bias))
    pc, nrm = synth.synth_bottle(n, seed)
    pc, nrm = torch.from_numpy(pc).to(DEV), torch.from_numpy(nrm).to(DEV)
    nbrs = pe.knn(pc)
    with torch.no_grad():
        fused = pe.forward_nbrs(pc[None], nrm[None], nbrs[None])[0]
    sd = {kk: v.cpu() for kk, v in pe.state_dict().items()}
    ref = ref_model.point_encode_nbrs(pc.cpu(), nrm.cpu(), nbrs.cpu(), sd)      # models/model.py:63-77 restated (oracle)
    np.testing.assert_allclose(fused.cpu().numpy(), ref.numpy(), rtol=2e-4, atol=5e-5)
    assert torch.equal(fused[:, 32:], fused[:1, 32:].expand(n, 8))              # the global-max columns are shared


@pytest.mark.parametrize("n,k,seed", [(4096, 60, 0), (1000, 60, 1), (65, 64, 2), (33, 1, 3)])
def test_knn_selects_the_k_smallest_exact_distances(n, k, seed):
    pc, _ = synth.synth_bottle(n, seed)
    t = torch.from_numpy(pc).to(DEV)
    pe = model.PointEncoder(k=k, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(DEV)
    idx = pe.knn(t)
    assert idx.shape == (n, k) and int(idx.min()) >= 0 and int(idx.max()) < n
    d2 = ((t[:, None, :].double() - t[None, :, :].double()) ** 2).sum(-1)       # fp64 ground truth
    got = torch.sort(torch.gather(d2, 1, idx), dim=1)[0]
    want = torch.topk(d2, k, dim=1, largest=False, sorted=True)[0]
    # same multiset of distances up to fp32 rounding of d^2 at the k-th boundary
    np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-5, atol=1e-12)
    assert all(len(set(r)) == k for r in idx.cpu().numpy()[:: max(1, n // 50)])    # no duplicates within a row
    assert bool((idx == torch.arange(n, device=DEV)[:, None]).any(1).all())     # self is always a neighbour (d = 0)


def test_knn_ties_resolve_to_lowest_indices():
    pts = np.zeros((40, 3), np.float32)
    pts[:, 0] = np.repeat(np.arange(10), 4)                                      # 10 clusters of 4 identical points
    t = torch.from_numpy(pts).to(DEV)
    pe = model.PointEncoder(k=6, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(DEV)
    idx = torch.sort(pe.knn(t), dim=1)[0].cpu().numpy()
    # query 0 (cluster 0): its 4 duplicates at d=0, then the two lowest-index members of cluster 1 at d=1
    assert idx[0].tolist() == [0, 1, 2, 3, 4, 5]
    # query 20 (cluster 5): 4 at d=0; clusters 4 and 6 tie at d=1 -> lowest indices 16, 17
    assert idx[20].tolist() == [16, 17, 20, 21, 22, 23]


def test_knn_boundary_bin_overflow_falls_back_to_radix_select():
    """500 points at exactly the same distance from a small cluster: the k-th neighbour's coarse-histogram bin holds more
    points than the shared-memory candidate list, so the query takes the full radix-select path -- same answer (ties to
    the lowest indices)."""
    pts = np.zeros((510, 3), np.float32)
    ang = np.linspace(0, 2 * np.pi, 500, endpoint=False)
    pts[10:, 0], pts[10:, 1] = np.cos(ang), np.sin(ang)          # |p| = 1 up to rounding -> nearly identical d^2 from the origin
    pts[10:] = np.float32([1, 0, 0])                              # make them exactly identical: 500 coincident points
    t = torch.from_numpy(pts).to(DEV)
    pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(DEV)
    idx = torch.sort(pe.knn(t), dim=1)[0].cpu().numpy()
    assert idx[0].tolist() == list(range(60))                    # 10 at d=0, then the 50 lowest-index coincident points
    assert idx[300].tolist() == list(range(10, 70))              # a coincident point: 60 lowest indices among its 500 twins


@pytest.mark.parametrize("n,k,seed", [(4096, 60, 0), (1001, 60, 1), (257, 64, 2), (90, 7, 3), (1, 1, 4)])
def test_tensor_core_kernel_equals_ffma_kernel(n, k, seed, monkeypatch):
    """point_encode_tc_kernel (tcgen05, 3xTF32, two points per 128-row tile) against point_encode_kernel (fp32 FFMA, one
    warp per point) on the same neighbour lists: ragged last tile (odd n), k = 64 (no padding rows), k < 32, one point."""
    torch.manual_seed(seed)
    pe = model.PointEncoder(k=k, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).to(DEV).eval()
    with torch.no_grad():
        for p in pe.parameters():
            p.mul_(1.3)
    pc, nrm = synth.synth_bottle(max(n, 2), seed)
    pc, nrm = torch.from_numpy(pc[:n]).to(DEV), torch.from_numpy(nrm[:n]).to(DEV)
    nbrs = pe.knn(pc)
    with torch.no_grad():
        monkeypatch.setenv("CPPF_PE_IMPL", "simt")
        ffma = pe.encode_fused(pc, nrm, nbrs)
        monkeypatch.setenv("CPPF_PE_IMPL", "tc")
        tc = pe.encode_fused(pc, nrm, nbrs)
        tc2 = pe.encode_fused(pc, nrm, nbrs)
    assert torch.equal(tc, tc2)                                                  # deterministic (the global max is order-free)
    np.testing.assert_allclose(tc.cpu().numpy(), ffma.cpu().numpy(), rtol=2e-5, atol=2e-5)
