"""CPU: the packed weight blob (cppf_b200/csrc/mlp_layout.h, model.pack_ppf_weights) is
interpreted here in numpy exactly the way the kernels walk it -- pre-projection, PPF
columns, permuted k-major matrices, 48-wide final chunks -- and must reproduce the
reference logits of the golden fixture."""
import numpy as np
import torch

from conftest import load_golden, split_state
from cppf_b200 import model


def _unperm(m, no):
    k = m.shape[0]
    return m.reshape(k, 8, no).transpose(0, 2, 1).reshape(k, 8 * no)     # stored og*no+c -> logical og+8c


def test_blob_interpreter_reproduces_reference_logits():
    d = load_golden("encoder_bottle.npz")
    sd = split_state(d, "ppf/")
    out_dim = 141
    blob = model.pack_ppf_weights(sd, out_dim)
    F = 40
    o = 0
    take = lambda n, shape: (blob[o:o + n].reshape(shape))
    pre_wa = blob[0:F * 64].reshape(F, 64)
    pre_wb = blob[F * 64:2 * F * 64].reshape(F, 64)
    pre_b = blob[2 * F * 64:2 * F * 64 + 64]
    p = blob[2 * F * 64 + 64:]
    sec = {}
    off = 0
    for name, n in [("wppf", 256), ("w2_0", 1024), ("w1_1", 1024), ("b1_1", 32), ("w2_1", 1024), ("b2_1", 32),
                    ("w10_2", 1024), ("b10_2", 32), ("w2_2", 256)]:
        sec[name] = p[off:off + n]
        off += n
    outp = (out_dim + 47) // 48 * 48
    wf = p[off:off + 16 * outp].reshape(16, outp)
    bf = p[off + 16 * outp:off + 17 * outp]
    assert off + 17 * outp == p.size

    pc, nrm, feat, idxs = d["pc"], d["nrm"], d["feat"], d["idxs"]
    table = np.concatenate([feat @ pre_wa + pre_b, feat @ pre_wb], 1)        # [N,128]
    a, b = idxs[:, 0], idxs[:, 1]
    dv = pc[a] - pc[b]
    dn = np.sqrt((dv * dv).sum(-1)).astype(np.float32)
    dh = dv / (dn[:, None] + np.float32(1e-7))
    ppf = np.stack([(nrm[a] * dh).sum(-1), (nrm[b] * dh).sum(-1), (nrm[a] * nrm[b]).sum(-1), dn], -1)
    l0 = table[a, :64] + table[b, 64:] + ppf @ sec["wppf"].reshape(4, 64)
    h, r = np.maximum(l0[:, :32], 0), l0[:, 32:]
    x1 = h @ _unperm(sec["w2_0"].reshape(32, 32), 4) + r
    u = np.maximum(x1 @ _unperm(sec["w1_1"].reshape(32, 32), 4) + _unperm(sec["b1_1"][None], 4), 0)
    x2 = u @ _unperm(sec["w2_1"].reshape(32, 32), 4) + _unperm(sec["b2_1"][None], 4) + x1
    ur = x2 @ _unperm(sec["w10_2"].reshape(32, 32), 4) + _unperm(sec["b10_2"][None], 4)
    x3 = np.maximum(ur[:, :16], 0) @ _unperm(sec["w2_2"].reshape(16, 16), 2) + ur[:, 16:]
    unperm_f = lambda m: m.reshape(m.shape[0], outp // 48, 8, 6).transpose(0, 1, 3, 2).reshape(m.shape[0], outp)
    logits = (x3 @ unperm_f(wf) + unperm_f(bf[None]))[:, :out_dim]
    np.testing.assert_allclose(logits, d["logits"], rtol=2e-4, atol=2e-5)


def test_state_dict_keys_match_reference_checkpoints():
    d = load_golden("encoder_bottle.npz")
    ppf = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141)
    pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32)
    assert set(ppf.state_dict()) == set(split_state(d, "ppf/"))
    assert set(pe.state_dict()) == set(split_state(d, "pe/"))
    ppf.load_state_dict(split_state(d, "ppf/"))
    pe.load_state_dict(split_state(d, "pe/"))
    for k, v in pe.state_dict().items():
        assert tuple(v.shape) == d["pe/" + k].shape


def test_point_encoder_has_no_eager_path():
    """The product PointEncoder runs only through its sm_100a kernel: CPU tensors (or a batch, or an unsupported
    configuration) raise -- the torch-op composition lives in the oracle (oracle/ref_model.py:point_encode_nbrs, pinned to
    the reference module by test_oracle_golden.py::test_point_encoder_matches_reference)."""
    import pytest
    d = load_golden("encoder_bottle.npz")
    pe = model.PointEncoder(k=60, spfcs=[32, 64, 32, 32], num_layers=1, out_dim=32).eval()
    pe.load_state_dict(split_state(d, "pe/"))
    with torch.no_grad(), pytest.raises(NotImplementedError):
        pe.forward_nbrs(torch.from_numpy(d["pc"])[None], torch.from_numpy(d["nrm"])[None], torch.from_numpy(d["nbrs"])[None])


def test_unsupported_architecture_fails_loudly():
    import pytest
    m = model.PPFEncoder(ppffcs=[84, 64, 32, 16], out_dim=141)
    with pytest.raises(NotImplementedError):
        m.weight_blob(torch.device("cpu"))


# ---------------------------------------------------------------------------------------
# tcgen05 encoder blob (cppf_b200/model.py:pack_tc_weights, csrc/encode_tc.cu "chain algebra"): emulate the
# kernel's 4 steps in numpy straight from the packed blob and compare with the oracle's layer-by-layer stack.
def _uncanon(block, n, k):
    """[K/4][N][4] -> [N, K]"""
    return block.reshape(k // 4, n, 4).transpose(1, 0, 2).reshape(n, k)


def _tc_operand(blob, off, n, k):
    hi = _uncanon(blob[off:off + n * k], n, k)
    lo = _uncanon(blob[off + n * k:off + 2 * n * k], n, k)
    assert not (hi.view(np.uint32) & 0x1FFF).any()                 # hi is exactly representable in tf32
    return hi.astype(np.float64) + lo.astype(np.float64)


def test_tc_blob_chain_algebra_matches_oracle():
    import torch

    from cppf_b200 import model
    from oracle import ref_model

    torch.manual_seed(3)
    m = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141)
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(1.7)
    sd = m.state_dict()
    blob = model.pack_tc_weights(sd)
    rng = np.random.default_rng(0)
    n, p = 50, 400
    feat = rng.standard_normal((n, 40)).astype(np.float32)
    ppf = rng.standard_normal((p, 4)).astype(np.float32)
    ia, ib = rng.integers(0, n, p), rng.integers(0, n, p)
    TAB, OFF_S = 192, 40 * 192 + 192
    pre_w = blob[:40 * TAB].reshape(40, TAB).astype(np.float64)
    pre_b = blob[40 * TAB:40 * TAB + TAB].astype(np.float64)
    table = feat.astype(np.float64) @ pre_w + pre_b
    ta, tb = table[ia, :96], table[ib, 96:]
    o = OFF_S
    wp = _tc_operand(blob, o, 96, 8); o += 2 * 96 * 8
    ws1 = _tc_operand(blob, o, 64, 32); o += 2 * 64 * 32
    ws2 = _tc_operand(blob, o, 32, 32); o += 2 * 32 * 32
    wh = _tc_operand(blob, o, 112, 32); o += 2 * 112 * 32
    wr = _tc_operand(blob, o, 48, 32); o += 2 * 48 * 32
    bh = blob[o:o + 112 * 8].reshape(2, 112, 4).astype(np.float64); o += 112 * 8      # k = 0: hi, k = 4: lo
    br = blob[o:o + 48 * 8].reshape(2, 48, 4).astype(np.float64); o += 48 * 8
    assert o == blob.size and not bh[:, :, 1:].any() and not br[:, :, 1:].any()
    assert not (blob[o - 48 * 8 - 112 * 8:o - 48 * 8 - 112 * 4].view(np.uint32) & 0x1FFF).any()     # hi plane is tf32-exact
    bias = np.concatenate([bh[0, :, 0] + bh[1, :, 0], br[0, :, 0] + br[1, :, 0]])
    assert not wp[:, 4:].any()                                      # K padding of the ppf operand
    d = np.concatenate([ppf, np.zeros((p, 4), np.float32)], 1).astype(np.float64) @ wp.T          # step 0
    h = np.maximum(d[:, :32] + ta[:, :32] + tb[:, :32], 0)
    d[:, 32:96] += h @ ws1.T                                                                      # step 1
    u = np.maximum(d[:, 32:64] + ta[:, 32:64] + tb[:, 32:64], 0)
    d[:, 64:96] += u @ ws2.T                                                                      # step 2
    t = d[:, 64:96] + ta[:, 64:96] + tb[:, 64:96]
    a3 = np.concatenate([np.maximum(t[:, :16], 0), t[:, 16:]], 1)
    heads = a3 @ wh.T + bias[:112]                                                                # step 3
    right = a3 @ wr.T + bias[112:160]
    # the categorical heads are packed in log2 units (x log2 e); the tail columns are not
    got = np.concatenate([heads[:, :100] / model.LOG2E, right[:, :36] / model.LOG2E, heads[:, 100:105]], 1)   # reference column order
    x = torch.from_numpy(np.concatenate([feat[ia], feat[ib], ppf], 1))
    want = ref_model.pair_mlp(x, {k: v for k, v in sd.items()}).numpy()
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5)
    assert not heads[:, 105:].any() and not right[:, 36:].any()     # zero padding columns stay zero


def test_packed_blobs_follow_the_weights():
    """The packed operand blobs are cached per module and rebuilt when the weights change in place
    (load_state_dict, optimiser-style updates) -- a stale blob would silently run old weights."""
    import torch
    from cppf_b200 import model
    torch.manual_seed(1)
    m = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141).eval()
    b0 = m.tc_blob("cpu")
    assert m.tc_blob("cpu") is b0                                   # cached
    ref0 = torch.from_numpy(model.pack_tc_weights(m.state_dict()))
    assert torch.equal(b0, ref0)
    other = model.PPFEncoder(ppffcs=[84, 32, 32, 16], out_dim=141)
    m.load_state_dict(other.state_dict())                           # in-place copy: version counters move
    b1 = m.tc_blob("cpu")
    assert b1 is not b0 and torch.equal(b1, torch.from_numpy(model.pack_tc_weights(other.state_dict())))
    with torch.no_grad():
        m.final.weight.mul_(2.0)
    b2 = m.tc_blob("cpu")
    assert b2 is not b1 and not torch.equal(b2, b1)
    assert torch.equal(b2, torch.from_numpy(model.pack_tc_weights(m.state_dict())))


def test_point_encoder_tensor_core_section_reproduces_reference_features():
    """CPU: the second section of the PointEncoder blob (model.pack_pe_weights: tcgen05 operands of
    csrc/point_encoder.cu, namespace tcpe) interpreted in numpy the way point_encode_tc_kernel walks it -- canonical
    K-major [K/4][N][4] operands as hi + lo, Linear(64, 32) as two K = 32 halves, biases as ones-operand rows (k = 0 hi,
    k = 4 lo), LayerNorm vectors, output tail -- must reproduce the reference module's features on the golden fixture
    (models/sprin.py:40-107, models/model.py:63-77)."""
    d = load_golden("encoder_bottle.npz")
    sd = split_state(d, "pe/")
    blob = model.pack_pe_weights(sd)
    simt = 6 * 32 + 96 + 32 * 64 + 192 + 64 * 32 + 96 + 32 * 32 + 96 + 32 * 32 + 32 + 64 * 32 + 96 + 32 * 8 + 8   # pe::kBlobFloats
    tc = blob[simt:]
    o = [0]

    def take(n):
        v = tc[o[0]:o[0] + n]
        o[0] += n
        return v

    def operand(n, k):                                           # -> W[n][k] = hi + lo
        hi = take(n * k).reshape(k // 4, n, 4).transpose(1, 0, 2).reshape(n, k)
        lo = take(n * k).reshape(k // 4, n, 4).transpose(1, 0, 2).reshape(n, k)
        assert np.all((hi.view(np.uint32) & 0x1FFF) == 0)        # tf32: low 13 mantissa bits clear
        return hi.astype(np.float64) + lo.astype(np.float64)

    w1, w2, w3a, w3b, w4, w5 = operand(32, 8), operand(64, 32), operand(32, 32), operand(32, 32), operand(32, 32), operand(32, 32)
    assert np.all(w1[:, 6:] == 0)

    def bias(n):
        v = take(n * 8).reshape(2, n, 4)
        assert np.all(v[:, :, 1:] == 0)
        return v[0, :, 0].astype(np.float64) + v[1, :, 0].astype(np.float64)

    b1, b2, b3, b4, b5 = bias(32), bias(64), bias(32), bias(32), bias(32)
    g1, e1, g2, e2, g3, e3, g4, e4 = take(32), take(32), take(64), take(64), take(32), take(32), take(32), take(32)
    wo, bo, go, eo = take(64 * 32).reshape(64, 32), take(32), take(32), take(32)
    wa, ba = take(32 * 8).reshape(32, 8), take(8)
    assert o[0] == tc.size

    def ln(x, g, e):
        m = x.mean(-1, keepdims=True)
        v = ((x - m) ** 2).mean(-1, keepdims=True)
        return (x - m) / np.sqrt(v + 1e-5) * g + e

    pc, nrm, nbrs = d["pc"].astype(np.float64), d["nrm"].astype(np.float64), d["nbrs"]
    nb, nn = pc[nbrs], nrm[nbrs]                                 # [N, k, 3]
    ctr = pc[:, None, :]
    mean = nb.mean(1, keepdims=True)
    l1, l2, l3 = mean - nb, nb - ctr, ctr - mean
    n1, n2, n3 = [np.linalg.norm(v, axis=-1) for v in (l1, l2, l3)]
    n3 = np.broadcast_to(n3, n1.shape)
    ri = np.stack([n1, n2, n3, (l1 * l2).sum(-1) / (n1 * n2 + 1e-7), (l2 * l3).sum(-1) / (n2 * n3 + 1e-7),
                   (l3 * l1).sum(-1) / (n3 * n1 + 1e-7), np.zeros_like(n1), np.zeros_like(n1)], -1)      # K padded to 8
    x = np.maximum(ln(ri @ w1.T + b1, g1, e1), 0)
    x = np.maximum(ln(x @ w2.T + b2, g2, e2), 0)
    x = np.maximum(ln(x[..., :32] @ w3a.T + x[..., 32:] @ w3b.T + b3, g3, e3), 0)
    x = np.maximum(ln(x @ w4.T + b4, g4, e4), 0)
    kern = x @ w5.T + b5                                         # [N, k, 32]
    nf = np.stack([n2, (nn * nrm[:, None, :]).sum(-1)], -1)      # [N, k, 2]
    ct = np.einsum("nkr,nki->nri", kern, nf).reshape(len(pc), 64)
    y = ln(ct @ wo + bo, go, eo)
    glob = (y @ wa + ba).max(0)
    got = np.concatenate([y, np.broadcast_to(glob, (len(pc), 8))], 1)
    np.testing.assert_allclose(got, d["feat_nbrs"], rtol=2e-4, atol=2e-5)
