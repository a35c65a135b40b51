"""Host-side mirror of the reference's ``models/model.py`` (PointEncoder, PPFEncoder,
ResLayer) backed by the sm_100a kernels in ``libcppf_b200.so``.

Same class names, constructor kwargs, ``state_dict`` keys/shapes and call signatures as
the reference (``models/model.py:8-137``, ``models/sprin.py:63-107``), so checkpoints
load unchanged and ``nocs/inference.py:82-88,181-182,236`` can use these classes as is.
Inference only: the modules never build autograd graphs and raise if asked to.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib

# offsets of cppf_b200/csrc/mlp_layout.h
_F = 40
_OFF_PRE_WA = 0
_OFF_PRE_WB = _OFF_PRE_WA + _F * 64
_OFF_PRE_BIAS = _OFF_PRE_WB + _F * 64
_OFF_PAIR = _OFF_PRE_BIAS + 64
_FINAL_CHUNK = 48
SUPPORTED_PPFFCS = (2 * _F + 4, 32, 32, 16)


def _perm_cols(m: np.ndarray, no: int) -> np.ndarray:
    """[K, 8*no] logical column j = og + 8c  ->  stored column og*no + c (mlp_layout.h)."""
    k = m.shape[0]
    return np.ascontiguousarray(m.reshape(k, no, 8).transpose(0, 2, 1).reshape(k, 8 * no))


def pack_ppf_weights(sd, out_dim: int) -> np.ndarray:
    """Pack a PPFEncoder ``state_dict`` into the blob the kernels read (mlp_layout.h)."""
    g = lambda k: sd[k].detach().to("cpu", torch.float32).numpy()
    w1_0, b1_0 = g("res_layers.0.fc1.weight"), g("res_layers.0.fc1.bias")
    w0_0, b0_0 = g("res_layers.0.fc0.weight"), g("res_layers.0.fc0.bias")
    w2_0, b2_0 = g("res_layers.0.fc2.weight"), g("res_layers.0.fc2.bias")
    w1_1, b1_1 = g("res_layers.1.fc1.weight"), g("res_layers.1.fc1.bias")
    w2_1, b2_1 = g("res_layers.1.fc2.weight"), g("res_layers.1.fc2.bias")
    w1_2, b1_2 = g("res_layers.2.fc1.weight"), g("res_layers.2.fc1.bias")
    w0_2, b0_2 = g("res_layers.2.fc0.weight"), g("res_layers.2.fc0.bias")
    w2_2, b2_2 = g("res_layers.2.fc2.weight"), g("res_layers.2.fc2.bias")
    wf, bf = g("final.weight"), g("final.bias")
    if w1_0.shape != (32, 2 * _F + 4) or w1_1.shape != (32, 32) or w1_2.shape != (16, 32) or wf.shape != (out_dim, 16):
        raise NotImplementedError("the sm_100a pair MLP is specialised to ppffcs=[84,32,32,16] "
                                  f"(nocs/inference.py:83); got {w1_0.shape}, {w1_1.shape}, {w1_2.shape}, {wf.shape}")
    outp = (out_dim + _FINAL_CHUNK - 1) // _FINAL_CHUNK * _FINAL_CHUNK
    both0 = np.concatenate([w1_0, w0_0], 0)                       # [64, 84]
    parts = [
        both0[:, :_F].T,                                          # PRE_WA [40,64]
        both0[:, _F:2 * _F].T,                                    # PRE_WB [40,64]
        np.concatenate([b1_0, b0_0 + b2_0]),                      # PRE_BIAS
        both0[:, 2 * _F:].T,                                      # W_PPF [4,64]
        _perm_cols(w2_0.T, 4),
        _perm_cols(w1_1.T, 4), _perm_cols(b1_1[None], 4),
        _perm_cols(w2_1.T, 4), _perm_cols(b2_1[None], 4),
        _perm_cols(np.concatenate([w1_2, w0_2], 0).T, 4),
        _perm_cols(np.concatenate([b1_2, b0_2 + b2_2])[None], 4),
        _perm_cols(w2_2.T, 2),
    ]
    wf_p = np.zeros((16, outp), np.float32)
    wf_p[:, :out_dim] = wf.T
    bf_p = np.zeros((1, outp), np.float32)
    bf_p[0, :out_dim] = bf
    nch = outp // _FINAL_CHUNK
    perm_f = lambda m: m.reshape(m.shape[0], nch, 6, 8).transpose(0, 1, 3, 2).reshape(m.shape[0], outp)
    parts += [perm_f(wf_p), perm_f(bf_p)]
    blob = np.concatenate([np.ascontiguousarray(p, dtype=np.float32).reshape(-1) for p in parts])
    assert blob.size == _lib.lib().cppf_ppf_blob_floats(out_dim), (blob.size, out_dim)
    return blob


TR_BINS, ROT_BINS = 32, 36          # config/config.yaml:7-8 of the reference
HEAD_OUT_DIM = 2 * TR_BINS + 2 * ROT_BINS + 2 + 3


def pack_head_weights(sd) -> np.ndarray:
    """`final` layer re-cut head by head for the fused encode+sample kernel (csrc/fused.cu):
    mu | nu : W[16][32] (permuted, NO=4) + b[32];  up | right : Wa[16][32] + ba[32] + Wb[16][8] + bb[8]
    (bins 32..35 in columns 0..3 of Wb);  tail : W[16][8] + b[8] (aux_up, aux_right, log-scale x3)."""
    wf = sd["final.weight"].detach().to("cpu", torch.float32).numpy()
    bf = sd["final.bias"].detach().to("cpu", torch.float32).numpy()
    if wf.shape != (HEAD_OUT_DIM, 16):
        raise NotImplementedError(f"fused heads need out_dim == 2*{TR_BINS}+2*{ROT_BINS}+5 = {HEAD_OUT_DIM} "
                                  f"(nocs/inference.py:83), got {wf.shape[0]}")
    parts = []
    for r0 in (0, TR_BINS):
        parts += [_perm_cols(wf[r0:r0 + 32].T, 4), _perm_cols(bf[None, r0:r0 + 32], 4)]
    for r0 in (2 * TR_BINS, 2 * TR_BINS + ROT_BINS):
        wb = np.zeros((16, 8), np.float32)
        bb = np.zeros(8, np.float32)
        wb[:, :ROT_BINS - 32] = wf[r0 + 32:r0 + ROT_BINS].T
        bb[:ROT_BINS - 32] = bf[r0 + 32:r0 + ROT_BINS]
        parts += [_perm_cols(wf[r0:r0 + 32].T, 4), _perm_cols(bf[None, r0:r0 + 32], 4), wb, bb]
    wt = np.zeros((16, 8), np.float32)
    bt = np.zeros(8, np.float32)
    wt[:, :5] = wf[HEAD_OUT_DIM - 5:].T
    bt[:5] = bf[HEAD_OUT_DIM - 5:]
    parts += [wt, bt]
    blob = np.concatenate([np.ascontiguousarray(q, dtype=np.float32).reshape(-1) for q in parts])
    assert blob.size == _lib.lib().cppf_head_blob_floats(), blob.size
    return blob


def _row_constant_operand(v: np.ndarray, n_pad: int) -> np.ndarray:
    """Vector v -> [2 K planes][n_pad][4] B operand of csrc/encode_tc.cu's ones-operand MMA: k = 0 holds the tf32 `hi` part
    of v, k = 4 its fp32 remainder `lo` (the A side is 1 in k = 0 and k = 4), everything else zero."""
    m = np.zeros(n_pad, np.float32)
    m[:v.shape[0]] = v
    hi = (m.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    out = np.zeros((2, n_pad, 4), np.float32)
    out[0, :, 0] = hi
    out[1, :, 0] = m - hi
    return out.reshape(-1)


def _canon_hi_lo(w: np.ndarray, n_pad: int, k_pad: int) -> list:
    """nn.Linear weight [n, k] -> tcgen05 no-swizzle K-major operand [k_pad/4][n_pad][4], split into the
    tf32 `hi` part (low 13 mantissa bits cleared) and the fp32 remainder `lo` (csrc/encode_tc.cu)."""
    m = np.zeros((n_pad, k_pad), np.float32)
    m[:w.shape[0], :w.shape[1]] = w
    hi = (m.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = (m - hi).astype(np.float32)
    canon = lambda a: np.ascontiguousarray(a.reshape(n_pad, k_pad // 4, 4).transpose(1, 0, 2)).reshape(-1)
    return [canon(hi), canon(lo)]


LOG2E = 1.4426950408889634


def _param_key(module, device):
    """Cheap cache key of a module's weights (the packed blobs are rebuilt when it changes): in-place updates
    (load_state_dict, optimiser steps) bump every parameter's version counter, .to()/.cuda() move all storages
    (first and last data_ptr).  ~5 us instead of ~60 us for walking module.parameters() with data_ptr() each."""
    pl = module.__dict__.get("_cppf_plist")
    if pl is None:
        pl = list(module.parameters())
        module.__dict__["_cppf_plist"] = pl
    return (str(device), pl[0].data_ptr(), pl[-1].data_ptr()) + tuple(p._version for p in pl)


def pack_tc_weights(sd) -> np.ndarray:
    """Pack a PPFEncoder ``state_dict`` for the tcgen05 encoder (csrc/encode_tc.cu, "chain algebra"):
    adjacent linear maps of models/model.py:26-31,134-137 are composed here in float64 and rounded once
    to fp32, so the kernel runs 4 MMA steps per tile instead of 8.  Layout = the kOff* constants there."""
    g = lambda k: sd[k].detach().to("cpu", torch.float64).numpy()
    W1_0, b1_0 = g("res_layers.0.fc1.weight"), g("res_layers.0.fc1.bias")
    W0_0, b0_0 = g("res_layers.0.fc0.weight"), g("res_layers.0.fc0.bias")
    W2_0, b2_0 = g("res_layers.0.fc2.weight"), g("res_layers.0.fc2.bias")
    W1_1, b1_1 = g("res_layers.1.fc1.weight"), g("res_layers.1.fc1.bias")
    W2_1, b2_1 = g("res_layers.1.fc2.weight"), g("res_layers.1.fc2.bias")
    W1_2, b1_2 = g("res_layers.2.fc1.weight"), g("res_layers.2.fc1.bias")
    W0_2, b0_2 = g("res_layers.2.fc0.weight"), g("res_layers.2.fc0.bias")
    W2_2, b2_2 = g("res_layers.2.fc2.weight"), g("res_layers.2.fc2.bias")
    Wf, bf = g("final.weight"), g("final.bias")
    if W1_0.shape != (32, 2 * _F + 4) or W1_1.shape != (32, 32) or W1_2.shape != (16, 32) or Wf.shape != (HEAD_OUT_DIM, 16):
        raise NotImplementedError("the tcgen05 pair encoder is specialised to ppffcs=[84,32,32,16], "
                                  f"out_dim={HEAD_OUT_DIM} (nocs/inference.py:83)")
    W10_2 = np.concatenate([W1_2, W0_2], 0)                           # [32,32]
    b10_2 = np.concatenate([b1_2, b0_2 + b2_2])
    br = b0_0 + b2_0                                                  # bias of r = fc0_0(x) + b2_0
    # rows of the three front-end quantities as functions of x = [feat_a(40) feat_b(40) ppf(4)]
    V1, Q1, Q2 = W1_0, W1_1 @ W0_0, W10_2 @ W0_0                      # each [32,84]
    bV1, bQ1, bQ2 = b1_0, W1_1 @ br + b1_1, W10_2 @ (br + b2_1) + b10_2
    front = np.concatenate([V1, Q1, Q2], 0)                           # [96,84]
    pre_w = np.concatenate([front[:, :_F].T, front[:, _F:2 * _F].T], 1)          # [40,192]: A side | B side
    pre_b = np.concatenate([bV1, bQ1, bQ2, np.zeros(96)])
    r_up, r_rt, r_tail = 2 * TR_BINS, 2 * TR_BINS + ROT_BINS, 2 * TR_BINS + 2 * ROT_BINS
    # the categorical heads are only ever consumed as softmax probabilities (nocs/inference.py:185-186,245-256), which the
    # kernel evaluates as 2^(l' - max l') with l' = l * log2(e): the factor is folded into their rows here (the tail rows --
    # aux logits and log-scales -- are outputs and stay unscaled)
    head_rows = np.concatenate([Wf[:r_rt] * LOG2E, Wf[r_tail:]], 0)   # mu | nu | up | tail  (105 rows)
    WH = np.concatenate([head_rows @ W2_2, head_rows], 1)             # acts on [u2 ; r2]  [105,32]
    WR = np.concatenate([Wf[r_rt:r_tail] @ W2_2, Wf[r_rt:r_tail]], 1) * LOG2E    # right head [36,32]
    f32 = lambda a: np.asarray(a, np.float64).astype(np.float32)
    parts = [f32(pre_w).reshape(-1), f32(pre_b)]
    parts += _canon_hi_lo(f32(front[:, 2 * _F:]), 96, 8)              # Wp : ppf columns
    parts += _canon_hi_lo(f32(np.concatenate([W1_1 @ W2_0, W10_2 @ W2_0], 0)), 64, 32)    # Ws1
    parts += _canon_hi_lo(f32(W10_2 @ W2_1), 32, 32)                  # Ws2
    parts += _canon_hi_lo(f32(WH), 112, 32)
    parts += _canon_hi_lo(f32(WR), 48, 32)
    # head biases as K = 8 MMA operands (bias in k = 0), multiplied by a constant ones operand in the kernel
    bh = np.zeros((112, 8))
    bh[0:r_rt, 0] = bf[:r_rt] * LOG2E
    bh[r_rt:r_rt + 5, 0] = bf[r_tail:]
    br_ = np.zeros((48, 8))
    br_[0:ROT_BINS, 0] = bf[r_rt:r_tail] * LOG2E
    parts += [_row_constant_operand(f32(bh[:, 0]), 112), _row_constant_operand(f32(br_[:, 0]), 48)]
    blob = np.concatenate([np.ascontiguousarray(q, dtype=np.float32).reshape(-1) for q in parts])
    assert blob.size == _lib.lib().cppf_tc_blob_floats(), blob.size
    return blob


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def _f32c(t, device):
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(t)
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def _no_grad_only(*tensors):
    if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        raise RuntimeError("cppf_b200 is inference-only (training/backward is out of scope): "
                           "call under torch.no_grad() with inputs that do not require grad")


class ResLayer(nn.Module):
    """Parameter container with the reference's names (models/model.py:8-25).  The
    arithmetic (models/model.py:26-31) runs fused inside the pair-MLP kernel."""

    def __init__(self, dim_in, dim_out, bn=False) -> None:
        super().__init__()
        assert bn is False
        self.fc1 = nn.Linear(dim_in, dim_out)
        self.fc2 = nn.Linear(dim_out, dim_out)
        self.fc0 = nn.Linear(dim_in, dim_out) if dim_in != dim_out else None

    def forward(self, x):
        raise RuntimeError("ResLayer is evaluated inside PPFEncoder's fused sm_100a kernel; "
                           "there is no stand-alone (PyTorch) path")


class PPFEncoder(nn.Module):
    """Drop-in for reference ``models/model.py:80-137``."""

    def __init__(self, ppffcs, out_dim) -> None:
        super().__init__()
        self.ppffcs = tuple(int(v) for v in ppffcs)
        self.out_dim = int(out_dim)
        self.res_layers = nn.ModuleList(ResLayer(i, o, bn=False) for i, o in zip(ppffcs[:-1], ppffcs[1:]))
        self.final = nn.Linear(ppffcs[-1], out_dim)
        self._blob = None
        self._blob_key = None
        self._hblob = None
        self._hblob_key = None
        self._tcblob = None
        self._tcblob_key = None

    # ---- weight blob cache (re-packed whenever parameters move or change)
    def weight_blob(self, device) -> torch.Tensor:
        key = _param_key(self, device)
        if self._blob is None or self._blob_key != key:
            if self.ppffcs != SUPPORTED_PPFFCS:
                raise NotImplementedError(f"pair MLP kernels are specialised to ppffcs={list(SUPPORTED_PPFFCS)}, "
                                          f"got {list(self.ppffcs)}")
            blob = pack_ppf_weights(self.state_dict(), self.out_dim)
            self._blob = torch.from_numpy(blob).to(device)
            self._blob_key = key
        return self._blob

    def head_blob(self, device) -> torch.Tensor:
        """Head-wise `final` weights for the fused encode+sample kernel (pack_head_weights)."""
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in self.final.parameters())
        if self._hblob is None or self._hblob_key != key:
            self._hblob = torch.from_numpy(pack_head_weights(self.state_dict())).to(device)
            self._hblob_key = key
        return self._hblob

    def tc_blob(self, device) -> torch.Tensor:
        """Weights in tcgen05 operand layout for the tensor-core encoder (pack_tc_weights)."""
        key = _param_key(self, device)
        if self._tcblob is None or self._tcblob_key != key:
            if self.ppffcs != SUPPORTED_PPFFCS:
                raise NotImplementedError(f"pair MLP kernels are specialised to ppffcs={list(SUPPORTED_PPFFCS)}")
            self._tcblob = torch.from_numpy(pack_tc_weights(self.state_dict())).to(device)
            self._tcblob_key = key
        return self._tcblob

    def tc_preproject(self, feat: torch.Tensor) -> torch.Tensor:
        """Per-point table of the tcgen05 encoder's front end (cppf_tc_preproject): 192 columns per point,
        stored planar as [48 float4 chunks][N][4] so that dense-mode gathers are coalesced."""
        n = feat.shape[0]
        L = _lib.lib()
        table = torch.empty((L.cppf_tc_table_cols() // 4, n, 4), dtype=torch.float32, device=feat.device)
        _lib.check(L.cppf_tc_preproject(feat.data_ptr(), self.tc_blob(feat.device).data_ptr(), table.data_ptr(), n,
                                        _stream_ptr(feat.device)), "cppf_tc_preproject")
        return table

    def preproject(self, feat: torch.Tensor) -> torch.Tensor:
        """Per-point table of ResLayer-0's feature columns (cppf_ppf_preproject)."""
        n = feat.shape[0]
        table = torch.empty((n, 128), dtype=torch.float32, device=feat.device)
        L = _lib.lib()
        _lib.check(L.cppf_ppf_preproject(feat.data_ptr(), self.weight_blob(feat.device).data_ptr(), table.data_ptr(),
                                         n, _stream_ptr(feat.device)), "cppf_ppf_preproject")
        return table

    def _encode(self, pc, pc_normal, feat, idxs, dist, cols=None):
        _no_grad_only(pc, pc_normal, feat, *self.parameters())
        device = feat.device if isinstance(feat, torch.Tensor) else torch.device("cuda")
        if device.type != "cuda":
            raise RuntimeError("cppf_b200.PPFEncoder runs on CUDA tensors only (no CPU fallback)")
        with torch.cuda.device(device):
            pc, pc_normal, feat = _f32c(pc, device), _f32c(pc_normal, device), _f32c(feat, device)
            n = pc.shape[0]
            if feat.shape != (n, _F) or pc.shape != (n, 3) or pc_normal.shape != (n, 3):
                raise ValueError(f"expected pc/normal [N,3] and feat [N,{_F}], got {tuple(pc.shape)}, "
                                 f"{tuple(pc_normal.shape)}, {tuple(feat.shape)}")
            table = self.preproject(feat)
            col0, ncol = (0, self.out_dim) if cols is None else cols
            if idxs is None:
                n_pairs, idx_ptr, is64 = n * n, None, 0
                if dist is not None:
                    dist = _f32c(dist, device)
                    if dist.shape != (n, n):
                        raise ValueError("dist must be [N,N]")
            else:
                if isinstance(idxs, np.ndarray):
                    idxs = torch.from_numpy(np.ascontiguousarray(idxs))
                idxs = idxs.to(device)
                if idxs.dtype not in (torch.int64, torch.int32):
                    idxs = idxs.long()
                idxs = idxs.contiguous()
                if idxs.dim() != 2 or idxs.shape[1] != 2:
                    raise ValueError("idxs must be [P,2]")
                n_pairs, idx_ptr, is64 = idxs.shape[0], idxs.data_ptr(), int(idxs.dtype == torch.int64)
            out = torch.empty((n_pairs, ncol), dtype=torch.float32, device=device)
            L = _lib.lib()
            _lib.check(L.cppf_ppf_encode(pc.data_ptr(), pc_normal.data_ptr(), table.data_ptr(),
                                         self.weight_blob(device).data_ptr(), idx_ptr, is64,
                                         dist.data_ptr() if dist is not None else None, out.data_ptr(),
                                         n, n_pairs, self.out_dim, col0, ncol, _stream_ptr(device)), "cppf_ppf_encode")
        return out

    def forward(self, pc, pc_normal, feat, dist=None, idxs=None):
        """models/model.py:89-115.  Batched [1,N,*] inputs; with ``idxs`` -> [1,P,out_dim],
        dense -> [1,N,N,out_dim] (all ordered pairs, row = point a)."""
        if idxs is not None:
            return self.forward_with_idx(pc[0], pc_normal[0], feat[0], idxs)[None]
        if pc.shape[0] != 1:
            raise NotImplementedError("the reference drives batch size 1 (nocs/inference.py:174); got batch "
                                      f"{pc.shape[0]}")
        n = pc.shape[1]
        out = self._encode(pc[0], pc_normal[0], feat[0], None, None if dist is None else dist[0])
        return out.view(1, n, n, self.out_dim)

    def forward_with_idx(self, pc, pc_normal, feat, idxs):
        """models/model.py:117-137.  pc,pc_normal [N,3], feat [N,40], idxs [P,2] -> [P,out_dim]."""
        return self._encode(pc, pc_normal, feat, idxs, None)


# ---------------------------------------------------------------------------------------
# Point encoder (reference models/model.py:34-77, models/sprin.py:63-107).
def conv_kernel(iunit, ounit, *hunits):
    """models/sprin.py:63-71 layout, so Sequential indices (state_dict keys) match."""
    layers = []
    for unit in hunits:
        layers += [nn.Linear(iunit, unit), nn.LayerNorm(unit), nn.ReLU()]
        iunit = unit
    layers.append(nn.Linear(iunit, ounit))
    return nn.Sequential(*layers)


class GlobalInfoProp(nn.Module):
    def __init__(self, n_in, n_global):
        super().__init__()
        self.linear = nn.Linear(n_in, n_global)


class SparseSO3Conv(nn.Module):
    def __init__(self, rank, n_in, n_out, *kernel_interns, layer_norm=True):
        super().__init__()
        self.kernel = conv_kernel(6, rank, *kernel_interns)
        self.outnet = nn.Linear(rank * n_in, n_out)
        self.rank = rank
        self.layer_norm = nn.LayerNorm(n_out) if layer_norm else None


def pack_pe_weights(sd) -> np.ndarray:
    """Pack a PointEncoder ``state_dict`` (spfcs=[32,64,32,32], rank 32, 2 neighbour features, out_dim 32) for the
    fused kNN+SPRIN kernel; layout = the k* constants of csrc/point_encoder.cu."""
    g = lambda k: sd[k].detach().to("cpu", torch.float32).numpy()
    parts = []
    dims = [(6, 32), (32, 64), (64, 32), (32, 32), (32, 32)]
    for li, (seq, (din, dout)) in enumerate(zip((0, 3, 6, 9, 12), dims)):
        w, b = g(f"spconvs.0.kernel.{seq}.weight"), g(f"spconvs.0.kernel.{seq}.bias")
        if w.shape != (dout, din):
            raise NotImplementedError(f"fused point encoder needs spfcs=[32,64,32,32], rank 32; kernel.{seq} is {w.shape}")
        no = dout // 8
        parts += [_perm_cols(w.T, no), _perm_cols(b[None], no)]
        if li < 4:
            parts += [g(f"spconvs.0.kernel.{seq + 1}.weight"), g(f"spconvs.0.kernel.{seq + 1}.bias")]
    wo = g("spconvs.0.outnet.weight")
    if wo.shape != (32, 64) or "spconvs.0.layer_norm.weight" not in sd or g("aggrs.0.linear.weight").shape != (8, 32):
        raise NotImplementedError("fused point encoder needs num_nbr_feats=2, out_dim=32 with layer_norm")
    tail = [wo.T, g("spconvs.0.outnet.bias"), g("spconvs.0.layer_norm.weight"), g("spconvs.0.layer_norm.bias"),
            g("aggrs.0.linear.weight").T, g("aggrs.0.linear.bias")]
    parts += tail
    # second section: the same layers as tcgen05 operands (point_encode_tc_kernel, tcpe::o* offsets): canonical K-major hi / lo
    # weights -- Linear(64, 32) as two K = 32 halves --, biases as ones-operand rows, LayerNorm vectors, then the tail again
    ws = [g(f"spconvs.0.kernel.{seq}.weight") for seq in (0, 3, 6, 9, 12)]
    bs = [g(f"spconvs.0.kernel.{seq}.bias") for seq in (0, 3, 6, 9, 12)]
    parts += _canon_hi_lo(ws[0], 32, 8) + _canon_hi_lo(ws[1], 64, 32) + _canon_hi_lo(ws[2][:, :32], 32, 32)
    parts += _canon_hi_lo(ws[2][:, 32:], 32, 32) + _canon_hi_lo(ws[3], 32, 32) + _canon_hi_lo(ws[4], 32, 32)
    parts += [_row_constant_operand(b, b.shape[0]) for b in bs]
    for seq in (1, 4, 7, 10):
        parts += [g(f"spconvs.0.kernel.{seq}.weight"), g(f"spconvs.0.kernel.{seq}.bias")]
    parts += tail
    blob = np.concatenate([np.ascontiguousarray(q, dtype=np.float32).reshape(-1) for q in parts])
    assert blob.size == _lib.lib().cppf_pe_blob_floats(), blob.size
    return blob


class PointEncoder(nn.Module):
    """Drop-in for reference ``models/model.py:34-77`` (num_layers=1 as at every call site,
    nocs/inference.py:82).  Runs as two sm_100a kernels (csrc/point_encoder.cu): ``cppf_knn`` (exact k nearest
    neighbours without the N x N matrix) and ``cppf_point_encode`` (SPRIN convolution + LayerNorm + global max, one warp
    per point).  The module only holds the parameters under the reference's state_dict keys; there is no eager path."""

    def __init__(self, k, spfcs, out_dim, num_layers=2, num_nbr_feats=2) -> None:
        super().__init__()
        if num_layers != 1:
            raise NotImplementedError("every reference call site uses num_layers=1 (nocs/inference.py:82)")
        self.k = k
        self.spconvs = nn.ModuleList([SparseSO3Conv(32, num_nbr_feats, out_dim, *spfcs)])
        self.aggrs = nn.ModuleList([GlobalInfoProp(out_dim, out_dim // 4)])

    # ---- fused sm_100a path (csrc/point_encoder.cu)
    def _fused_ok(self):
        c = self.spconvs[0]
        return (self.k <= 64 and c.rank == 32 and c.layer_norm is not None and tuple(c.outnet.weight.shape) == (32, 64)
                and [tuple(m.weight.shape) for m in c.kernel if isinstance(m, nn.Linear)] ==
                [(32, 6), (64, 32), (32, 64), (32, 32), (32, 32)])

    def pe_blob(self, device) -> torch.Tensor:
        key = _param_key(self, device)
        if getattr(self, "_peblob", None) is None or self._peblob_key != key:
            self._peblob = torch.from_numpy(pack_pe_weights(self.state_dict())).to(device)
            self._peblob_key = key
        return self._peblob

    def knn(self, pc: torch.Tensor) -> torch.Tensor:
        """Exact k nearest neighbours of every point (self included), [N,k] int64, unordered -- the index set of
        torch.topk(dist, k, largest=False, sorted=False) (models/model.py:47) computed without the N x N matrix."""
        n = pc.shape[0]
        out = torch.empty((n, self.k), dtype=torch.int64, device=pc.device)
        with torch.cuda.device(pc.device):
            _lib.check(_lib.lib().cppf_knn(pc.data_ptr(), n, self.k, out.data_ptr(), _stream_ptr(pc.device)), "cppf_knn")
        return out

    def encode_fused(self, pc: torch.Tensor, pc_normal: torch.Tensor, nbrs_idx: torch.Tensor = None) -> torch.Tensor:
        """pc, pc_normal [N,3] CUDA float32, nbrs_idx [N,k] int64 (None -> exact kNN) -> feat [N,40]."""
        _no_grad_only(pc, pc_normal, *self.parameters())
        if pc.device.type != "cuda":
            raise RuntimeError("cppf_b200.PointEncoder's fused path runs on CUDA tensors only (no CPU fallback)")
        pc, pc_normal = _f32c(pc, pc.device), _f32c(pc_normal, pc.device)
        n = pc.shape[0]
        if nbrs_idx is None:
            nbrs_idx = self.knn(pc)
        nbrs_idx = nbrs_idx.to(torch.int64).contiguous()
        if nbrs_idx.shape != (n, self.k):
            raise ValueError(f"nbrs_idx must be [N,{self.k}]")
        feat = torch.empty((n, 40), dtype=torch.float32, device=pc.device)
        glob = torch.empty(8, dtype=torch.float32, device=pc.device)
        with torch.cuda.device(pc.device):
            _lib.check(_lib.lib().cppf_point_encode(pc.data_ptr(), pc_normal.data_ptr(), nbrs_idx.data_ptr(),
                                                    self.pe_blob(pc.device).data_ptr(), feat.data_ptr(), glob.data_ptr(), n,
                                                    self.k, _stream_ptr(pc.device)), "cppf_point_encode")
        return feat

    def forward(self, pc, pc_normal, dist):
        """models/model.py:46-61: k nearest (self included) from the caller's distance matrix."""
        _no_grad_only(pc, pc_normal, dist)
        nbrs_idx = torch.topk(dist, self.k, largest=False, sorted=False)[1]
        return self.forward_nbrs(pc, pc_normal, nbrs_idx)

    def forward_nbrs(self, pc, pc_normal, nbrs_idx):
        """models/model.py:63-77.  pc,pc_normal [B,N,3], nbrs_idx [B,N,K] -> [B,N,out+out//4]."""
        _no_grad_only(pc, pc_normal)
        if pc.is_cuda and pc.shape[0] == 1 and self._fused_ok():
            return self.encode_fused(pc[0], pc_normal[0], nbrs_idx[0])[None]
        raise NotImplementedError(
            "cppf_b200.PointEncoder runs only through its sm_100a kernel (csrc/point_encoder.cu): CUDA tensors, batch 1, "
            "k <= 64, spfcs=[32,64,32,32], rank 32, out_dim 32 with LayerNorm -- the configuration of every reference call "
            "site (nocs/inference.py:82).  There is no eager fallback.")
