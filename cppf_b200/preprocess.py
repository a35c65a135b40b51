"""Per-object pre-processing on the device: what nocs/inference.py:131-142 does with
numpy, MinkowskiEngine and open3d before the hot path -- back-projection of the masked depth, one point per voxel,
kNN-PCA normals -- as thin wrappers over the ``cppf_backproject`` / ``cppf_voxel_first`` / ``cppf_normals_pca`` entry
points (include/cppf_b200.h).  CUDA tensors in, CUDA tensors out; the only host round trip is the point count."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

NOCS_INTRINSICS = np.array([[591.0125, 0, 322.525], [0, 590.16775, 244.11084], [0, 0, 1]])      # nocs/inference.py:98


def _sp(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def backproject(depth: torch.Tensor, mask: torch.Tensor, intrinsics=NOCS_INTRINSICS, depth_scale: float = 1000.0):
    """utils/util.py:598-631 + nocs/inference.py:132,136-137.  depth [H,W] uint16/int16 (NOCS png, mm) or float32,
    mask [H,W] bool/uint8, both CUDA.  -> (pts float64 [M,3] metres, pix int64 [M] = row*W + col, in np.where order)."""
    dev = depth.device
    if dev.type != "cuda":
        raise RuntimeError("cppf_b200.preprocess runs on CUDA tensors only (no CPU fallback)")
    h, w = depth.shape
    if depth.dtype in (torch.uint16, torch.int16):
        is_u16 = 1
    elif depth.dtype == torch.float32:
        is_u16 = 0
    else:
        raise TypeError(f"depth must be uint16/int16 or float32, got {depth.dtype}")
    depth = depth.contiguous()
    mask = mask.to(torch.uint8).contiguous()
    L = _lib.lib()
    kinv = np.ascontiguousarray(np.linalg.inv(np.asarray(intrinsics, np.float64)))          # utils/util.py:599
    pts = torch.empty((h * w, 3), dtype=torch.float64, device=dev)
    pix = torch.empty(h * w, dtype=torch.int64, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    scratch = torch.empty(L.cppf_backproject_scratch_bytes(h, w), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.cppf_backproject(depth.data_ptr(), is_u16, mask.data_ptr(), h, w, kinv.ctypes.data_as(C.c_void_p),
                                      float(depth_scale), pts.data_ptr(), pix.data_ptr(), cnt.data_ptr(), scratch.data_ptr(),
                                      _sp(dev)), "cppf_backproject")
    m = int(cnt.item())
    return pts[:m], pix[:m]


def sparse_quantize(pts: torch.Tensor, voxel: float):
    """Stand-in for ME.utils.sparse_quantize(pc, return_index=True, quantization_size=voxel)[1] followed by
    pc[indices].astype(float32) (nocs/inference.py:140-141).  pts float64 [M,3] CUDA -> (pc float32 [N,3], index int64 [N]):
    the first point of every occupied voxel, indices increasing."""
    dev = pts.device
    pts = pts.to(torch.float64).contiguous()
    m = pts.shape[0]
    L = _lib.lib()
    pc = torch.empty((m, 3), dtype=torch.float32, device=dev)
    index = torch.empty(m, dtype=torch.int64, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    scratch = torch.empty(L.cppf_voxel_scratch_bytes(m), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.cppf_voxel_first(pts.data_ptr(), None, m, float(voxel), pc.data_ptr(), index.data_ptr(), cnt.data_ptr(),
                                      scratch.data_ptr(), _sp(dev)), "cppf_voxel_first")
    n = int(cnt.item())
    return pc[:n], index[:n]


def estimate_normals(pc: torch.Tensor, knn: int, orient: bool = False) -> torch.Tensor:
    """Stand-in for utils/util.py:61-65 (open3d KNN-PCA normals).  pc float32 [N,3] CUDA -> float32 [N,3] unit normals;
    orient=False keeps the raw eigenvector sign (open3d leaves it unspecified), True points them at the camera."""
    dev = pc.device
    pc = pc.to(torch.float32).contiguous()
    n = pc.shape[0]
    k = min(int(knn), n)
    nbrs = torch.empty((n, k), dtype=torch.int64, device=dev)
    out = torch.empty((n, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cppf_normals_pca(pc.data_ptr(), n, k, int(bool(orient)), nbrs.data_ptr(), out.data_ptr(), _sp(dev)),
                   "cppf_normals_pca")
    return out


def object_cloud(depth, mask, res: float, knn: int, intrinsics=NOCS_INTRINSICS, jitter=None, orient: bool = False):
    """nocs/inference.py:131-142 for one instance mask: back-project, (optional injected jitter, :134), one point per
    `res` voxel, normals.  -> (pc float32 [N,3], normals float32 [N,3], index int64 [N] into the masked pixels)."""
    pts, _ = backproject(depth, mask, intrinsics)
    if jitter is not None:                                      # the reference adds it in the flipped frame: x, y change sign
        j = jitter.to(pts.device, torch.float64)
        pts = pts + j * torch.tensor([-1.0, -1.0, 1.0], dtype=torch.float64, device=pts.device)
    pc, index = sparse_quantize(pts, res)
    return pc, estimate_normals(pc, knn, orient), index
