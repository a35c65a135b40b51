"""Per-image driver: the loop body of the reference's ``nocs/inference.py:108-345`` (one detected instance at a time:
mask -> cloud -> pose) without CuPy / MinkowskiEngine / open3d -- pre-processing from ``cppf_b200.preprocess``, the hot
path from ``cppf_b200.pipeline``.  Dataset I/O (segmentation pickles in, ``results_*.pkl`` out), the laptop 2-D
segmenter (:144-172, :314-323) and evaluation stay outside (SURVEY.md section 2 marks them out of scope)."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import preprocess
from .pipeline import NoSurvivorsError, PoseEstimator

SYNSET_NAMES = ["BG", "bottle", "bowl", "camera", "can", "laptop", "mug"]          # nocs/inference.py:73-81


@torch.no_grad()
def estimate_image(depth, masks, class_ids: Sequence[int], estimators: Dict[str, PoseEstimator],
                   intrinsics=preprocess.NOCS_INTRINSICS, synset_names: Sequence[str] = SYNSET_NAMES, seed: int = 0,
                   jitter: bool = True, bboxes: Optional[np.ndarray] = None, bbox_mask: bool = False, orient_normals: bool = False):
    """depth [H,W] uint16 (mm) and masks [H,W,K] bool, numpy or CUDA tensors; class_ids [K].
    -> dict(pred_RTs float32 [K,4,4], pred_scales float32 [K,3], n_points [K], valid bool [K]) like the arrays the reference
    pickles (nocs/inference.py:113-118, 336-345).  Instances with fewer points than knn, and instances for which no pair
    voted for the winning centre, keep the identity pose and are flagged valid = False.
    All objects of the image are enqueued before the first pose record is read."""
    any_est = next(iter(estimators.values()))
    dev = any_est.device
    depth_d = torch.as_tensor(np.ascontiguousarray(depth) if isinstance(depth, np.ndarray) else depth).to(dev)
    masks_d = torch.as_tensor(np.ascontiguousarray(masks) if isinstance(masks, np.ndarray) else masks).to(dev)
    k = masks_d.shape[2]
    RTs = np.tile(np.eye(4, dtype=np.float32), (k, 1, 1))                       # :113-116
    scales = np.ones((k, 3), dtype=np.float32)                                  # :117
    n_points = np.zeros(k, np.int64)
    valid = np.zeros(k, bool)
    pending = []
    for i in range(k):
        m = masks_d[:, :, i].clone()
        if bbox_mask and bboxes is not None:                                    # :121-122
            b = bboxes[i]
            m[b[0]:b[2], b[1]:b[3]] = True
        est = estimators[synset_names[int(class_ids[i])]]                       # :124-129
        cfg = est.cfg
        pts, _ = preprocess.backproject(depth_d, m, intrinsics)                 # :131-132
        if pts.shape[0] == 0:
            continue
        if jitter:                                                              # :134
            g = torch.Generator(device=dev).manual_seed(seed * 1000003 + i)
            noise = torch.clamp(cfg.res / 4 * torch.randn(pts.shape, generator=g, device=dev, dtype=torch.float64),
                                -cfg.res / 2, cfg.res / 2)
            pts = pts + noise * torch.tensor([-1.0, -1.0, 1.0], dtype=torch.float64, device=dev)     # :134-137
        pc, _ = preprocess.sparse_quantize(pts, cfg.res)                        # :140-141
        n_points[i] = pc.shape[0]
        if pc.shape[0] < cfg.knn:
            continue
        nrm = preprocess.estimate_normals(pc, cfg.knn, orient_normals)          # :142
        pending.append((i, est.estimate_fused(pc, nrm, seed=seed * 1000003 + i, sync=False)))      # :174-339
    for i, p in pending:
        try:
            out = p.result() if hasattr(p, "result") else p
        except NoSurvivorsError:
            continue
        RTs[i], scales[i], valid[i] = out["RT"], out["scales"], True            # :336-339
    return {"pred_RTs": RTs, "pred_scales": scales, "n_points": n_points, "valid": valid}
