"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the per-object pre-processing that feeds the hot path
(SURVEY.md section 8 row f2).  Only tests/ may import this.

* ``backproject``        follows utils/util.py:598-631 of the reference line by line (pinned: the reference function
                         itself, ast-extracted, is executed on the demo depth frame by oracle/make_golden.py ->
                         tests/golden/preprocess_demo.npz).
* ``sparse_quantize``    stands in for ``ME.utils.sparse_quantize(pc, return_index=True, quantization_size=res)[1]``
                         (nocs/inference.py:140).  MinkowskiEngine==0.5.4 (README.md:77) is a third-party dependency
                         that is NOT vendored in the reference tree and not installed here; its published behaviour is
                         restated: voxel = floor(coord / quantization_size) as int32, one point kept per occupied
                         voxel.  Which point of a voxel is kept and in what order the indices come back is an
                         implementation detail of ME's hash map; this restatement keeps the FIRST point of every voxel
                         and returns the indices in increasing order.  PARITY UNPINNED (no reference test or vector).
* ``estimate_normals``   stands in for open3d==0.12.0 ``estimate_normals(KDTreeSearchParamKNN(knn))``
                         (utils/util.py:61-65; README.md:76; not vendored, not installed): for every point the
                         covariance of its knn nearest neighbours (the point itself included) and the eigenvector of
                         the smallest eigenvalue; open3d leaves the sign of that eigenvector unspecified.  PARITY
                         UNPINNED; tests compare up to sign.
"""
from __future__ import annotations

import numpy as np


def backproject(depth, intrinsics, instance_mask):
    """utils/util.py:598-631.  -> (pts float64 [M,3], (rows, cols))."""
    intrinsics_inv = np.linalg.inv(intrinsics)                                  # :599
    non_zero_mask = depth > 0                                                   # :609
    final_instance_mask = np.logical_and(instance_mask, non_zero_mask)          # :610
    idxs = np.where(final_instance_mask)                                        # :612
    grid = np.array([idxs[1], idxs[0]])                                         # :613  (u = column, v = row)
    uv_grid = np.concatenate((grid, np.ones([1, grid.shape[1]])), axis=0)       # :619-621
    xyz = np.transpose(intrinsics_inv @ uv_grid)                                # :623-624
    z = depth[idxs[0], idxs[1]]                                                 # :626
    pts = xyz * z[:, np.newaxis] / xyz[:, -1:]                                  # :629
    pts[:, 0] = -pts[:, 0]                                                      # :630
    pts[:, 1] = -pts[:, 1]                                                      # :631
    return pts, idxs


def object_cloud(depth, intrinsics, instance_mask, noise=None):
    """nocs/inference.py:131-137 without the random jitter unless `noise` ([M,3], the reference's
    clip(res/4 * randn, -res/2, res/2)) is injected: metres, camera axes flipped back."""
    pc, idxs = backproject(depth, intrinsics, instance_mask)                    # :131
    pc /= 1000                                                                  # :132
    if noise is not None:
        pc = pc + noise                                                         # :134
    pc[:, 0] = -pc[:, 0]                                                        # :136
    pc[:, 1] = -pc[:, 1]                                                        # :137
    return pc, idxs


def sparse_quantize(coords, quantization_size):
    """-> int64 indices of the first point of every occupied voxel, increasing."""
    vox = np.floor(np.asarray(coords, np.float64) / quantization_size).astype(np.int32)
    _, first = np.unique(vox, axis=0, return_index=True)
    return np.sort(first).astype(np.int64)


def knn_indices(pc, k):
    """Exact k nearest neighbours (self included), ties towards the lower index."""
    pc = np.asarray(pc, np.float32)
    d2 = ((pc[:, None, :] - pc[None, :, :]) ** 2)
    d2 = d2[..., 0] + d2[..., 1] + d2[..., 2]
    return np.argsort(d2, axis=1, kind="stable")[:, :k]


def estimate_normals(pc, knn, nbrs=None):
    """-> float64 [N,3] unit normals (sign unspecified)."""
    pc64 = np.asarray(pc, np.float64)
    if nbrs is None:
        nbrs = knn_indices(pc, min(knn, len(pc)))
    q = pc64[nbrs]                                           # [N,k,3]
    c = q - q.mean(1, keepdims=True)
    cov = np.einsum("nki,nkj->nij", c, c) / q.shape[1]
    w, v = np.linalg.eigh(cov)
    return v[:, :, 0]
