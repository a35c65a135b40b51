"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the scene-scale pieces of nocs/zero_shot.ipynb (SURVEY.md section 8
row f4).  Only tests/ may import this.  "cell N" = N-th code cell of the notebook (it is JSON; there are no line numbers).
The Gaussian smoothing is scipy.ndimage.gaussian_filter itself (scipy is installed in the image), so that piece is
pinned by the very library the reference calls.  The reference ships no tests or vectors for the rest: PARITY UNPINNED
beyond these line-by-line restatements."""
from __future__ import annotations

import numpy as np


def pair_filter(pc, pc_normal, point_idxs):
    """cell 6: boolean mask of the pairs that are KEPT."""
    n1 = pc_normal[point_idxs[:, 0]]
    n2 = pc_normal[point_idxs[:, 1]]
    ab = pc[point_idxs[:, 0]] - pc[point_idxs[:, 1]]
    ab = ab / (np.linalg.norm(ab, axis=-1, keepdims=True) + 1e-7)
    ppf = np.stack([np.sum(n1 * n2, -1), np.sum(ab * n1, -1), np.sum(ab * n2, -1)], -1)
    mask = (np.abs(ppf[:, 0]) > 0.9) & (np.abs(ppf[:, 1]) < 0.1) & (np.abs(ppf[:, 2]) < 0.1)
    return ~mask


def proposals(smoothed_grid, thresh=50, margin=10, rel_stop=0.7, max_props=64):
    """cell 9, verbatim control flow (the grid is modified in place).  -> list of (loc int[3], value, contrast)."""
    out = []
    max_val = None
    while len(out) < max_props:
        loc = np.array(np.unravel_index([np.argmax(smoothed_grid, axis=None)], smoothed_grid.shape)).T[::-1][0]
        lll = np.maximum(np.array([0, 0, 0]), loc - margin)
        rrr = np.minimum(np.array(smoothed_grid.shape) - 1, loc + margin)
        with np.errstate(all="ignore"):
            nbr_val = (np.mean(smoothed_grid[lll[0]:rrr[0], lll[1], lll[2]])
                       + np.mean(smoothed_grid[lll[0]:rrr[0], lll[1], rrr[2]])
                       + np.mean(smoothed_grid[lll[0]:rrr[0], rrr[1], lll[2]])
                       + np.mean(smoothed_grid[lll[0]:rrr[0], rrr[1], rrr[2]])
                       + np.mean(smoothed_grid[lll[0], lll[1]:rrr[1], lll[2]])
                       + np.mean(smoothed_grid[lll[0], lll[1]:rrr[1], rrr[2]])
                       + np.mean(smoothed_grid[rrr[0], lll[1]:rrr[1], lll[2]])
                       + np.mean(smoothed_grid[rrr[0], lll[1]:rrr[1], rrr[2]])
                       + np.mean(smoothed_grid[lll[0], lll[1], lll[2]:rrr[2]])
                       + np.mean(smoothed_grid[lll[0], rrr[1], lll[2]:rrr[2]])
                       + np.mean(smoothed_grid[rrr[0], lll[1], lll[2]:rrr[2]])
                       + np.mean(smoothed_grid[rrr[0], rrr[1], lll[2]:rrr[2]])) / 12
        diff = smoothed_grid[loc[0], loc[1], loc[2]] - nbr_val
        if diff > thresh:
            if max_val is None:
                max_val = diff
            out.append((loc.copy(), float(smoothed_grid[loc[0], loc[1], loc[2]]), float(diff)))
        if not (diff >= thresh) or (max_val is not None and diff < max_val * rel_stop):
            break
        smoothed_grid[lll[0]:rrr[0], lll[1]:rrr[1], lll[2]:rrr[2]] = 0
    return out


def instance_points(point_idxs_masked, n_points, min_contrib=12):
    """cell 11, "unsupervised instance segmentation": points that occur in more than `min_contrib` surviving pair slots,
    and the surviving pairs that touch such a point."""
    contrib = np.bincount(point_idxs_masked.reshape(-1), minlength=n_points)
    keep_pt = contrib > min_contrib
    keep_pair = keep_pt[point_idxs_masked[:, 0]] | keep_pt[point_idxs_masked[:, 1]]
    return keep_pt, keep_pair


def blob_grid(shape, centres, heights, sigma=2.0):
    """Test helper: a vote grid with Gaussian blobs of the given peak heights."""
    g = np.zeros(shape, np.float32)
    x, y, z = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    for c, h in zip(centres, heights):
        g += h * np.exp(-((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) / (2 * sigma ** 2)).astype(np.float32)
    return g
