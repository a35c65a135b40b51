"""Per-object pose pipeline: the script body of the reference's ``nocs/inference.py:174-339``
(and ``sunrgbd/inference.py:142-287``) as a function over the cppf_b200 kernels.

    point encoder -> pair MLP -> sample (mu, nu) -> centre vote -> argmax -> back-vote
    -> compaction -> pair MLP on survivors -> sample axis angle -> orientation candidates
    -> sphere histogram -> aux sign -> scale -> RT

Three entry points over the same kernels: ``estimate`` (materialised logits, the literal two-pass
flow of the reference; one host round trip for the survivor count), ``estimate_fused(staged=True)``
(fused kernels launched one by one, intermediates inspectable) and ``estimate_fused`` /
``enqueue_fused`` (ONE library call per object, ``cppf_pose_fused``: nothing between the
host->device copy of the cloud and the device->host copy of the 128-byte record waits for the GPU).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

import ctypes as C

from . import _lib, fast, voting
from .synth import vote_grid_geometry


def fibonacci_sphere(samples: int) -> np.ndarray:
    """Orientation bins, same construction as the reference's utils/util.py:102-118."""
    i = np.arange(samples, dtype=np.float64)
    y = 1 - (i / float(samples - 1)) * 2
    r = np.sqrt(1 - y * y)
    th = math.pi * (3.0 - math.sqrt(5.0)) * i
    return np.stack([np.cos(th) * r, y, np.sin(th) * r], -1)


@dataclass
class PoseConfig:
    """Per-category constants (config/config.yaml:1-28 + config/category/*.yaml of the reference)
    and the inference flags of nocs/inference.py:32-43."""
    res: float = 4e-3
    vote_range: tuple = (0.25, 0.25)
    scale_mean: tuple = (0.05, 0.15, 0.05)
    up_sym: bool = True
    regress_right: bool = False
    z_right: bool = False
    tr_num_bins: int = 32
    rot_num_bins: int = 36
    knn: int = 60
    num_rots: int = 72
    angle_prec: float = 1.5
    adaptive_voting: bool = True
    n_pairs: int = 100000               # nocs/inference.py:177; 0 = all N^2 ordered pairs
    rot_subsample: int = 10000          # nocs/inference.py:279-281; 0 = use every survivor
    scale_mul: float = 2.0              # nocs/inference.py:335 (SUN RGB-D: 1.0, sunrgbd/inference.py:281)
    category: str = "bottle"
    extra: dict = field(default_factory=dict)

    @classmethod
    def from_dict(cls, d):
        keys = cls.__dataclass_fields__.keys()
        return cls(**{k: (tuple(v) if isinstance(v, (list, tuple)) else v) for k, v in d.items() if k in keys})


_WORKSPACES = {}          # (device index, stream) -> uint8 workspace of cppf_pose_fused


def release_workspaces():
    """Drop the cached device buffers (cppf_pose_fused workspaces, the row-split pair lists)."""
    _WORKSPACES.clear()
    from . import rowsplit
    rowsplit._PAIRS_CACHE.clear()


class NoSurvivorsError(RuntimeError):
    """No pair voted for the winning centre (degenerate input): the reference's script would carry NaNs into the pose
    (empty tensors through nocs/inference.py:236-335); every entry point here raises instead."""


class PendingPose:
    """A pose whose kernels are enqueued; .result() waits for its record and runs the host tail.  The vote-grid dims the
    flat argmax is unravelled with travel WITH the pending pose (the staged path knows them when it enqueues, the one-call
    path reads them from its own record), never through the estimator, so poses may be read in any order."""

    def __init__(self, est, record_host, done, n_dirs, keep=(), staged=False, dims=None):
        self.est, self.record_host, self.done, self.n_dirs, self._keep, self.staged = est, record_host, done, n_dirs, keep, staged
        self.dims = dims

    def result(self):
        self.done.synchronize()
        r = self.record_host.numpy()
        self._keep = ()
        if self.staged:
            return self.est._pose_from_record(r, self.n_dirs, self.dims)
        return self.est._pose_from_record16(r, self.n_dirs)


class PoseEstimator:
    """One category's encoders + constants.  ``estimate`` handles one object."""

    def __init__(self, point_encoder, ppf_encoder, cfg: PoseConfig, device="cuda"):
        self.pe, self.ppf, self.cfg = point_encoder, ppf_encoder, cfg
        self.device = torch.device(device)
        n_bins = int(4 * np.pi / (cfg.angle_prec / 180 * np.pi))                  # nocs/inference.py:100-101
        self.sphere_np = fibonacci_sphere(n_bins)
        self.sphere = torch.from_numpy(self.sphere_np.astype(np.float32)).to(self.device)
        self.cos_thr = float(np.float32(np.cos(cfg.angle_prec / 180 * np.pi)))     # :283
        self.scale_mean = torch.tensor(cfg.scale_mean, dtype=torch.float32, device=self.device)
        self.timers = None            # optional {name: [(start_event, end_event), ...]} filled by estimate()
        self.lut = fast.decode_lut(cfg.vote_range, cfg.tr_num_bins, cfg.rot_num_bins).to(self.device)
        self.encoder_impl = "tc"      # "tc": tcgen05 3xTF32 encoder (csrc/encode_tc.cu); "simt": fp32 FFMA (csrc/fused.cu)
        self.timing = None            # optional cppf_timing_create() handle passed to cppf_pose_fused
        self._routed_scratch, self._routed_key = None, None
        self._pe_fused_ok = None
        self._rec_ring, self._rec_next = None, 0      # pinned pose records handed out round-robin (<= 1024 objects in flight)
        self._gen = None              # CUDA generator for the pair sampling of nocs/inference.py:177 (re-seeded per object)

    def _timed(self, name):
        est = self

        class _Ctx:
            def __enter__(self):
                if est.timers is not None:
                    self.a = torch.cuda.Event(enable_timing=True)
                    self.b = torch.cuda.Event(enable_timing=True)
                    self.a.record()

            def __exit__(self, *exc):
                if est.timers is not None:
                    self.b.record()
                    est.timers.setdefault(name, []).append((self.a, self.b))
        return _Ctx()

    # ------------------------------------------------------------------ stages
    @torch.no_grad()
    def point_features(self, pc, nrm):
        """nocs/inference.py:180-181.  The reference materialises torch.cdist (N x N) and takes topk of it; the
        fused path selects the same k neighbours from exact distances in-kernel (cppf_knn) and runs the SPRIN
        convolution in one kernel.  Neighbour sets can differ from cdist's only on near-ties of the k-th
        distance (cdist uses the less accurate |x|^2 + |y|^2 - 2 x.y form, SURVEY.md section 8a notes)."""
        if self.pe._fused_ok():
            return self.pe.encode_fused(pc, nrm)
        dist = torch.cdist(pc[None], pc[None])
        return self.pe(pc[None], nrm[None], dist)[0]

    @torch.no_grad()
    def estimate(self, pc_host: np.ndarray, nrm_host: np.ndarray, seed: int = 0, idxs=None, noise=None,
                 return_debug: bool = False):
        """pc_host, nrm_host: float32 [N,3] host arrays (pinned or not).  Returns a dict with
        RT [4,4], scales [3] (nocs/inference.py:336-339) and a 17-float `record`."""
        cfg, dev = self.cfg, self.device
        noise = noise or {}
        n = pc_host.shape[0]
        pc = torch.as_tensor(pc_host).to(dev, non_blocking=True)                    # :174-175
        nrm = torch.as_tensor(nrm_host).to(dev, non_blocking=True)
        if isinstance(pc_host, torch.Tensor) and pc_host.is_cuda:                   # cloud already resident in HBM
            corner = pc.min(0)[0]
            dims = voting.grid_dims(pc, corner, cfg.res)
        else:
            corner_np, dims = vote_grid_geometry(np.asarray(pc_host), cfg.res)      # :194-195
            corner = torch.from_numpy(corner_np).to(dev, non_blocking=True)
        if idxs is None and cfg.n_pairs > 0:                                        # :177
            g = torch.Generator(device=dev).manual_seed(seed)
            idxs = torch.randint(0, n, (cfg.n_pairs, 2), generator=g, device=dev, dtype=torch.int32)
        elif idxs is not None:
            idxs = torch.as_tensor(idxs).to(dev)
        feat = self.point_features(pc, nrm)

        # ---- first pass: translation heads only (columns 0:64, :183-188)
        B = cfg.tr_num_bins
        with self._timed("ppf_encode_pass1"):
            logits_tr = self.ppf._encode(pc, nrm, feat, idxs, None, cols=(0, 2 * B))
        mu_nu = torch.empty((logits_tr.shape[0], 2), dtype=torch.float32, device=dev)
        mu_nu[:, 0] = voting.sample_bins(logits_tr, 0, B, q=noise.get("q_mu"), u=noise.get("u_mu"), seed=seed, stream_id=0,
                                         div=B - 1, mul_a=2.0, mul_b=cfg.vote_range[0], sub=cfg.vote_range[0])
        mu_nu[:, 1] = voting.sample_bins(logits_tr, B, B, q=noise.get("q_nu"), u=noise.get("u_nu"), seed=seed, stream_id=1,
                                         div=B - 1, mul_a=cfg.vote_range[1])
        del logits_tr

        # ---- centre voting + argmax (:191-211)
        grid = torch.zeros(dims, dtype=torch.float32, device=dev)
        with self._timed("ppf_vote"):
            voting.ppf_vote(pc, mu_nu, idxs, grid, corner, cfg.res, cfg.num_rots, cfg.adaptive_voting)
        flat = voting.grid_argmax(grid)
        gyz = dims[1] * dims[2]
        cell = torch.stack([flat // gyz, (flat % gyz) // dims[2], flat % dims[2]], -1)[0]
        T_est = corner.double() + cell.double() * cfg.res                           # :209 float64 like numpy
        centre = T_est.float()

        # ---- back-vote filter + compaction (:216-231)
        _, mask = voting.backvote(pc, mu_nu, idxs, dims, corner, cfg.res, centre, 3 * cfg.res, cfg.num_rots,
                                  want_offsets=False)
        kept, cnt, pos = voting.compact_pairs(mask, idxs, n, want_pos=True)
        n_kept = int(cnt.item())                                                    # the one host sync
        kept, pos = kept[:n_kept], pos[:n_kept]

        def rows(t):            # injected per-pair noise is indexed by the ORIGINAL pair row: survivors take their rows
            return None if t is None else torch.as_tensor(t).to(dev)[pos].contiguous()
        out = {"T": T_est, "n_survivors": n_kept, "grid_dims": dims, "argmax": flat}
        if n_kept == 0:
            raise NoSurvivorsError("no pair voted for the winning centre (degenerate input)")

        # ---- second pass on the survivors (:236-256): rotation / aux / scale heads
        R0 = 2 * B
        RB = cfg.rot_num_bins
        heads = self.ppf._encode(pc, nrm, feat, kept, None, cols=(R0, self.ppf.out_dim - R0))
        preds_up_aux, preds_right_aux = heads[:, -5], heads[:, -4]
        log_scale = heads[:, -3:].mean(0)                                           # :335
        dirs = []
        for j, (col, aux, tag) in enumerate([(0, preds_up_aux, "up"), (RB, preds_right_aux, "right")]):
            if j == 1 and not cfg.regress_right:                                    # :260-261
                continue
            rot = voting.sample_bins(heads, col, RB, q=rows(noise.get(f"q_{tag}")), u=rows(noise.get(f"u_{tag}")), seed=seed,
                                     stream_id=2 + j, div=RB - 1, mul_a=float(np.float32(np.pi)))     # :250-256
            sub = kept
            if cfg.rot_subsample and n_kept > cfg.rot_subsample:                    # :277-281
                if noise.get("sub_key") is not None:        # injected shuffle: the survivors with the smallest keys
                    sel = torch.argsort(rows(noise["sub_key"]), stable=True)[:cfg.rot_subsample]
                else:
                    g = torch.Generator(device=dev).manual_seed(seed + 17 + j)
                    sel = torch.randperm(n_kept, generator=g, device=dev)[:cfg.rot_subsample]
                sub, rot_sub = kept[sel].contiguous(), rot[sel].contiguous()
            else:
                rot_sub = rot
            cand = voting.rot_vote(pc, rot_sub, sub, cfg.num_rots)                  # :265-275
            counts = voting.sphere_count(cand, self.sphere, self.cos_thr)           # :282-283
            best = torch.argmax(counts)                                             # :284 (first max)
            best_dir = self.sphere[best]
            # aux sign (:286-302)
            a_i, b_i = kept[:, 0].long(), kept[:, 1].long()
            ab = pc[a_i] - pc[b_i]
            abn = ab / (ab.pow(2).sum(-1).sqrt() + 1e-7)[:, None]
            pn = nrm[a_i]
            pn = torch.where(((pn * abn).sum(-1) < 0)[:, None], -pn, pn)
            target = ((pn * best_dir).sum(-1) > 0).float()
            up_loss = torch.nn.functional.binary_cross_entropy_with_logits(aux, target)
            down_loss = torch.nn.functional.binary_cross_entropy_with_logits(aux, 1.0 - target)
            sign = torch.where(down_loss < up_loss, -1.0, 1.0)
            dirs.append((best, sign))
            if return_debug:
                out[f"counts_{tag}"] = counts
        # ---- pose assembly (:305-339), on the host in float64 like the reference
        rec = torch.cat([torch.stack([d[0].double() for d in dirs]), torch.stack([d[1].double() for d in dirs]),
                         T_est, log_scale.double()]).cpu().numpy()                  # single D2H
        k = len(dirs)
        up = self.sphere_np[int(rec[0])] * rec[k]
        if cfg.regress_right:
            right = self.sphere_np[int(rec[1])] * rec[k + 1]
            right = right - np.dot(up, right) * up
            right /= (np.linalg.norm(right) + 1e-9)
        else:
            right = np.array([0, -up[2], up[1]])
            right /= (np.linalg.norm(right) + 1e-9)
        if np.linalg.norm(right) < 1e-7:
            right = np.array([up[1], -up[0], 0.0])
            right /= (np.linalg.norm(right) + 1e-9)
        R = np.stack([np.cross(up, right), up, right], -1) if cfg.z_right else np.stack([right, up, np.cross(right, up)], -1)
        T = rec[2 * k:2 * k + 3]
        pred_scale = np.exp(rec[2 * k + 3:2 * k + 6].astype(np.float32)) * np.asarray(cfg.scale_mean) * cfg.scale_mul
        sn = np.linalg.norm(pred_scale)
        RT = np.eye(4, dtype=np.float32)
        RT[:3, :3] = R * sn
        RT[:3, 3] = T
        out.update(RT=RT, scales=(pred_scale / sn).astype(np.float32), up=up, right=right, T_host=T,
                   pred_scale=pred_scale,
                   record=np.concatenate([[0.0, float(n_kept)], pred_scale, R.reshape(-1), T]).astype(np.float32))
        if return_debug:
            out.update(grid=grid, mu_nu=mu_nu, kept=kept, idxs=idxs, feat=feat)
        return out


    # ------------------------------------------------------------------ fused path
    def _pose_from_record(self, rec, n_dirs, dims):
        """Host tail of nocs/inference.py:305-339 from one small device->host record:
        rec = [flat, best_up, (best_right), scale_sum x3, count, S_up, S_right, corner x3] (float64);
        dims = the vote-grid dims of THIS object."""
        cfg = self.cfg
        flat = int(rec[0])
        bests = [int(rec[1 + j]) for j in range(n_dirs)]
        st = rec[1 + n_dirs:7 + n_dirs]
        corner = rec[7 + n_dirs:10 + n_dirs]
        cell = np.array(np.unravel_index(flat, dims))
        T = corner + cell * cfg.res                                                 # :209
        if st[3] <= 0:
            raise NoSurvivorsError("no pair voted for the winning centre (degenerate input)")
        cnt = st[3]
        up = self.sphere_np[bests[0]] * (-1.0 if st[4] < 0 else 1.0)                # :299-302
        if cfg.regress_right:
            right = self.sphere_np[bests[1]] * (-1.0 if st[5] < 0 else 1.0)
            right = right - np.dot(up, right) * up
            right /= (np.linalg.norm(right) + 1e-9)
        else:
            right = np.array([0, -up[2], up[1]])
            right /= (np.linalg.norm(right) + 1e-9)
        if np.linalg.norm(right) < 1e-7:
            right = np.array([up[1], -up[0], 0.0])
            right /= (np.linalg.norm(right) + 1e-9)
        R = np.stack([np.cross(up, right), up, right], -1) if cfg.z_right else np.stack([right, up, np.cross(right, up)], -1)
        pred_scale = np.exp((st[:3] / cnt).astype(np.float32)) * np.asarray(cfg.scale_mean) * cfg.scale_mul    # :335
        sn = np.linalg.norm(pred_scale)
        RT = np.eye(4, dtype=np.float32)
        RT[:3, :3] = R * sn
        RT[:3, 3] = T
        return dict(RT=RT, scales=(pred_scale / sn).astype(np.float32), up=up, right=right, T_host=T, pred_scale=pred_scale,
                    n_survivors=int(st[3]), argmax_flat=flat, best_bins=bests, grid_dims=tuple(dims),
                    record=np.concatenate([[0.0, st[3]], pred_scale, R.reshape(-1), T]).astype(np.float32))

    # ------------------------------------------------------------------ one call per object
    def _onecall_ok(self):
        if self._pe_fused_ok is None:
            self._pe_fused_ok = bool(self.pe._fused_ok())
        return (self.encoder_impl == "tc" and self._pe_fused_ok and self.cfg.num_rots <= 72 and self.ppf.out_dim == 141)

    def _workspace(self, n, n_pairs, max_cells, routed_max_cells):
        """One workspace per (device, stream), shared by every estimator (categories differ only in weights) and
        grown on demand: work on one stream is ordered, so successive objects can reuse it."""
        nb = _lib.lib().cppf_pose_workspace_bytes(n, n_pairs, self.cfg.knn, max_cells, routed_max_cells, self.cfg.num_rots,
                                                  self.sphere.shape[0], int(self.cfg.rot_subsample or 0))
        slot = (self.device.index or 0, torch.cuda.current_stream(self.device).cuda_stream)
        ws = _WORKSPACES.get(slot)
        if ws is None or ws.numel() < nb:
            _WORKSPACES[slot] = None
            ws = _WORKSPACES[slot] = torch.empty(nb, dtype=torch.uint8, device=self.device)
        return ws

    def grid_capacity(self, pc_in):
        """(max_cells, routed_max_cells) for cppf_pose_fused from the cloud's bounding box (nocs/inference.py:194-195):
        the exact cell count in the slot of the vote kernel that will run, or None when the grid needs the
        global-reduction kernel of the staged path.  Costs one small device->host copy for a CUDA tensor."""
        if isinstance(pc_in, torch.Tensor) and pc_in.is_cuda:
            lo = pc_in.min(0)[0]
            dims = voting.grid_dims(pc_in, lo, self.cfg.res)
        else:
            _, dims = vote_grid_geometry(np.asarray(pc_in), self.cfg.res)
        cells = dims[0] * dims[1] * dims[2]
        if fast.vote_fits_private(dims):
            return cells, 0
        if fast.vote_routed_supported(dims):
            return 1, cells
        return None

    @torch.no_grad()
    def enqueue_fused(self, pc_in, nrm_in, seed: int = 0, idxs=None, uniforms=None, inject_bins=None, max_cells=None,
                      routed_max_cells=0, record_host=None, device_pairs: bool = False):
        """Enqueue the whole per-object path (cppf_pose_fused) on the current stream and return a PendingPose;
        nothing here waits for the GPU when max_cells is given.  pc_in / nrm_in: float32 [N,3], host (numpy or
        pinned torch) or CUDA.  max_cells / routed_max_cells: capacity of the vote grid in the shared-memory kernel /
        in the routed-slab kernel (see grid_capacity; default: derived from the cloud's bounding box).
        device_pairs: draw the cfg.n_pairs random pairs of nocs/inference.py:177 inside the library (Philox keyed by
        `seed`) instead of with torch.randint -- the pairs cppf_pose_batch / enqueue_batch use."""
        cfg, dev = self.cfg, self.device
        L = _lib.lib()
        n = pc_in.shape[0]
        if max_cells is None:
            cap = self.grid_capacity(pc_in)
            if cap is None:
                raise RuntimeError("vote grid too large for cppf_pose_fused; use estimate_fused(staged=True)")
            max_cells, routed_max_cells = cap
        if not self._onecall_ok():
            raise RuntimeError("cppf_pose_fused needs the tcgen05 encoder and the reference PointEncoder / head "
                               "configuration; use estimate_fused(staged=True)")
        pc = torch.as_tensor(pc_in).to(dev, torch.float32, non_blocking=True).contiguous()
        nrm = torch.as_tensor(nrm_in).to(dev, torch.float32, non_blocking=True).contiguous()
        sample_dev = bool(device_pairs and idxs is None and cfg.n_pairs > 0)
        if sample_dev:
            pass
        elif idxs is None and cfg.n_pairs > 0:                                      # nocs/inference.py:177
            if self._gen is None:
                self._gen = torch.Generator(device=dev)
            idxs = torch.randint(0, n, (cfg.n_pairs, 2), generator=self._gen.manual_seed(seed), device=dev, dtype=torch.int32)
        elif idxs is not None:
            idxs = torch.as_tensor(idxs).to(dev).contiguous()
            assert idxs.dtype in (torch.int32, torch.int64)
        n_pairs = cfg.n_pairs if sample_dev else (n * n if idxs is None else idxs.shape[0])
        if uniforms is not None:
            assert uniforms.shape == (n_pairs, 4) and uniforms.is_contiguous() and uniforms.dtype == torch.float32
        if inject_bins is not None:
            assert inject_bins.dtype == torch.uint8 and inject_bins.is_contiguous() and inject_bins.shape[0] == n_pairs
        ws = self._workspace(n, n_pairs, int(max_cells), int(routed_max_cells))
        rec = torch.empty(16, dtype=torch.float64, device=dev)
        a_sample = int(sample_dev)
        a = _lib.PoseArgs()
        a.struct_bytes = C.sizeof(_lib.PoseArgs)
        a.pc, a.nrm = pc.data_ptr(), nrm.data_ptr()
        a.idx = idxs.data_ptr() if idxs is not None else None
        a.pe_blob, a.tc_blob = self.pe.pe_blob(dev).data_ptr(), self.ppf.tc_blob(dev).data_ptr()
        a.lut, a.sphere = self.lut.data_ptr(), self.sphere.data_ptr()
        a.uniforms = uniforms.data_ptr() if uniforms is not None else None
        a.inject_bins = inject_bins.data_ptr() if inject_bins is not None else None
        a.workspace, a.record, a.timing = ws.data_ptr(), rec.data_ptr(), self.timing
        a.n_pairs, a.workspace_bytes = n_pairs, ws.numel()
        a.rot_subsample, a.seed = int(cfg.rot_subsample or 0), int(seed)
        a.n_points, a.idx_is_64 = n, int(idxs is not None and idxs.dtype == torch.int64)
        a.knn, a.n_rots, a.adaptive, a.regress_right = cfg.knn, cfg.num_rots, int(cfg.adaptive_voting), int(cfg.regress_right)
        a.n_sphere, a.inject_cols = self.sphere.shape[0], (inject_bins.shape[1] if inject_bins is not None else 0)
        a.max_cells, a.routed_max_cells, a.sample_pairs = int(max_cells), int(routed_max_cells), a_sample
        a.res, a.tol, a.cos_thr, a.res_host = float(cfg.res), float(3 * cfg.res), self.cos_thr, float(cfg.res)
        with torch.cuda.device(dev):
            _lib.check(L.cppf_pose_fused(C.byref(a), torch.cuda.current_stream(dev).cuda_stream), "cppf_pose_fused")
        if record_host is None:         # a slot of a pinned ring (a fresh pinned allocation per object costs a cudaHostAlloc)
            if self._rec_ring is None:
                self._rec_ring = torch.empty((1024, 16), dtype=torch.float64, pin_memory=True)
            record_host = self._rec_ring[self._rec_next % 1024]
            self._rec_next += 1
        record_host.copy_(rec, non_blocking=True)
        done = torch.cuda.Event()
        done.record()
        return PendingPose(self, record_host, done, 2 if cfg.regress_right else 1, keep=(pc, nrm, idxs, uniforms, inject_bins, rec))

    def _pose_from_record16(self, r, n_dirs):
        if r[15] != 0:
            raise RuntimeError("vote grid larger than the capacity given to cppf_pose_fused (status %d)" % int(r[15]))
        old = np.concatenate([[r[0]], r[1:1 + n_dirs], r[3:9], r[9:12]])
        return self._pose_from_record(old, n_dirs, tuple(int(v) for v in r[12:15]))

    def _records17_from_records16(self, r):
        """[m,16] pose records of this category -> [m,17] float32 gather records (sunrgbd/inference.py:287:
        class_id, score = survivor count, scale x3, R x9 row-major, T x3), column-wise: the host tail of
        nocs/inference.py:305-339 for a whole batch.  Rows without survivors / over capacity get score 0 and identity R."""
        cfg = self.cfg
        m = r.shape[0]
        ok = (r[:, 15] == 0) & (r[:, 6] > 0)
        gy, gz = np.maximum(r[:, 13], 1).astype(np.int64), np.maximum(r[:, 14], 1).astype(np.int64)
        flat = r[:, 0].astype(np.int64)
        cell = np.stack([flat // (gy * gz), (flat % (gy * gz)) // gz, flat % gz], -1)
        T = r[:, 9:12] + cell * cfg.res                                             # :209
        up = self.sphere_np[np.clip(r[:, 1].astype(np.int64), 0, len(self.sphere_np) - 1)] * np.where(r[:, 7] < 0, -1.0, 1.0)[:, None]
        if cfg.regress_right:
            right = self.sphere_np[np.clip(r[:, 2].astype(np.int64), 0, len(self.sphere_np) - 1)] * np.where(r[:, 8] < 0, -1.0, 1.0)[:, None]
            right = right - (up * right).sum(-1, keepdims=True) * up
        else:
            right = np.stack([np.zeros(m), -up[:, 2], up[:, 1]], -1)
        right = right / (np.linalg.norm(right, axis=-1, keepdims=True) + 1e-9)
        bad = np.linalg.norm(right, axis=-1) < 1e-7
        if bad.any():
            alt = np.stack([up[:, 1], -up[:, 0], np.zeros(m)], -1)
            right[bad] = alt[bad] / (np.linalg.norm(alt[bad], axis=-1, keepdims=True) + 1e-9)
        if cfg.z_right:
            R = np.stack([np.cross(up, right), up, right], -1)
        else:
            R = np.stack([right, up, np.cross(right, up)], -1)
        cnt = np.where(ok, r[:, 6], 1.0)
        scale = np.exp((r[:, 3:6] / cnt[:, None]).astype(np.float32)) * np.asarray(cfg.scale_mean) * cfg.scale_mul     # :335
        out = np.zeros((m, 17), np.float32)
        out[:, 1] = np.where(ok, r[:, 6], 0.0)
        out[:, 2:5] = scale
        out[:, 5:14] = np.where(ok[:, None], R.reshape(m, 9), np.eye(3).reshape(1, 9))
        out[:, 14:17] = T
        return out

    @torch.no_grad()
    def estimate_fused(self, pc_in, nrm_in, seed: int = 0, idxs=None, uniforms=None, return_debug: bool = False,
                       sync: bool = True, inject_bins=None, staged: bool = False, max_cells=None, routed_max_cells=0,
                       device_pairs: bool = False):
        """Same pose as `estimate`, through the fused kernels: logits, (mu,nu) floats and rotation
        candidates never reach HBM; a single 128-byte record comes back to the host.  By default the whole
        object is ONE library call (cppf_pose_fused); staged=True / return_debug=True run it kernel by kernel
        (same kernels) and expose the intermediates.  sync=False returns a PendingPose (.result())."""
        if not (staged or return_debug) and self._onecall_ok():
            cap = (max_cells, routed_max_cells) if max_cells is not None else self.grid_capacity(pc_in)
            if cap is not None:
                pend = self.enqueue_fused(pc_in, nrm_in, seed=seed, idxs=idxs, uniforms=uniforms, inject_bins=inject_bins,
                                          max_cells=cap[0], routed_max_cells=cap[1], device_pairs=device_pairs)
                return pend.result() if sync else pend
        return self._estimate_fused_staged(pc_in, nrm_in, seed=seed, idxs=idxs, uniforms=uniforms, return_debug=return_debug,
                                           sync=sync, inject_bins=inject_bins)

    @torch.no_grad()
    def _estimate_fused_staged(self, pc_in, nrm_in, seed: int = 0, idxs=None, uniforms=None, return_debug: bool = False,
                               sync: bool = True, inject_bins=None):
        cfg, dev = self.cfg, self.device
        n = pc_in.shape[0]
        pc = torch.as_tensor(pc_in).to(dev, non_blocking=True)
        nrm = torch.as_tensor(nrm_in).to(dev, non_blocking=True)
        if isinstance(pc_in, torch.Tensor) and pc_in.is_cuda:
            corner = pc.min(0)[0]
            dims = voting.grid_dims(pc, corner, cfg.res)
        else:
            corner_np, dims = vote_grid_geometry(np.asarray(pc_in), cfg.res)
            corner = torch.from_numpy(corner_np).to(dev, non_blocking=True)
        if idxs is None and cfg.n_pairs > 0:
            g = torch.Generator(device=dev).manual_seed(seed)
            idxs = torch.randint(0, n, (cfg.n_pairs, 2), generator=g, device=dev, dtype=torch.int32)
        elif idxs is not None:
            idxs = torch.as_tensor(idxs).to(dev).contiguous()
        with self._timed("point_encoder"):
            feat = self.point_features(pc, nrm)
        table = self.ppf.tc_preproject(feat) if self.encoder_impl == "tc" else self.ppf.preproject(feat)
        heads = fast.HEAD_TR | fast.HEAD_UP | fast.HEAD_TAIL | (fast.HEAD_RIGHT if cfg.regress_right else 0)
        with self._timed("encode_sample"):
            bins, tail = fast.encode_sample(self.ppf, pc, nrm, table, idxs, heads=heads, uniforms=uniforms, seed=seed,
                                            impl=self.encoder_impl)
        if inject_bins is not None:           # benchmark aid: vote load of a trained network (SURVEY.md 8d i)
            bins[:, :inject_bins.shape[1]] = inject_bins
        grid = torch.zeros(dims, dtype=torch.float32, device=dev)
        with self._timed("vote"):
            if fast.vote_fits_private(dims) and cfg.num_rots <= 72:
                fast.vote_fast(pc, idxs, grid, corner, cfg.res, bins=bins, lut=self.lut, n_rots=cfg.num_rots,
                               adaptive=cfg.adaptive_voting)
            elif fast.vote_routed_supported(dims) and cfg.num_rots <= 72:     # up to 8 shared-memory slabs (64^3)
                if self._routed_scratch is None or self._routed_key != (n, idxs is None, dims):
                    n_pairs = n * n if idxs is None else idxs.shape[0]
                    nb = _lib.lib().cppf_vote_routed_scratch_bytes(n_pairs, cfg.num_rots, *dims)
                    self._routed_scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
                    self._routed_key = (n, idxs is None, dims)
                fast.vote_routed(pc, idxs, grid, corner, cfg.res, bins=bins, lut=self.lut, n_rots=cfg.num_rots,
                                 adaptive=cfg.adaptive_voting, scratch=self._routed_scratch)
            else:                                   # larger still (scene-scale grids): global fp32 reductions
                b = bins.long()
                mu_nu = torch.stack([self.lut[b[:, 0]], self.lut[32 + b[:, 1]]], -1).contiguous()
                voting.ppf_vote(pc, mu_nu, idxs, grid, corner, cfg.res, cfg.num_rots, cfg.adaptive_voting)
        flat = voting.grid_argmax(grid)
        with self._timed("backvote"):
            mask = fast.backvote_bins(pc, bins, self.lut, idxs, dims, corner, flat, cfg.res, 3 * cfg.res, cfg.num_rots)
            _, cnt, pos = voting.compact_pairs(mask, idxs, n, want_pos=True, want_idx=False)
        bests = []
        with self._timed("rot_hist"):
            for j in range(2 if cfg.regress_right else 1):
                counts = fast.rot_hist(pc, bins, self.lut, idxs, pos, cnt, self.sphere, which=j, n_rots=cfg.num_rots,
                                       max_samples=cfg.rot_subsample or (1 << 40), offset_seed=seed * 7919 + j,
                                       thr=self.cos_thr)
                bests.append(voting.grid_argmax(counts))
            stats = fast.survivor_stats(pc, nrm, tail, idxs, pos, cnt, self.sphere, bests[0],
                                        bests[1] if cfg.regress_right else None)
        rec_dev = torch.cat([flat.double()] + [b.double() for b in bests] + [stats, corner.double()])
        nd = len(bests)
        if not sync:
            host = torch.empty(rec_dev.shape, dtype=torch.float64, pin_memory=True)
            host.copy_(rec_dev, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
            return PendingPose(self, host, done, nd, keep=(rec_dev,), staged=True, dims=dims)
        out = self._pose_from_record(rec_dev.cpu().numpy(), nd, dims)
        if return_debug:
            out.update(grid=grid, bins=bins, tail=tail, mask=mask, pos=pos, count=cnt, feat=feat, idxs=idxs, table=table)
        return out


_ARGS_DTYPE = None
_PINNED = {}              # (device index, tag) -> [pinned tensor, event of its last use]: cudaHostAlloc costs milliseconds


def _pinned(dev, tag, shape, dtype):
    """A cached page-locked host buffer of at least `shape` elements (grow-only).  Before it is handed out again the
    previous user's stream work is waited for (the buffer is the source / target of an asynchronous copy)."""
    key = (dev.index or 0, tag)
    need = int(np.prod(shape))
    ent = _PINNED.get(key)
    if ent is None or ent[0].numel() < need or ent[0].dtype != dtype:
        ent = _PINNED[key] = [torch.empty(max(need, 1), dtype=dtype, pin_memory=True), None]
    if ent[1] is not None:
        ent[1].synchronize()
    return ent, ent[0][:need].view(*shape)



def _args_dtype():
    """numpy mirror of struct cppf_pose_args (include/cppf_b200.h), so that a whole batch of argument blocks is filled
    column-wise; checked against the compiled layout."""
    global _ARGS_DTYPE
    if _ARGS_DTYPE is None:
        ct = {C.c_int64: "<i8", C.c_uint64: "<u8", C.c_void_p: "<u8", C.c_int: "<i4", C.c_float: "<f4", C.c_double: "<f8"}
        dt = np.dtype({"names": [f for f, _ in _lib.PoseArgs._fields_],
                       "formats": [ct[t] for _, t in _lib.PoseArgs._fields_],
                       "offsets": [getattr(_lib.PoseArgs, f).offset for f, _ in _lib.PoseArgs._fields_],
                       "itemsize": C.sizeof(_lib.PoseArgs)})
        assert dt.itemsize == _lib.lib().cppf_pose_args_bytes()
        _ARGS_DTYPE = dt
    return _ARGS_DTYPE


class PendingBatch:
    """A batch enqueued by ONE cppf_pose_batch call; .results() waits for the [n,16] record block (one device->host copy)
    and runs the host tail of nocs/inference.py:305-339 for every object."""

    def __init__(self, ests, records_host, done, keep):
        self.ests, self.records_host, self.done, self._keep = ests, records_host, done, keep

    def results(self, on_error="raise"):
        """-> list of pose dicts.  on_error="none": objects whose pose could not be formed (no survivors / grid over
        capacity) yield None instead of raising."""
        self.done.synchronize()
        self._keep = ()
        recs = self.records_host.numpy()
        out = []
        for est, r in zip(self.ests, recs):
            try:
                out.append(est._pose_from_record16(r, 2 if est.cfg.regress_right else 1))
            except RuntimeError:
                if on_error == "raise":
                    raise
                out.append(None)
        return out

    def records17(self):
        """The gather records (sunrgbd/inference.py:287 layout) of the whole batch as one float32 [n,17] array, computed
        column-wise (no per-object Python): what shard.gather_records sends."""
        self.done.synchronize()
        self._keep = ()
        recs = self.records_host.numpy()
        out = np.zeros((len(self.ests), 17), np.float32)
        groups = {}
        for i, e in enumerate(self.ests):
            groups.setdefault(id(e), (e, []))[1].append(i)
        for est, rows in groups.values():
            rows = np.asarray(rows)
            out[rows] = est._records17_from_records16(recs[rows])
        return out


@torch.no_grad()
def enqueue_batch(items, n_streams: int = 4, n_threads: int = 4, n_pairs=None, inject_bins=None, capacities=None):
    """The object loop of nocs/inference.py:120-129 as ONE library call (cppf_pose_batch).
    items = [(estimator, pc, normals, seed), ...]: float32 [N_i,3] clouds, host (numpy / pinned torch) or CUDA; objects may
    belong to different categories (estimators).  Host clouds are packed into one pinned block and copied with ONE
    host->device copy; the point pairs of nocs/inference.py:177 are drawn on the device (cfg.n_pairs > 0; 0 = all N^2
    pairs); all records come back with ONE device->host copy.  inject_bins: optional list (one uint8 CUDA tensor [P_i, c]
    or None per object) overwriting the sampled bins (benchmark aid, see enqueue_fused).  capacities: optional list of
    (max_cells, routed_max_cells) per object (see grid_capacity; deriving them from a CUDA-resident cloud costs a device->host
    copy per object).  -> PendingBatch."""
    if not items:
        raise ValueError("empty batch")
    dev = items[0][0].device
    L = _lib.lib()
    n_obj = len(items)
    n_streams = max(1, min(int(n_streams), 16, n_obj))
    n_threads = max(1, min(int(n_threads), n_streams))
    ns = np.array([it[1].shape[0] for it in items], np.int64)
    # ---- clouds: one pinned staging block, one H2D (CUDA clouds are used in place)
    on_dev = [isinstance(it[1], torch.Tensor) and it[1].is_cuda for it in items]
    host_rows = int(sum(n for n, d in zip(ns, on_dev) if not d))
    keep = []
    ptr_pc, ptr_nrm = np.zeros(n_obj, np.uint64), np.zeros(n_obj, np.uint64)
    lo_hi = np.zeros((n_obj, 2, 3), np.float32)                      # bounding boxes of the host clouds (nocs/inference.py:194)
    if host_rows:
        stage_ent, stage = _pinned(dev, "clouds", (2 * host_rows, 3), torch.float32)
        sn = stage.numpy()
        off = 0
        offs = []
        for k_, ((est, pc, nrm, _), d, n) in enumerate(zip(items, on_dev, ns)):
            if d:
                offs.append(-1)
                continue
            blk = sn[off:off + n]
            blk[:] = pc.numpy() if isinstance(pc, torch.Tensor) else pc
            sn[off + n:off + 2 * n] = nrm.numpy() if isinstance(nrm, torch.Tensor) else nrm
            if capacities is None:
                bt = np.ascontiguousarray(blk.T)                 # [3, N]: a reduction along axis 0 of [N, 3] is ~25x slower
                lo_hi[k_, 0], lo_hi[k_, 1] = bt.min(1), bt.max(1)
            offs.append(off)
            off += 2 * n
        dstage = stage.to(dev, non_blocking=True)
        stage_ent[1] = torch.cuda.Event()
        stage_ent[1].record()
        keep += [dstage]
        base = dstage.data_ptr()
        for i, (o, n) in enumerate(zip(offs, ns)):
            if o >= 0:
                ptr_pc[i], ptr_nrm[i] = base + 12 * o, base + 12 * (o + int(n))
    for i, ((est, pc, nrm, _), d) in enumerate(zip(items, on_dev)):
        if d:
            pcd, nd = pc.to(torch.float32).contiguous(), nrm.to(torch.float32).contiguous()
            keep += [pcd, nd]
            ptr_pc[i], ptr_nrm[i] = pcd.data_ptr(), nd.data_ptr()
    # ---- grid capacities (nocs/inference.py:194-195) and pair counts
    if not all(it[0]._onecall_ok() for it in {id(it[0]): it for it in items}.values()):
        raise RuntimeError("cppf_pose_batch needs the tcgen05 encoder and the reference PointEncoder / head configuration")
    if capacities is not None:
        caps = [tuple(c) for c in capacities]
    else:
        # host clouds: grid dims for the whole batch at once (float32 like numpy at :195); CUDA clouds: one round trip each
        res_v = np.array([it[0].cfg.res for it in items], np.float64)
        dims_v = ((lo_hi[:, 1] - lo_hi[:, 0]) / res_v[:, None].astype(np.float32)).astype(np.int32) + 1
        cells_v = dims_v.astype(np.int64).prod(1)
        priv_max = L.cppf_vote_private_max_cells()
        caps = []
        for k_, ((est, pc, nrm, _), d) in enumerate(zip(items, on_dev)):
            if d:
                cap = est.grid_capacity(pc)
            elif cells_v[k_] <= priv_max:
                cap = (int(cells_v[k_]), 0)
            else:
                cap = (1, int(cells_v[k_])) if fast.vote_routed_supported(tuple(int(v) for v in dims_v[k_])) else None
            if cap is None:
                raise RuntimeError("vote grid too large for cppf_pose_batch; use estimate_fused(staged=True) for this object")
            caps.append(cap)
    caps = np.asarray(caps, np.int64)
    pairs = np.array([(it[0].cfg.n_pairs if n_pairs is None else n_pairs) for it in items], np.int64)
    dense = pairs <= 0
    pairs = np.where(dense, ns * ns, pairs)
    # ---- one workspace per worker stream, sized for the largest object it serves
    ws = []
    ws_bytes = {}
    for s_ in range(n_streams):
        rows = np.arange(s_, n_obj, n_streams)
        nb = 0
        for i in rows:
            cfg = items[i][0].cfg
            key = (int(ns[i]), 0 if dense[i] else int(pairs[i]), cfg.knn, int(caps[i, 0]), int(caps[i, 1]), cfg.num_rots,
                   items[i][0].sphere.shape[0], int(cfg.rot_subsample or 0))
            if key not in ws_bytes:
                ws_bytes[key] = L.cppf_pose_workspace_bytes(*key)
            nb = max(nb, ws_bytes[key])
        slot = ("batch", dev.index or 0, s_)
        w = _WORKSPACES.get(slot)
        if w is None or w.numel() < nb:
            _WORKSPACES[slot] = None
            w = _WORKSPACES[slot] = torch.empty(nb, dtype=torch.uint8, device=dev)
        ws.append(w)
    rec = torch.empty((n_obj, 16), dtype=torch.float64, device=dev)
    # ---- argument blocks, filled column-wise per category
    a = np.zeros(n_obj, _args_dtype())
    a["struct_bytes"] = a.dtype.itemsize
    a["pc"], a["nrm"] = ptr_pc, ptr_nrm
    a["n_points"], a["n_pairs"] = ns, pairs
    a["sample_pairs"] = (~dense).astype(np.int32)
    a["max_cells"], a["routed_max_cells"] = caps[:, 0], caps[:, 1]
    a["seed"] = np.array([int(it[3]) for it in items], np.uint64)
    a["record"] = rec.data_ptr() + 128 * np.arange(n_obj, dtype=np.uint64)
    wsp = np.array([w.data_ptr() for w in ws], np.uint64)
    wsb = np.array([w.numel() for w in ws], np.int64)
    a["workspace"], a["workspace_bytes"] = wsp[np.arange(n_obj) % n_streams], wsb[np.arange(n_obj) % n_streams]
    if inject_bins is not None:
        for i, t in enumerate(inject_bins):
            if t is not None:
                assert t.dtype == torch.uint8 and t.is_contiguous() and t.shape[0] == int(pairs[i])
                a["inject_bins"][i], a["inject_cols"][i] = t.data_ptr(), t.shape[1]
        keep.append(list(inject_bins))
    groups = {}
    for i, it in enumerate(items):
        groups.setdefault(id(it[0]), (it[0], []))[1].append(i)
    for est, rows in groups.values():
        cfg = est.cfg
        rows = np.asarray(rows)
        a["pe_blob"][rows], a["tc_blob"][rows] = est.pe.pe_blob(dev).data_ptr(), est.ppf.tc_blob(dev).data_ptr()
        a["lut"][rows], a["sphere"][rows] = est.lut.data_ptr(), est.sphere.data_ptr()
        a["rot_subsample"][rows] = int(cfg.rot_subsample or 0)
        a["knn"][rows], a["n_rots"][rows] = cfg.knn, cfg.num_rots
        a["adaptive"][rows], a["regress_right"][rows] = int(cfg.adaptive_voting), int(cfg.regress_right)
        a["n_sphere"][rows] = est.sphere.shape[0]
        a["res"][rows], a["tol"][rows], a["cos_thr"][rows] = float(cfg.res), float(3 * cfg.res), est.cos_thr
        a["res_host"][rows] = float(cfg.res)
    with torch.cuda.device(dev):
        _lib.check(L.cppf_pose_batch(a.ctypes.data, n_obj, n_streams, n_threads, torch.cuda.current_stream(dev).cuda_stream),
                   "cppf_pose_batch")
    # page-locked target of the one D2H copy, sized in steps of 256 objects so that torch's caching host allocator hands the
    # block of an earlier batch back instead of calling cudaHostAlloc (milliseconds) for every new batch size
    host = torch.empty((-(-n_obj // 256) * 256, 16), dtype=torch.float64, pin_memory=True)[:n_obj]
    host.copy_(rec, non_blocking=True)
    done = torch.cuda.Event()
    done.record()
    return PendingBatch([it[0] for it in items], host, done, keep + [rec, a, ws])


def estimate_many(items, sync: bool = True, batched: bool = True, n_streams: int = 4, n_threads: int = 4):
    """Batch of objects, possibly of different categories: items = [(estimator, pc, normals, seed), ...]
    (the reference loops over them one at a time and picks the category's weights, nocs/inference.py:120-129).
    batched=True: ONE cppf_pose_batch call (objects overlap on worker streams; pairs drawn on the device).
    batched=False: one cppf_pose_fused call per object on the current stream, pairs drawn by torch.randint.
    Either way every object is enqueued before the first record is read.  Returns the list of pose dicts (sync=True),
    or a PendingBatch / list of PendingPose."""
    if batched and all(it[0]._onecall_ok() for it in items):
        try:
            pend = enqueue_batch(items, n_streams=n_streams, n_threads=n_threads)
            return pend.results() if sync else pend
        except RuntimeError as e:           # a vote grid too large for the one-call path: per-object calls (staged fallback)
            if "too large" not in str(e):
                raise
    pend = [est.estimate_fused(pc, nrm, seed=seed, sync=False) for est, pc, nrm, seed in items]
    return [p.result() for p in pend] if sync else pend
