"""Keyframe selection of the SLAM front end on the GPU: the overlap score of
must3r/slam/model.py:62-91 (`get_overlap_score`), the decision of :123-128 (`choose_keyframe_from_overlap`) and the
nearest-neighbour searchers of must3r/slam/nns.py (`get_searcher`, `KDTree_scipy`, `QuandrantSearcher`), with the same names,
arguments and results - but the point database stays on the device and the per-frame query is a brute-force scan kernel
(`m3r_nn_min_dist`) instead of a scipy KD-tree on the CPU (`.cpu().numpy()` of every prediction + `KDTree.query`).
Results are torch tensors on the query's device; the score itself is a Python float (the caller branches on it).
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from .. import _lib


def nn_min_dist(queries: torch.Tensor, db: torch.Tensor) -> torch.Tensor:
    """[Q,3], [P,3] fp32 CUDA -> [Q] distance to the nearest database point (inf if P == 0)."""
    if not queries.is_cuda:
        raise RuntimeError("must3r_b200 keyframe search needs CUDA tensors (no CPU fallback)")
    q = queries.reshape(-1, 3).float().contiguous()
    d = db.reshape(-1, 3).float().contiguous()
    out = torch.empty((q.shape[0],), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(_lib.lib().m3r_nn_min_dist(C.c_void_p(q.data_ptr()), q.shape[0], C.c_void_p(d.data_ptr()) if d.numel() else None,
                                              d.shape[0], C.c_void_p(out.data_ptr()),
                                              C.c_void_p(torch.cuda.current_stream(q.device).cuda_stream)), "nn_min_dist")
    return out


def get_quadrant_id(rays: torch.Tensor, quadrant_divider: int = 4, eps: float = 1e-5) -> torch.Tensor:
    """must3r/slam/tools.py:9-31: viewing-direction quadrant of every ray (spherical coordinates, quantised)."""
    rays = rays / rays.norm(dim=-1, keepdim=True).clip(eps)
    thetas = (torch.acos(rays[:, -1]) / math.pi).clip(eps, 1 - eps)
    phis = (torch.atan2(rays[:, 1], rays[:, 0]) / math.pi).clip(-1 + eps, 1 - eps)
    theta_idx = torch.floor(thetas * quadrant_divider).long()
    phis_idx = torch.floor(phis * quadrant_divider).long() + quadrant_divider
    return theta_idx + phis_idx * quadrant_divider


class DeviceNN:
    """must3r/slam/nns.py:40-62 (`KDTree_scipy`): `add_pts` grows the database, `query` returns nearest distances."""

    def __init__(self):
        self.chunks, self.n = [], 0
        self._flat = None

    def add_pts(self, pts, **kw):
        p = pts.reshape(-1, 3).float()
        if p.shape[0]:
            self.chunks.append(p)
            self.n += p.shape[0]
            self._flat = None

    def _db(self, like):
        if self._flat is None:
            self._flat = torch.cat(self.chunks) if self.chunks else torch.zeros((0, 3), dtype=torch.float32, device=like.device)
            self.chunks = [self._flat] if self.n else []
        return self._flat

    def query(self, pts, **kw):
        return nn_min_dist(pts, self._db(pts))


class QuadrantSearcher:
    """must3r/slam/nns.py:65-95 (`QuandrantSearcher`): one database per viewing-direction quadrant; a query point only
    sees database points observed from a similar direction."""

    def __init__(self, method="quadrant_x4-kdtree-scipy"):
        self.quadrant_divider = int(method.split('quadrant_x')[-1].split('-')[0])
        self.search_structs = [DeviceNN() for _ in range(2 * self.quadrant_divider ** 2)]

    def _ids(self, pts, cam_center):
        return get_quadrant_id(pts.reshape(-1, 3).float() - cam_center.reshape(1, 3).to(pts.device).float(), self.quadrant_divider)

    def add_pts(self, pts, cam_center, **kw):
        pts = pts.reshape(-1, 3)
        ids = self._ids(pts, cam_center)
        for quad in torch.unique(ids).tolist():
            self.search_structs[quad].add_pts(pts[ids == quad])

    def query(self, pts, cam_center, **kw):
        pts = pts.reshape(-1, 3)
        ids = self._ids(pts, cam_center)
        dists = torch.zeros((pts.shape[0],), dtype=torch.float32, device=pts.device)
        for quad in torch.unique(ids).tolist():
            sel = ids == quad
            dists[sel] = self.search_structs[quad].query(pts[sel])
        return dists


QuandrantSearcher = QuadrantSearcher          # the reference's spelling (nns.py:65)


def get_searcher(method, isquadrant=False):
    """must3r/slam/nns.py:9-19"""
    if 'quadrant_x' in method and not isquadrant:
        return QuadrantSearcher(method)
    if "kdtree-scipy" in method:
        return DeviceNN()
    if method == 'none':
        return None
    raise ValueError(f"Unknown searcher method {method}")


def get_overlap_score(res, overlap_tree, cam_center, mode='nn', kf_x_subsamp=None, min_conf_keyframe=1.5, percentile=70,
                      eps=1e-9):
    """must3r/slam/model.py:62-91: how much of the new frame is NOT yet covered by the keyframes' points = the `percentile`
    of the distances from its confident points to their nearest stored neighbour (optionally / depth), or a confidence
    statistic."""
    if mode == 'meanconf':
        return res['conf'].mean()
    if mode == 'medianconf':
        return res['conf'].median()
    if 'nn' not in mode:
        raise ValueError(f"Unknown overlap score method {mode}")
    pts3d = res['pts3d'][0, 0, ::kf_x_subsamp, ::kf_x_subsamp] if kf_x_subsamp else res['pts3d']
    msk = res['conf'][0, 0, ::kf_x_subsamp, ::kf_x_subsamp] if kf_x_subsamp else res['conf']
    msk = msk > min_conf_keyframe
    if int(msk.sum()) == 0:
        return 0.
    dists = overlap_tree.query(pts3d[msk], cam_center=cam_center)
    if 'norm' in mode:
        depths = res['pts3d_local'][0, 0, ::kf_x_subsamp, ::kf_x_subsamp, -1]
        dists = dists / (depths[msk].float() + eps)
    dists = torch.where(torch.isposinf(dists), torch.full_like(dists, torch.finfo(dists.dtype).max), dists)
    return float(torch.quantile(dists.double(), percentile / 100.0, interpolation='linear'))      # np.percentile's default


def choose_keyframe_from_overlap(overlap_score, thr, overlap_mode):
    """must3r/slam/model.py:123-128"""
    return overlap_score > thr if 'nn' in overlap_mode else overlap_score < thr
