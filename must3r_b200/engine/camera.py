"""Camera recovery of ``postprocess(compute_cam=True)`` (must3r/engine/inference.py:29-47): focal length by the
Weiszfeld re-weighted fit of ``estimate_focal_knowing_depth(..., focal_mode='weiszfeld')``
(dust3r/dust3r/post_process.py:12-60) and camera-to-world pose by the weighted rigid registration that the reference gets
from ``roma.rigid_points_registration(x, y, weights, compute_scaling=False)`` (roma is not vendored in the reference;
its documented algorithm = weighted Kabsch / orthogonal Procrustes: R = argmin sum_i w_i |R x_i + t - y_i|^2).

Outside the timed hot path: a few dozen small torch ops on the tensors' own device; the rotation is
Horn's closed-form quaternion solution evaluated with batched 4x4 matmuls - no SVD / cuSOLVER call on the GPU.
"""
from __future__ import annotations

import math

import torch


def estimate_focal_weiszfeld(pts3d_local: torch.Tensor, pp: torch.Tensor, iters: int = 10) -> torch.Tensor:
    """pts3d_local [B,H,W,3] (camera frame), pp = principal point (x, y) -> focal [B] in pixels.
    focal = argmin sum |pixel - focal * (x, y) / z|: closed-form L2 start, then `iters` IRLS steps with weights 1 / distance."""
    B, H, W, _ = pts3d_local.shape
    dev = pts3d_local.device
    xs = torch.arange(W, device=dev, dtype=torch.float32)
    ys = torch.arange(H, device=dev, dtype=torch.float32)
    px = torch.stack([xs[None, :].expand(H, W), ys[:, None].expand(H, W)], dim=-1).reshape(1, H * W, 2) - pp.reshape(1, 1, 2).float()
    p = pts3d_local.reshape(B, H * W, 3).float()
    ray = torch.nan_to_num(p[..., :2] / p[..., 2:3], posinf=0.0, neginf=0.0)
    num = (ray * px).sum(-1)                       # <(x,y)/z, pixel>
    den = ray.square().sum(-1)
    focal = num.mean(1) / den.mean(1)
    for _ in range(iters):
        dist = (px - focal.view(B, 1, 1) * ray).norm(dim=-1)
        w = dist.clamp(min=1e-8).reciprocal()
        focal = (w * num).mean(1) / (w * den).mean(1)
    # the reference clips to [min_focal, max_focal] * focal_base with defaults (0, inf): only the lower bound acts
    return focal.clamp(min=0.0)


def _rotation_horn(S: torch.Tensor) -> torch.Tensor:
    """Proper rotation maximising tr(R S) for S = sum_i w_i x_i y_i^T (Horn 1987: the unit quaternion is the eigenvector of the
    largest eigenvalue of a symmetric 4x4 matrix built from S).  No LAPACK / cuSOLVER: the dominant eigenvector comes from ten
    squarings of the (shifted, normalised) 4x4 matrix = power iteration with exponent 1024, a handful of batched 4x4 matmuls on
    whatever device the points live on.  Always det R = +1 (the reflection case of Kabsch needs no special handling)."""
    Sxx, Sxy, Sxz = S[..., 0, 0], S[..., 0, 1], S[..., 0, 2]
    Syx, Syy, Syz = S[..., 1, 0], S[..., 1, 1], S[..., 1, 2]
    Szx, Szy, Szz = S[..., 2, 0], S[..., 2, 1], S[..., 2, 2]
    N = torch.stack([
        torch.stack([Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx], -1),
        torch.stack([Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz], -1),
        torch.stack([Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy], -1),
        torch.stack([Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz], -1)], -2).double()
    eye = torch.eye(4, dtype=N.dtype, device=N.device)
    nrm = N.flatten(-2).norm(dim=-1).clamp(min=1e-300)[..., None, None]
    P = N / nrm + eye                                   # eigenvalues in [0, 2]: the largest algebraic one dominates
    for _ in range(10):
        P = P @ P
        P = P / P.flatten(-2).norm(dim=-1).clamp(min=1e-300)[..., None, None]
    # P ~ v v^T: its column of largest norm is the eigenvector (up to sign)
    col = P.square().sum(dim=-2).argmax(dim=-1)
    q = torch.gather(P, -1, col[..., None, None].expand(*P.shape[:-1], 1)).squeeze(-1)
    q = q / q.norm(dim=-1, keepdim=True).clamp(min=1e-300)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
        torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], -2)
    return R.float()


def rigid_points_registration(x: torch.Tensor, y: torch.Tensor, weights: torch.Tensor, method: str = "horn"):
    """Weighted rigid registration: (R [...,3,3], t [...,3]) minimising sum_i w_i |R x_i + t - y_i|^2, det R = +1.  x, y [...,n,3].
    method 'horn' (default): closed-form quaternion solution, no SVD call (no cuSOLVER on the GPU); 'svd': weighted Kabsch through
    torch.linalg.svd - the same minimiser, kept as the cross-check of the tests."""
    w = weights.unsqueeze(-1).float()
    x, y = x.float(), y.float()
    wsum = w.sum(dim=-2, keepdim=True)
    xc = (w * x).sum(dim=-2, keepdim=True) / wsum
    yc = (w * y).sum(dim=-2, keepdim=True) / wsum
    cov = ((w * (y - yc)).transpose(-1, -2) @ (x - xc))              # sum w (y - yc)(x - xc)^T
    if method == "svd":
        U, _, Vh = torch.linalg.svd(cov)
        d = torch.sign(torch.linalg.det(U @ Vh))
        D = torch.diag_embed(torch.stack([torch.ones_like(d), torch.ones_like(d), d], dim=-1))
        R = U @ D @ Vh
    else:
        R = _rotation_horn(cov.transpose(-1, -2))
    t = yc.squeeze(-2) - (R @ xc.transpose(-1, -2)).squeeze(-1)
    return R, t


def camera_from_pointmaps(out: dict) -> dict:
    """Adds 'focal' [batch dims] and 'c2w' [batch dims, 4, 4] to a postprocess dict holding pts3d, pts3d_local, conf."""
    pts, loc, conf = out["pts3d"], out["pts3d_local"], out["conf"]
    batch = pts.shape[:-3]
    H, W = conf.shape[-2:]
    nb = math.prod(batch)
    pp = torch.tensor((W / 2, H / 2), device=pts.device)
    out["focal"] = estimate_focal_weiszfeld(loc.reshape(nb, H, W, 3), pp).reshape(*batch)
    R, t = rigid_points_registration(loc.reshape(*batch, -1, 3), pts.reshape(*batch, -1, 3), conf.reshape(*batch, -1) - 1.0)
    c2w = torch.eye(4, device=pts.device).view(*([1] * len(batch)), 4, 4).repeat(*batch, 1, 1)
    c2w[..., :3, :3] = R
    c2w[..., :3, 3] = t
    out["c2w"] = c2w
    return out
