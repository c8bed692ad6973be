"""CPU oracle for the MUSt3R multi-view inference hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``must3r_b200``) never does, and fails loudly
when its CUDA extension is missing.

It is a plain-PyTorch fp32 *restatement* (not a copy) of the reference algorithm:
explicit matmul/softmax/erf arithmetic over a flat ``state_dict`` whose keys are the
reference's.  Every function cites the reference file:line it follows
(paths relative to /root/reference).

Parity pinning: the reference ships no tests or golden vectors, so this
oracle is pinned against outputs of the reference itself, generated in the build
container by ``tests/golden/make_golden.py`` (which imports /root/reference) and committed
under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------
@dataclass
class EncoderConfig:
    """must3r/model/encoder.py:14-23 ctor defaults (ViT-L)."""
    img_size: Tuple[int, int] = (224, 224)
    patch_size: int = 16
    embed_dim: int = 1024
    depth: int = 24
    num_heads: int = 16
    mlp_ratio: int = 4
    rope_base: float = 100.0
    rope_f0: float = 1.0
    ln_eps: float = 1e-6
    patch_embed: str = "PatchEmbedDust3R"     # or 'ManyAR_PatchEmbed' (dust3r/dust3r/patch_embed.py:13-16)


@dataclass
class DecoderConfig:
    """must3r/model/decoder.py:19-36 ctor defaults (ViT-B memory decoder)."""
    img_size: Tuple[int, int] = (224, 224)
    enc_embed_dim: int = 1024
    patch_size: int = 16
    embed_dim: int = 768
    output_dim: int = 1792
    depth: int = 12
    num_heads: int = 12
    mlp_ratio: int = 4
    rope_base: float = 100.0
    rope_f0: float = 1.0
    ln_eps: float = 1e-6
    feedback_type: Optional[str] = "single_mlp"
    feedback_ln_eps: float = 1e-5      # must3r/model/feedback_mechanism.py:14 (nn.LayerNorm default)
    memory_mode: str = "kv"


def parse_pos_embed(name: str) -> Tuple[float, float]:
    """must3r/model/blocks/pos_embed.py:7-22: 'RoPE100' -> (100, 1); 'RoPE100_224:512' -> (100, 224/512)."""
    assert name.startswith("RoPE")
    f0 = 1.0
    if "_" in name:
        name, res = name.split("_")
        old, new = res.split(":")
        f0 = float(old) / float(new)
    return float(name[len("RoPE"):]), f0


# --------------------------------------------------------------------------------------
# primitive ops
# --------------------------------------------------------------------------------------
def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    y = x @ w.t()
    return y if b is None else y + b


def layernorm(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    """nn.LayerNorm over the last dim (biased variance)."""
    mu = x.mean(dim=-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=-1, keepdim=True)
    return xc * torch.rsqrt(var + eps) * w + b


def gelu_erf(x: Tensor) -> Tensor:
    """nn.GELU() default = exact erf form (dust3r/croco/models/blocks.py:68)."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def rope2d(tokens: Tensor, pos: Tensor, base: float, f0: float) -> Tensor:
    """2-D RoPE with curope semantics (dust3r/croco/models/curope/kernels.cu:40-80,
    curope.cpp:11-47): per head-dim D, Q=D/4; dims [0,Q)|[Q,2Q) are the (u,v) pair rotated by
    pos_y, dims [2Q,3Q)|[3Q,4Q) the pair rotated by pos_x; angle = pos * f0 / base**(d/Q) in fp32.
    tokens: [B,H,N,D]; pos: [B,N,2] int64 (y,x).  Returns a new tensor."""
    B, H, N, D = tokens.shape
    Q = D // 4
    d = torch.arange(Q, dtype=torch.float32, device=tokens.device)
    inv_freq = f0 / torch.pow(torch.tensor(base, dtype=torch.float32, device=tokens.device), d / Q)
    t = tokens.float()
    out = torch.empty_like(t)
    for axis in range(2):
        ang = pos[:, :, axis].to(torch.float32)[:, None, :, None] * inv_freq  # [B,1,N,Q]
        c, s = torch.cos(ang), torch.sin(ang)
        u = t[..., axis * 2 * Q: axis * 2 * Q + Q]
        v = t[..., axis * 2 * Q + Q: axis * 2 * Q + 2 * Q]
        out[..., axis * 2 * Q: axis * 2 * Q + Q] = u * c - v * s
        out[..., axis * 2 * Q + Q: axis * 2 * Q + 2 * Q] = v * c + u * s
    return out.to(tokens.dtype)


def sdpa(q: Tensor, k: Tensor, v: Tensor, key_mask: Optional[Tensor] = None) -> Tensor:
    """softmax(q k^T / sqrt(hd)) v on [B,H,N,hd] (must3r/model/blocks/attention.py:65-78).
    key_mask: optional bool [B,Nk], True = key is attended."""
    scale = q.shape[-1] ** -0.5
    s = (q @ k.transpose(-2, -1)) * scale
    if key_mask is not None:
        s = s.masked_fill(~key_mask[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    return p @ v


def self_attention(sd: Dict[str, Tensor], pfx: str, x: Tensor, pos: Tensor, num_heads: int,
                   base: float, f0: float) -> Tensor:
    """Attention.forward, must3r/model/blocks/attention.py:92-99 (qkv -> [B,N,3,H,hd], RoPE on q,k)."""
    B, N, C = x.shape
    hd = C // num_heads
    qkv = linear(x, sd[pfx + "qkv.weight"], sd.get(pfx + "qkv.bias"))
    qkv = qkv.reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)  # 3,B,H,N,hd
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = rope2d(q, pos, base, f0)
    k = rope2d(k, pos, base, f0)
    o = sdpa(q, k, v).transpose(1, 2).reshape(B, N, C)
    return linear(o, sd[pfx + "proj.weight"], sd[pfx + "proj.bias"])


def mlp(sd: Dict[str, Tensor], pfx: str, x: Tensor) -> Tensor:
    """Mlp.forward, dust3r/croco/models/blocks.py:74-80 (dropouts are identity at eval)."""
    h = gelu_erf(linear(x, sd[pfx + "fc1.weight"], sd[pfx + "fc1.bias"]))
    return linear(h, sd[pfx + "fc2.weight"], sd[pfx + "fc2.bias"])


# --------------------------------------------------------------------------------------
# encoder
# --------------------------------------------------------------------------------------
def patch_embed(sd: Dict[str, Tensor], img: Tensor, patch: int) -> Tuple[Tensor, Tensor]:
    """PatchEmbedDust3R.forward (dust3r/dust3r/patch_embed.py:20-29): 16x16/stride-16 conv as an
    im2col GEMM; tokens in raster order; pos = cartesian_prod(arange(h), arange(w)) = (y, x)
    (dust3r/croco/models/blocks.py:195-207)."""
    V, C, H, W = img.shape
    assert H % patch == 0 and W % patch == 0
    h, w = H // patch, W // patch
    cols = img.reshape(V, C, h, patch, w, patch).permute(0, 2, 4, 1, 3, 5).reshape(V, h * w, C * patch * patch)
    wgt = sd["patch_embed.proj.weight"].reshape(sd["patch_embed.proj.weight"].shape[0], -1)
    x = cols @ wgt.t() + sd["patch_embed.proj.bias"]
    ys = torch.arange(h, device=img.device)
    xs = torch.arange(w, device=img.device)
    pos = torch.stack([ys[:, None].expand(h, w), xs[None, :].expand(h, w)], dim=-1).reshape(1, h * w, 2)
    return x, pos.expand(V, -1, -1).clone()


def patch_embed_many_ar(sd: Dict[str, Tensor], img: Tensor, true_shape: Tensor, patch: int) -> Tuple[Tensor, Tensor]:
    """ManyAR_PatchEmbed.forward (dust3r/dust3r/patch_embed.py:42-70): the batch is stored landscape (W >= H); views whose
    true_shape is portrait are transposed before the projection and get the positions of the transposed grid."""
    V, C, H, W = img.shape
    assert W >= H, f"img should be in landscape mode, but got {W=} {H=}"
    assert tuple(true_shape.shape) == (V, 2)
    n_tok = (H // patch) * (W // patch)
    x = img.new_zeros((V, n_tok, sd["patch_embed.proj.bias"].shape[0]))
    pos = torch.zeros((V, n_tok, 2), dtype=torch.int64, device=img.device)
    portrait = true_shape[:, 1] < true_shape[:, 0]
    if bool((~portrait).any()):
        x[~portrait], pos[~portrait] = patch_embed(sd, img[~portrait], patch)
    if bool(portrait.any()):
        x[portrait], pos[portrait] = patch_embed(sd, img[portrait].swapaxes(-1, -2), patch)
    return x, pos


def encoder_forward(sd: Dict[str, Tensor], cfg: EncoderConfig, img: Tensor, true_shape: Tensor
                    ) -> Tuple[Tensor, Tensor]:
    """Dust3rEncoder.forward, must3r/model/encoder.py:46-52 (always fp32)."""
    if cfg.patch_embed == "ManyAR_PatchEmbed":
        x, pos = patch_embed_many_ar(sd, img.float(), true_shape, cfg.patch_size)
    else:
        x, pos = patch_embed(sd, img.float(), cfg.patch_size)
    for i in range(cfg.depth):
        p = f"blocks_enc.{i}."
        # Block.forward, must3r/model/blocks/layers.py:51-54
        h = layernorm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"], cfg.ln_eps)
        x = x + self_attention(sd, p + "attn.", h, pos, cfg.num_heads, cfg.rope_base, cfg.rope_f0)
        h = layernorm(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"], cfg.ln_eps)
        x = x + mlp(sd, p + "mlp.", h)
    x = layernorm(x, sd["norm_enc.weight"], sd["norm_enc.bias"], cfg.ln_eps)
    return x, pos


# --------------------------------------------------------------------------------------
# decoder
# --------------------------------------------------------------------------------------
def _prepare_y(sd: Dict[str, Tensor], cfg: DecoderConfig, l: int, y: Tensor) -> Tensor:
    """CachedDecoderBlock.prepare_y, must3r/model/blocks/layers.py:81-88."""
    p = f"blocks_dec.{l}."
    if cfg.memory_mode == "raw":
        return y
    y_ = layernorm(y, sd[p + "norm_y.weight"], sd[p + "norm_y.bias"], cfg.ln_eps)
    if cfg.memory_mode == "norm_y":
        return y_
    k = linear(y_, sd[p + "cross_attn.projk.weight"], sd[p + "cross_attn.projk.bias"])
    v = linear(y_, sd[p + "cross_attn.projv.weight"], sd[p + "cross_attn.projv.bias"])
    return torch.cat([k, v], dim=-1)


def _mem_to_kv(sd: Dict[str, Tensor], cfg: DecoderConfig, l: int, mem: Tensor) -> Tuple[Tensor, Tensor]:
    """CachedDecoderBlock.forward lines layers.py:91-96: how stored memory becomes (key, value)."""
    p = f"blocks_dec.{l}."
    D = cfg.embed_dim
    if cfg.memory_mode == "kv":
        return mem[..., :D], mem[..., D:]
    y_ = mem
    if cfg.memory_mode == "raw":
        y_ = layernorm(mem, sd[p + "norm_y.weight"], sd[p + "norm_y.bias"], cfg.ln_eps)
    k = linear(y_, sd[p + "cross_attn.projk.weight"], sd[p + "cross_attn.projk.bias"])
    v = linear(y_, sd[p + "cross_attn.projv.weight"], sd[p + "cross_attn.projv.bias"])
    return k, v


def _decoder_block(sd: Dict[str, Tensor], cfg: DecoderConfig, l: int, x: Tensor, pos: Tensor,
                   key: Tensor, value: Tensor, key_mask: Optional[Tensor]) -> Tensor:
    """CachedDecoderBlock.forward, must3r/model/blocks/layers.py:90-99; cross-attention has
    pos_embed=None (layers.py:72) so no RoPE on memory keys; CachedCrossAttention.forward
    attention.py:139-149.  x: [BV,N,D]; key/value: [BV,Nk,D]; key_mask: [BV,Nk] True=attend."""
    p = f"blocks_dec.{l}."
    H = cfg.num_heads
    h = layernorm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"], cfg.ln_eps)
    x = x + self_attention(sd, p + "attn.", h, pos, H, cfg.rope_base, cfg.rope_f0)
    BV, N, D = x.shape
    hd = D // H
    h = layernorm(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"], cfg.ln_eps)
    q = linear(h, sd[p + "cross_attn.projq.weight"], sd[p + "cross_attn.projq.bias"])
    q = q.reshape(BV, N, H, hd).permute(0, 2, 1, 3)
    k = key.reshape(BV, -1, H, hd).permute(0, 2, 1, 3)
    v = value.reshape(BV, -1, H, hd).permute(0, 2, 1, 3)
    o = sdpa(q, k, v, key_mask).transpose(1, 2).reshape(BV, N, D)
    x = x + linear(o, sd[p + "cross_attn.proj.weight"], sd[p + "cross_attn.proj.bias"])
    h = layernorm(x, sd[p + "norm3.weight"], sd[p + "norm3.bias"], cfg.ln_eps)
    return x + mlp(sd, p + "mlp.", h)


def _feedback(sd: Dict[str, Tensor], cfg: DecoderConfig, new_mem: List[Tensor]) -> List[Tensor]:
    """run_feedback_layers, must3r/model/feedback_mechanism.py:39-53."""
    if not cfg.feedback_type:
        return new_mem
    h = layernorm(new_mem[-1], sd["feedback_norm.weight"], sd["feedback_norm.bias"], cfg.feedback_ln_eps)
    if cfg.feedback_type == "single_mlp":
        off = mlp(sd, "feedback_layer.", h)
    elif cfg.feedback_type == "single_linear":
        off = linear(h, sd["feedback_layer.weight"], sd["feedback_layer.bias"])
    else:
        raise ValueError(cfg.feedback_type)
    return [m + off for m in new_mem[:-1]] + [new_mem[-1]]


def _head(sd: Dict[str, Tensor], cfg: DecoderConfig, x: Tensor, H: int, W: int) -> Tensor:
    """_compute_prediction_head decoder.py:149-156 + LinearHead.forward head.py:69-72 +
    unpatchify tools/image.py:9-14: out[b,16y+i,16x+j,c] = proj[b, y*w+x, c*256+16i+j]."""
    P = cfg.patch_size
    h = layernorm(x, sd["norm_dec.weight"], sd["norm_dec.bias"], cfg.ln_eps).float()
    y = linear(h, sd["head_dec.proj.weight"], sd["head_dec.proj.bias"])  # [BV,N,C*P*P]
    BV, N, _ = y.shape
    gh, gw = H // P, W // P
    C = cfg.output_dim // (P * P)
    y = y.reshape(BV, gh, gw, C, P, P).permute(0, 1, 4, 2, 5, 3)  # BV,gh,P,gw,P,C
    return y.reshape(BV, H, W, C)


Memory = Tuple[List[Tensor], Tensor, int, int, int]


def decoder_forward_list(sd: Dict[str, Tensor], cfg: DecoderConfig, x: Sequence[Tensor],
                         pos: Sequence[Tensor], true_shape: Sequence[Tensor],
                         current_mem: Optional[Memory] = None, render: bool = False
                         ) -> Tuple[Memory, List[Tensor]]:
    """MUSt3R.forward_list, must3r/model/decoder.py:158-265 (one entry per aspect-ratio group,
    x[i]: [B,n_i,N_i,Denc]).  The tensor form (decoder.py:267-350) is the 1-group special case."""
    G = len(x)
    D = cfg.embed_dim
    mem_D = 2 * D if cfg.memory_mode == "kv" else D
    B = x[0].shape[0]
    nimgs = [xi.shape[1] for xi in x]
    Ns = [xi.shape[2] for xi in x]
    xs, ps = [], []
    for i in range(G):
        xi = linear(x[i].float().reshape(B * nimgs[i], Ns[i], -1), sd["feat_embed_enc_to_dec.weight"],
                    sd["feat_embed_enc_to_dec.bias"]).reshape(B, nimgs[i], Ns[i], D)
        emb = sd["image2_embed"].reshape(1, 1, 1, D)
        if current_mem is None and i == 0:
            xi = torch.cat([xi[:, :1], xi[:, 1:] + emb], dim=1)   # decoder.py:176-178,280-282
        else:
            xi = xi + emb                                          # decoder.py:179-180,287
        xs.append(xi.reshape(B * nimgs[i], Ns[i], D))
        ps.append(pos[i].reshape(B * nimgs[i], Ns[i], 2))
    if current_mem is None:
        mem_vals = [torch.zeros((B, 0, mem_D), dtype=torch.float32, device=xs[0].device) for _ in range(cfg.depth)]
        mem_labels = torch.zeros((B, 0), dtype=torch.int64, device=xs[0].device)
        mem_nimgs, mem_pi, mem_pt = 0, 0, 0                        # decoder.py:141-147
    else:
        mem_vals, mem_labels, mem_nimgs, mem_pi, mem_pt = current_mem
        mem_vals = [m.float() for m in mem_vals]
    Nm = mem_vals[0].shape[1]
    tok = [n * N for n, N in zip(nimgs, Ns)]
    Nt = sum(tok)
    use_mask = (not render) and (Nm > 0 or sum(nimgs) > 1)         # decoder.py:199-204,291-296
    # key-validity masks per group: [n_i, Nm+Nt]; an image never attends to its own new tokens
    masks: List[Optional[Tensor]] = [None] * G
    if use_mask:                                                   # make_mem_mask decoder.py:119-139
        off = 0
        for i in range(G):
            m = torch.ones((nimgs[i], Nm + Nt), dtype=torch.bool, device=xs[0].device)
            for j in range(nimgs[i]):
                m[j, Nm + off + j * Ns[i]: Nm + off + (j + 1) * Ns[i]] = False
            masks[i] = m
            off += tok[i]

    new_mem: List[Tensor] = []
    for l in range(cfg.depth):
        if not render:                                             # decoder.py:208-213,301-306
            x_cat = torch.cat([xi.reshape(B, -1, D) for xi in xs], dim=1)
            new_mem.append(x_cat)
            mem_l = torch.cat([mem_vals[l], _prepare_y(sd, cfg, l, x_cat)], dim=1)
        else:
            mem_l = mem_vals[l]
        key, value = _mem_to_kv(sd, cfg, l, mem_l)                 # [B,Nk,D]
        Nk = key.shape[1]
        for i in range(G):
            k_i = key[:, None].expand(B, nimgs[i], Nk, D).reshape(B * nimgs[i], Nk, D)
            v_i = value[:, None].expand(B, nimgs[i], Nk, D).reshape(B * nimgs[i], Nk, D)
            km = None if masks[i] is None else masks[i][None].expand(B, -1, -1).reshape(B * nimgs[i], Nk)
            xs[i] = _decoder_block(sd, cfg, l, xs[i], ps[i], k_i, v_i, km)

    if not render:                                                 # decoder.py:230-249,323-338
        new_mem = _feedback(sd, cfg, new_mem)
        mem_out = [torch.cat([mem_vals[l], _prepare_y(sd, cfg, l, new_mem[l])], dim=1) for l in range(cfg.depth)]
        labels, off = [], 0
        for i in range(G):
            li = torch.arange(nimgs[i], dtype=torch.int64, device=mem_labels.device)
            li = li.view(1, nimgs[i], 1).repeat(B, 1, Ns[i]).view(B, tok[i]) + mem_nimgs + off
            labels.append(li)
            off += nimgs[i]
        mem_labels_out = torch.cat([mem_labels] + labels, dim=1)
        n_tot = mem_nimgs + sum(nimgs)
        out: Memory = (mem_out, mem_labels_out, n_tot, n_tot, mem_labels_out.shape[1])
    else:
        out = (mem_vals if current_mem is None else list(current_mem[0]), mem_labels, mem_nimgs, mem_pi, mem_pt)

    preds = []
    for i in range(G):
        ts = true_shape[i].reshape(B * nimgs[i], 2)
        assert bool((ts == ts[:1]).all()), "true_shape must be all identical"  # head.py:31
        H, W = int(ts[0, 0]), int(ts[0, 1])
        pm = _head(sd, cfg, xs[i], H, W)
        preds.append(pm.reshape(B, nimgs[i], H, W, -1))
    return out, preds


def decoder_forward(sd, cfg, x, pos, true_shape, current_mem=None, render=False):
    """MUSt3R.forward, must3r/model/decoder.py:267-350 (tensor form or list form)."""
    if isinstance(x, (list, tuple)):
        return decoder_forward_list(sd, cfg, list(x), list(pos), list(true_shape), current_mem, render)
    assert not render or current_mem is not None                   # decoder.py:278
    out, preds = decoder_forward_list(sd, cfg, [x], [pos], [true_shape], current_mem, render)
    return out, preds[0]


# --------------------------------------------------------------------------------------
# postprocess
# --------------------------------------------------------------------------------------
def exp_to_norm(xyz: Tensor) -> Tensor:
    """apply_exp_to_norm, must3r/tools/geometry.py:14-18."""
    d = xyz.norm(dim=-1, keepdim=True)
    return xyz / d.clip(min=1e-8) * torch.expm1(d)


def postprocess(pointmaps: Tensor, activation: str = "norm_exp") -> Dict[str, Tensor]:
    """postprocess(compute_cam=False), must3r/engine/inference.py:16-27."""
    pm = pointmaps.float()
    act = exp_to_norm if activation == "norm_exp" else (lambda t: t)
    out = {"pts3d": act(pm[..., :3])}
    if pm.shape[-1] >= 6:
        out["pts3d_local"] = act(pm[..., 3:6])
    if pm.shape[-1] in (4, 7):
        out["conf"] = 1.0 + pm[..., -1].exp()
    return out


def focal_weiszfeld(pts3d_local: Tensor, pp_xy: Tuple[float, float], iters: int = 10):
    """estimate_focal_knowing_depth(focal_mode='weiszfeld'), dust3r/dust3r/post_process.py:12-60, in float64 numpy.
    pts3d_local [B,H,W,3]; pixel grid (x = column, y = row, dust3r/dust3r/utils/geometry.py:15-37) minus the principal
    point; focal = argmin sum |pixel - f (x,y)/z| by 10 IRLS steps from the L2 closed form."""
    import numpy as np
    p = pts3d_local.double().numpy()
    B, H, W, _ = p.shape
    gx, gy = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64), indexing="xy")
    px = np.stack([gx - pp_xy[0], gy - pp_xy[1]], -1).reshape(1, H * W, 2)
    p = p.reshape(B, H * W, 3)
    with np.errstate(divide="ignore", invalid="ignore"):
        ray = p[..., :2] / p[..., 2:3]
    ray = np.nan_to_num(ray, nan=0.0, posinf=0.0, neginf=0.0)
    num, den = (ray * px).sum(-1), (ray ** 2).sum(-1)
    f = num.mean(1) / den.mean(1)
    for _ in range(iters):
        w = 1.0 / np.clip(np.linalg.norm(px - f[:, None, None] * ray, axis=-1), 1e-8, None)
        f = (w * num).mean(1) / (w * den).mean(1)
    return torch.from_numpy(np.clip(f, 0.0, None))


def rigid_registration(x: Tensor, y: Tensor, w: Tensor):
    """roma.rigid_points_registration(x, y, weights=w, compute_scaling=False) as used at must3r/engine/inference.py:38-41
    (roma is a PyPI dependency of the reference, not vendored, version unpinned; its documented algorithm is the weighted
    orthogonal Procrustes / Kabsch solution).  One problem: x, y [n,3], w [n] -> R [3,3] (det +1), t [3]; float64 numpy."""
    import numpy as np
    x, y, w = x.double().numpy(), y.double().numpy(), w.double().numpy()[:, None]
    xc, yc = (w * x).sum(0) / w.sum(), (w * y).sum(0) / w.sum()
    H = (w * (y - yc)).T @ (x - xc)
    U, _, Vt = np.linalg.svd(H)
    D = np.diag([1.0, 1.0, np.sign(np.linalg.det(U @ Vt))])
    R = U @ D @ Vt
    return torch.from_numpy(R), torch.from_numpy(yc - R @ xc)


# --------------------------------------------------------------------------------------
# callable wrappers with the reference's model API (so engine code can drive the oracle)
# --------------------------------------------------------------------------------------
class OracleEncoder:
    def __init__(self, sd: Dict[str, Tensor], cfg: EncoderConfig):
        self.sd = {k: v.float() for k, v in sd.items()}
        self.cfg = cfg
        self.patch_size, self.embed_dim, self.depth = cfg.patch_size, cfg.embed_dim, cfg.depth

    @torch.no_grad()
    def __call__(self, img, true_shape):
        return encoder_forward(self.sd, self.cfg, img, true_shape)


class OracleDecoder:
    pointmaps_activation = "norm_exp"

    def __init__(self, sd: Dict[str, Tensor], cfg: DecoderConfig):
        self.sd = {k: v.float() for k, v in sd.items()}
        self.cfg = cfg
        self.memory_mode, self.embed_dim, self.depth = cfg.memory_mode, cfg.embed_dim, cfg.depth

    @torch.no_grad()
    def __call__(self, x, pos, true_shape, current_mem=None, render=False):
        return decoder_forward(self.sd, self.cfg, x, pos, true_shape, current_mem, render)
