"""Pin the CPU oracle (oracle/must3r_oracle.py) against outputs of the unmodified reference
(tests/golden/*.npz, made by tests/golden/make_golden.py).  fp32 vs fp32 on CPU: tolerance 2e-5 rel-L2
(different op order: explicit softmax/erf vs SDPA/nn.GELU, RoPE fallback vs curope pairing)."""
import numpy as np
import pytest
import torch

from helpers import load_golden, tiny_oracle, full_oracle, digest, rel
from must3r_b200 import synthetic as syn
from oracle import must3r_oracle as orc

TOL = 2e-5

VARIANTS = {
    "kv": ({}, {}),
    "normy": ({}, dict(memory_mode="norm_y")),
    "raw": ({}, dict(memory_mode="raw")),
    "f0": (dict(rope_f0=224 / 512), dict(rope_f0=224 / 512)),
    "nofb": ({}, dict(feedback_type=None)),
    "fblin": ({}, dict(feedback_type="single_linear")),
}


def check_mem(g, prefix, mem, depth=3):
    for l in range(depth):
        assert rel(mem[0][l], g[f"{prefix}.mem{l}"]) < TOL, (prefix, l)
    assert np.array_equal(mem[1].numpy(), g[f"{prefix}.labels"])
    assert list(mem[2:]) == g[f"{prefix}.tail"].tolist()


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_tiny_chain(variant):
    g = load_golden("tiny_model.npz")
    enc, dec = tiny_oracle(7, *VARIANTS[variant])
    imgs, ts = syn.synthetic_views(5, 32, 48, seed=11)
    x, pos = enc(imgs, ts)
    assert rel(x, g[f"{variant}.enc_x"]) < TOL
    assert np.array_equal(pos.numpy(), g[f"{variant}.enc_pos"])
    mem, pm = dec(x[None, 0:2], pos[None, 0:2], ts[None, 0:2], None)
    assert rel(pm, g[f"{variant}.pm_init"]) < TOL
    check_mem(g, f"{variant}.init", mem)
    mem, pm = dec(x[None, 2:3], pos[None, 2:3], ts[None, 2:3], mem)
    assert rel(pm, g[f"{variant}.pm_upd1"]) < TOL
    mem, pm = dec(x[None, 3:5], pos[None, 3:5], ts[None, 3:5], mem)
    assert rel(pm, g[f"{variant}.pm_upd2"]) < TOL
    check_mem(g, f"{variant}.final", mem)
    mem_r, pm = dec(x[None], pos[None], ts[None], mem, render=True)
    assert rel(pm, g[f"{variant}.pm_render"]) < TOL
    assert mem_r[0][0] is mem[0][0]  # render returns the memory untouched (decoder.py:340)


def test_tiny_batch2_single_image_init():
    g = load_golden("tiny_model.npz")
    enc, dec = tiny_oracle(7)
    imgs, ts = syn.synthetic_views(6, 32, 48, seed=12)
    x, pos = enc(imgs, ts)
    xb, pb, tb = x.view(2, 3, *x.shape[1:]), pos.view(2, 3, *pos.shape[1:]), ts.view(2, 3, 2)
    mem, pm = dec(xb[:, :1], pb[:, :1], tb[:, :1], None)
    assert rel(pm, g["b2.pm_init1"]) < TOL
    mem, pm = dec(xb[:, 1:3], pb[:, 1:3], tb[:, 1:3], mem)
    assert rel(pm, g["b2.pm_upd2"]) < TOL
    check_mem(g, "b2.final", mem)
    _, pm = dec(xb, pb, tb, mem, render=True)
    assert rel(pm, g["b2.pm_render"]) < TOL


def test_tiny_list_form_two_aspect_ratios():
    g = load_golden("tiny_model.npz")
    enc, dec = tiny_oracle(7)
    imgs_p, ts_p = syn.synthetic_views(2, 48, 32, seed=13)
    xp, pp = enc(imgs_p, ts_p)
    imgs_l, ts_l = syn.synthetic_views(2, 32, 48, seed=14)
    xl, pl = enc(imgs_l, ts_l)
    mem, pms = dec([xl[None], xp[None]], [pl[None], pp[None]], [ts_l[None], ts_p[None]], None)
    assert rel(pms[0], g["list.pm0"]) < TOL and rel(pms[1], g["list.pm1"]) < TOL
    check_mem(g, "list.init", mem)
    mem2, pms = dec([xp[None, :1], xl[None, :1]], [pp[None, :1], pl[None, :1]], [ts_p[None, :1], ts_l[None, :1]], mem)
    assert rel(pms[0], g["list.pm2"]) < TOL and rel(pms[1], g["list.pm3"]) < TOL
    check_mem(g, "list.final", mem2)


def test_many_ar_patch_embed():
    g = load_golden("tiny_model.npz")
    enc, _ = tiny_oracle(7, dict(patch_embed="ManyAR_PatchEmbed"))
    imgs, _ = syn.synthetic_views(3, 32, 48, seed=15)
    ts = torch.tensor([[32, 48], [48, 32], [32, 48]], dtype=torch.int64)
    x, pos = enc(imgs, ts)
    assert rel(x, g["manyar.enc_x"]) < TOL
    assert np.array_equal(pos.numpy(), g["manyar.enc_pos"])


def test_postprocess():
    g = load_golden("tiny_model.npz")
    out = orc.postprocess(torch.from_numpy(g["kv.pm_render"]))
    for k in ("pts3d", "pts3d_local", "conf"):
        assert rel(out[k], g[f"kv.post.{k}"]) < 1e-6


def test_focal_weiszfeld_vs_reference():
    """Oracle restatement of estimate_focal_knowing_depth(focal_mode='weiszfeld') against the reference's own outputs:
    noisy pinhole pointmaps with focals 40 / 55 / 70 px (incl. a zero depth) and the degenerate random-weight pointmaps."""
    g = load_golden("tiny_model.npz")
    f = orc.focal_weiszfeld(torch.from_numpy(g["cam.pts_local"]), (48 / 2, 32 / 2))
    assert np.allclose(f.numpy(), g["cam.focal"], rtol=2e-5)
    loc = torch.from_numpy(g["kv.post.pts3d_local"]).reshape(-1, 32, 48, 3)
    f = orc.focal_weiszfeld(loc, (48 / 2, 32 / 2))
    assert np.allclose(f.numpy(), g["kv.post.focal"], rtol=1e-3, atol=1e-5)


def test_rigid_registration_known_answer():
    """Weighted Kabsch restatement (roma.rigid_points_registration): recovers a known rotation + translation exactly from
    noiseless points, ignores zero-weight outliers, and never returns a reflection."""
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(200, 3, generator=gen, dtype=torch.float64)
    ang = 0.7
    R0 = torch.tensor([[np.cos(ang), -np.sin(ang), 0.0], [np.sin(ang), np.cos(ang), 0.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
    R0 = R0 @ torch.tensor([[1.0, 0.0, 0.0], [0.0, np.cos(0.3), -np.sin(0.3)], [0.0, np.sin(0.3), np.cos(0.3)]], dtype=torch.float64)
    t0 = torch.tensor([0.5, -1.0, 2.0], dtype=torch.float64)
    y = x @ R0.T + t0
    w = torch.rand(200, generator=gen, dtype=torch.float64) + 0.1
    y[:10] += 5.0
    w[:10] = 0.0                                                    # outliers with zero weight
    R, t = orc.rigid_registration(x, y, w)
    assert torch.allclose(R, R0, atol=1e-10) and torch.allclose(t, t0, atol=1e-10)
    assert abs(float(torch.linalg.det(R)) - 1.0) < 1e-10
    Rm, _ = orc.rigid_registration(x, x * torch.tensor([1.0, 1.0, -1.0], dtype=torch.float64), torch.ones(200, dtype=torch.float64))
    assert float(torch.linalg.det(Rm)) > 0.999                       # mirrored target: still a proper rotation


@pytest.mark.parametrize("tag,H,W,size", [("224", 224, 224, 224)])
def test_full_size_digest(tag, H, W, size):
    """Full ViT-L / ViT-B architecture (24+12 layers) against reference digests."""
    g = load_golden("full_model_digest.npz")
    torch.set_num_threads(max(1, torch.get_num_threads()))
    enc, dec = full_oracle(size)
    imgs, ts = syn.synthetic_views(3, H, W, seed=2)
    x, pos = enc(imgs, ts)
    assert rel(digest(x), g[f"{tag}.enc_x"]) < TOL
    mem, pm = dec(x[None, :2], pos[None, :2], ts[None, :2], None)
    assert rel(digest(pm), g[f"{tag}.pm_init"]) < 5e-5
    mem, pm = dec(x[None, 2:3], pos[None, 2:3], ts[None, 2:3], mem)
    assert rel(digest(pm), g[f"{tag}.pm_upd"]) < 5e-5
    assert rel(digest(mem[0][0]), g[f"{tag}.mem0"]) < 5e-5
    assert rel(digest(mem[0][11]), g[f"{tag}.mem11"]) < 5e-5
    assert np.array_equal(mem[1][:, ::97].numpy(), g[f"{tag}.labels"])
    _, pm = dec(x[None], pos[None], ts[None], mem, render=True)
    assert rel(digest(pm), g[f"{tag}.pm_render"]) < 5e-5


def test_oracle_reproduces_c2_chain_golden():
    """The benchmarked C2 schedule (10 views 224x224, [2]+[1]*8, render 10) through must3r_b200.engine with the oracle
    model, against the digest of the UNMODIFIED reference engine's run (tests/golden/chain_digest.npz): pins the oracle
    AND the engine's scheduling on a full-size chain (the 512x384 / 20-view twin is checked on the GPU only: minutes of CPU)."""
    from must3r_b200 import engine
    from helpers import full_oracle, digest
    g = load_golden("chain_digest.npz")
    enc, dec = full_oracle(224, 0)
    imgs, ts = syn.synthetic_views(10, 224, 224, seed=2)
    views, tss, ids = list(imgs.unbind(0)), list(ts.unbind(0)), [torch.tensor(i) for i in range(10)]
    mem, pm0, pm = engine.inference_multi_ar(enc, dec, views, ids, tss, [2] + [1] * 8, post_process_function=lambda p: {"raw": p},
                                             device="cpu", return_mem=True)
    raw_r = torch.stack([d["raw"] for d in pm])
    assert rel(digest(raw_r, 65536), g["c2.raw_render"]) < 3e-5
    assert rel(digest(torch.stack([d["raw"] for d in pm0]), 65536), g["c2.raw_first"]) < 3e-5
    assert rel(digest(mem[0][11], 16384), g["c2.mem11"]) < 3e-5
    assert np.array_equal(mem[1][:, ::193].numpy(), g["c2.labels"])
