"""Model-level parity of the CUDA path (through load_model / Dust3rEncoder / MUSt3R, i.e. the C ABI whole-model
entry points) against (i) outputs of the unmodified reference (tests/golden/*.npz) and (ii) the CPU oracle.

Tolerances (rel-L2 of raw pointmaps / memory vs the fp32 reference), see DESIGN.md "Numerics":
  tiny models (128-wide, 2-3 layers, errors do not average out): fp16 3e-3, bf16 2.5e-2
  full-size models (the benchmarked schedules, test_benchmarked_chains_vs_reference_engine and the 3-view digests):
    fp16 operands 1.0e-3 (1.2e-3 on the 3-view digests' sampled statistics), bf16 1.1e-2 = the reference's own bf16-vs-fp32 gap
"""
import argparse
import os

import numpy as np
import pytest
import torch

from helpers import load_golden, tiny_oracle, full_oracle, digest, rel, TINY_ENC, TINY_DEC
from must3r_b200 import synthetic as syn, engine
from must3r_b200.model import Dust3rEncoder, MUSt3R, load_model, set_precision, ActivationType

pytestmark = pytest.mark.gpu
TOL = {torch.float16: 3e-3, torch.bfloat16: 2.5e-2}
DT = [torch.float16, torch.bfloat16]


def tiny_cuda(seed=7, enc_kw=None, dec_kw=None):
    enc = Dust3rEncoder(img_size=(64, 64), embed_dim=128, depth=2, num_heads=2, **(enc_kw or {}))
    dkw = dict(img_size=(64, 64), enc_embed_dim=128, embed_dim=128, depth=3, num_heads=2, output_dim=1792,
               feedback_type="single_mlp", memory_mode="kv", landscape_only=False)
    dkw.update(dec_kw or {})
    dec = MUSt3R(**dkw)
    enc.load_state_dict(syn.encoder_state_dict(seed, embed_dim=128, depth=2), strict=True)
    dec.load_state_dict(syn.decoder_state_dict(seed, enc_embed_dim=128, embed_dim=128, depth=3,
                                               feedback_type=dkw["feedback_type"]), strict=True)
    return enc.cuda().eval(), dec.cuda().eval()


def check_mem(g, prefix, mem, tol, depth=3):
    for l in range(depth):
        assert rel(mem[0][l].float().cpu(), g[f"{prefix}.mem{l}"]) < tol, (prefix, l)
    assert np.array_equal(mem[1].cpu().numpy(), g[f"{prefix}.labels"])
    assert [int(v) for v in mem[2:]] == g[f"{prefix}.tail"].tolist()


VARIANTS = {
    "kv": ({}, {}),
    "f0": (dict(pos_embed="RoPE100_224:512"), dict(pos_embed="RoPE100_224:512")),
    "nofb": ({}, dict(feedback_type=None)),
    "fblin": ({}, dict(feedback_type="single_linear")),
    "normy": ({}, dict(memory_mode="norm_y")),          # memory rows = norm_y(x), K|V projected at use (layers.py:85,94)
    "raw": ({}, dict(memory_mode="raw")),               # memory rows = x, norm_y + projection at use (layers.py:82,92)
}


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("variant", list(VARIANTS))
def test_tiny_chain_vs_reference(dtype, variant):
    """init(2 views) -> update(1) -> update(2, masked with Nm>0) -> render(5), against the reference's outputs."""
    set_precision(dtype)
    tol = TOL[dtype]
    g = load_golden("tiny_model.npz")
    enc, dec = tiny_cuda(7, *VARIANTS[variant])
    imgs, ts = syn.synthetic_views(5, 32, 48, seed=11)
    imgs, ts = imgs.cuda(), ts.cuda()
    x, pos = enc(imgs, ts)
    assert x.dtype == torch.float32 and rel(x.cpu(), g[f"{variant}.enc_x"]) < tol
    assert np.array_equal(pos.cpu().numpy(), g[f"{variant}.enc_pos"])
    mem, pm = dec(x[None, 0:2], pos[None, 0:2], ts[None, 0:2], None)
    assert pm.dtype == torch.float32 and rel(pm.cpu(), g[f"{variant}.pm_init"]) < tol
    check_mem(g, f"{variant}.init", mem, tol)
    mem, pm = dec(x[None, 2:3], pos[None, 2:3], ts[None, 2:3], mem)
    assert rel(pm.cpu(), g[f"{variant}.pm_upd1"]) < tol
    mem, pm = dec(x[None, 3:5], pos[None, 3:5], ts[None, 3:5], mem)
    assert rel(pm.cpu(), g[f"{variant}.pm_upd2"]) < tol
    check_mem(g, f"{variant}.final", mem, tol)
    mem_r, pm = dec(x[None], pos[None], ts[None], mem, render=True)
    assert rel(pm.cpu(), g[f"{variant}.pm_render"]) < tol
    assert mem_r[0][0] is mem[0][0]
    out = engine.postprocess(pm, ActivationType.NORM_EXP)
    for k in ("pts3d", "pts3d_local", "conf"):
        assert rel(out[k].cpu(), g[f"{variant}.post.{k}"] if variant == "kv" else out[k].cpu()) < 2 * tol


@pytest.mark.parametrize("dtype", DT)
def test_tiny_batch2_and_list_form(dtype):
    set_precision(dtype)
    tol = TOL[dtype]
    g = load_golden("tiny_model.npz")
    enc, dec = tiny_cuda(7)
    imgs, ts = syn.synthetic_views(6, 32, 48, seed=12)
    x, pos = enc(imgs.cuda(), ts.cuda())
    ts = ts.cuda()
    xb, pb, tb = x.view(2, 3, *x.shape[1:]), pos.view(2, 3, *pos.shape[1:]), ts.view(2, 3, 2)
    mem, pm = dec(xb[:, :1], pb[:, :1], tb[:, :1], None)             # single-image init: no own-mask
    assert rel(pm.cpu(), g["b2.pm_init1"]) < tol
    mem, pm = dec(xb[:, 1:3], pb[:, 1:3], tb[:, 1:3], mem)
    assert rel(pm.cpu(), g["b2.pm_upd2"]) < tol
    check_mem(g, "b2.final", mem, tol)
    _, pm = dec(xb, pb, tb, mem, render=True)
    assert rel(pm.cpu(), g["b2.pm_render"]) < tol
    # two aspect ratios in one call
    imgs_p, ts_p = syn.synthetic_views(2, 48, 32, seed=13)
    imgs_l, ts_l = syn.synthetic_views(2, 32, 48, seed=14)
    xp, pp = enc(imgs_p.cuda(), ts_p.cuda())
    xl, pl = enc(imgs_l.cuda(), ts_l.cuda())
    ts_p, ts_l = ts_p.cuda(), ts_l.cuda()
    mem, pms = dec([xl[None], xp[None]], [pl[None], pp[None]], [ts_l[None], ts_p[None]], None)
    assert rel(pms[0].cpu(), g["list.pm0"]) < tol and rel(pms[1].cpu(), g["list.pm1"]) < tol
    check_mem(g, "list.init", mem, tol)
    mem2, pms = dec([xp[None, :1], xl[None, :1]], [pp[None, :1], pl[None, :1]], [ts_p[None, :1], ts_l[None, :1]], mem)
    assert rel(pms[0].cpu(), g["list.pm2"]) < tol and rel(pms[1].cpu(), g["list.pm3"]) < tol
    check_mem(g, "list.final", mem2, tol)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("tag,H,W,size", [("224", 224, 224, 224), ("512", 384, 512, 512)])
def test_full_size_vs_reference_digest(dtype, tag, H, W, size, tmp_path):
    """Real architecture (ViT-L encoder, ViT-B decoder) through load_model on a synthetic checkpoint file written
    in the reference's format; compared with digests of the reference's outputs."""
    set_precision(dtype)
    tol = {torch.float16: 1.2e-3, torch.bfloat16: 1.1e-2}[dtype]
    g = load_golden("full_model_digest.npz")
    ck = {"args": argparse.Namespace(
              encoder=f"Dust3rEncoder(img_size=({size}, {size}), patch_embed='PatchEmbedDust3R')",
              decoder=f"CausalMUSt3R(img_size=({size}, {size}), feedback_type='single_mlp', memory_mode='kv', mem_dropout=0.1)"),
          "encoder": syn.encoder_state_dict(0), "decoder": syn.decoder_state_dict(0)}
    path = os.path.join(tmp_path, "ckpt.pth")
    torch.save(ck, path)
    enc, dec = load_model(path, device="cuda", verbose=False)
    imgs, ts = syn.synthetic_views(3, H, W, seed=2)
    imgs, ts = imgs.cuda(), ts.cuda()
    x, pos = enc(imgs, ts)
    assert rel(digest(x), g[f"{tag}.enc_x"]) < tol
    mem, pm = dec(x[None, :2], pos[None, :2], ts[None, :2], None)
    assert rel(digest(pm), g[f"{tag}.pm_init"]) < tol
    mem, pm = dec(x[None, 2:3], pos[None, 2:3], ts[None, 2:3], mem)
    assert rel(digest(pm), g[f"{tag}.pm_upd"]) < tol
    assert rel(digest(mem[0][0]), g[f"{tag}.mem0"]) < tol
    assert rel(digest(mem[0][11]), g[f"{tag}.mem11"]) < tol
    assert np.array_equal(mem[1][:, ::97].cpu().numpy(), g[f"{tag}.labels"])
    _, pm = dec(x[None], pos[None], ts[None], mem, render=True)
    assert rel(digest(pm), g[f"{tag}.pm_render"]) < tol


@pytest.mark.parametrize("dtype", DT)
def test_engine_on_cuda_vs_reference_engine(dtype):
    """must3r_b200.engine driving the CUDA model, against the reference engine's outputs (engine.npz)."""
    set_precision(dtype)
    tol = 2 * TOL[dtype]
    g = load_golden("engine.npz")
    enc, dec = tiny_cuda(7)
    imgs, tss = [], []
    for i in range(6):
        H, W = (32, 48) if i % 3 != 2 else (48, 32)
        im, ts = syn.synthetic_views(1, H, W, seed=100 + i)
        imgs.append(im[0].cuda())
        tss.append(ts[0].cuda())
    ids = [torch.tensor(i) for i in range(6)]
    pp = lambda pm: engine.postprocess(pm, ActivationType.NORM_EXP)  # noqa: E731
    mem, pm0, pm = engine.inference_multi_ar(enc, dec, imgs, ids, tss, [2, 1, 1], max_bs=2, post_process_function=pp,
                                             device="cuda", return_mem=True, num_refinements_iterations=1)
    for i, d in enumerate(pm):
        for k, v in d.items():
            assert rel(v.cpu(), g[f"multi_ar.pm.{i}.{k}"]) < tol, (i, k)
    assert np.array_equal(mem[1].cpu().numpy(), g["multi_ar.mem.labels"])
    mem, pm0 = engine.inference_video_multi_ar(enc, dec, imgs, tss, [2, 1, 1, 1, 1], post_process_function=pp,
                                               device="cuda", return_mem=True, local_context_size=2)
    for i, d in enumerate(pm0):
        for k, v in d.items():
            assert rel(v.cpu(), g[f"video.pm0.{i}.{k}"]) < tol, (i, k)
    assert np.array_equal(mem[1].cpu().numpy(), g["video.mem.labels"])
    im, ts = syn.synthetic_views(8, 32, 48, seed=200)
    pm0, pm = engine.inference(enc, dec, im.view(2, 4, 3, 32, 48).cuda(), ts.view(2, 4, 2).cuda(), [2, 1, 1], max_bs=3)
    assert rel(pm0.cpu(), g["inference.pm0"]) < tol and rel(pm.cpu(), g["inference.pm"]) < tol


@pytest.mark.parametrize("dtype", DT)
def test_many_ar_patch_embed_vs_reference(dtype):
    """ManyAR_PatchEmbed (landscape-stored batch with a portrait view, dust3r/dust3r/patch_embed.py:32-70)."""
    set_precision(dtype)
    g = load_golden("tiny_model.npz")
    enc, _ = tiny_cuda(7, dict(patch_embed="ManyAR_PatchEmbed"))
    imgs, _ = syn.synthetic_views(3, 32, 48, seed=15)
    ts = torch.tensor([[32, 48], [48, 32], [32, 48]], dtype=torch.int64)
    x, pos = enc(imgs.cuda(), ts.cuda())
    assert rel(x.cpu(), g["manyar.enc_x"]) < TOL[dtype]
    assert np.array_equal(pos.cpu().numpy(), g["manyar.enc_pos"])


def test_cuda_vs_oracle_same_box():
    """The CPU oracle (pinned to the reference by test_oracle_golden) evaluated on this box vs the CUDA path."""
    set_precision(torch.float16)
    enc_o, dec_o = tiny_oracle(3)
    enc, dec = tiny_cuda(3)
    imgs, ts = syn.synthetic_views(3, 48, 64, seed=5)
    xo, po = enc_o(imgs, ts)
    x, pos = enc(imgs.cuda(), ts.cuda())
    assert rel(x.cpu(), xo) < TOL[torch.float16]
    mo, pmo = dec_o(xo[None], po[None], ts[None], None)
    m, pm = dec(x[None], pos[None], ts.cuda()[None], None)
    assert rel(pm.cpu(), pmo) < TOL[torch.float16]
    assert rel(m[0][2].float().cpu(), mo[0][2]) < TOL[torch.float16]


def test_no_cpu_fallback_and_arg_errors():
    enc, dec = tiny_cuda(7)
    imgs, ts = syn.synthetic_views(1, 32, 48, seed=1)
    with pytest.raises(RuntimeError, match="CUDA only"):
        enc(imgs, ts)
    with pytest.raises(AssertionError):
        enc(torch.zeros(1, 3, 40, 48, device="cuda"), ts.cuda())
    x, pos = enc(imgs.cuda(), ts.cuda())
    mem, _ = dec(x[None], pos[None], ts.cuda()[None], None)
    dec.change_memory_mode("norm_y")                      # a K|V memory handed to a norm_y decoder is a width error
    with pytest.raises(AssertionError, match="memory_mode"):
        dec(x[None], pos[None], ts.cuda()[None], mem)
    with pytest.raises(NotImplementedError):
        dec(x[None], pos[None], ts.cuda()[None], None, return_feats=True)


@pytest.mark.parametrize("mode", ["norm_y", "raw"])
def test_memory_modes_match_kv(mode):
    """the three memory modes give the same pointmaps (bit-identical in the fp32 reference; here the stored
    rows are rounded to 16 bit at a different point, so equal within the 16-bit tolerance), incl. batch 2 and an expanded
    (stride-0) memory at render time (engine/inference.py:668-683)."""
    set_precision(torch.float16)
    enc, dec = tiny_cuda(7)
    imgs, ts = syn.synthetic_views(6, 32, 48, seed=12)
    x, pos = enc(imgs.cuda(), ts.cuda())
    ts = ts.cuda()
    xb, pb, tb = x.view(2, 3, *x.shape[1:]), pos.view(2, 3, *pos.shape[1:]), ts.view(2, 3, 2)

    def chain(m):
        dec.change_memory_mode(m)
        mem, pm0 = dec(xb[:, :2], pb[:, :2], tb[:, :2], None)
        mem, pm1 = dec(xb[:, 2:3], pb[:, 2:3], tb[:, 2:3], mem)
        _, pm2 = dec(xb, pb, tb, mem, render=True)
        one = ([v[:1].expand(3, -1, -1) for v in mem[0]], mem[1][:1].expand(3, -1), *mem[2:])
        _, pm3 = dec(xb[0][:, None], pb[0][:, None], tb[0][:, None], one, render=True)     # 3 "scenes" sharing one memory
        return mem, [pm0, pm1, pm2, pm3]

    mem_kv, ref = chain("kv")
    mem_m, got = chain(mode)
    assert mem_m[0][0].shape[2] == dec.embed_dim and mem_kv[0][0].shape[2] == 2 * dec.embed_dim
    for a, b in zip(got, ref):
        assert rel(a.cpu(), b.cpu()) < TOL[torch.float16]
    assert rel(got[3].cpu()[:, 0], got[2].cpu()[0]) < TOL[torch.float16]


# bf16 gate of the full-size chains: the reference's own bf16-autocast-vs-fp32 gap on the raw head output (1.1e-2,
# BASELINE.md §5 / SURVEY.md §0 fact 10); fp16 operands are the north star's 1e-3 parity mode
CHAIN_TOL = {torch.float16: 1.0e-3, torch.bfloat16: 1.1e-2}
# pts3d_local = norm_exp activation of the head's channels 3:6: expm1 amplifies the relative error of the larger |v| of the
# camera-frame points (measured 1.0e-3 at 224x224, 1.2e-3 at 512x384 with fp16 operands vs 7.7e-4 on the raw head output)
CHAIN_TOL_LOCAL = {torch.float16: 1.5e-3, torch.bfloat16: 1.3e-2}


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("tag,V,H,W,size", [("c2", 10, 224, 224, 224), ("c3", 20, 384, 512, 512)])
def test_benchmarked_chains_vs_reference_engine(dtype, tag, V, H, W, size):
    """The schedules bench.py times (C2: 10 views 224x224, C3: 20 views 512x384; mem_batches [2]+[1]*(V-2), render all)
    through must3r_b200.engine, against digests of the UNMODIFIED reference engine's outputs (tests/golden/chain_digest.npz,
    made by tests/golden/make_golden.py chain): error after the full chain of one-view updates, not after 3 views."""
    set_precision(dtype)
    tol = CHAIN_TOL[dtype]
    g = load_golden("chain_digest.npz")
    enc = Dust3rEncoder(img_size=(size, size))
    dec = MUSt3R(img_size=(size, size), feedback_type="single_mlp", memory_mode="kv", landscape_only=False)
    enc.load_state_dict(syn.encoder_state_dict(0)); dec.load_state_dict(syn.decoder_state_dict(0))
    enc, dec = enc.cuda().eval(), dec.cuda().eval()
    imgs, ts = syn.synthetic_views(V, H, W, seed=2)
    views, tss, ids = list(imgs.cuda().unbind(0)), list(ts.unbind(0)), [torch.tensor(i) for i in range(V)]
    raw = lambda pm: {"raw": pm}  # noqa: E731
    mem, pm0, pm = engine.inference_multi_ar(enc, dec, views, ids, tss, [2] + [1] * (V - 2), post_process_function=raw,
                                             device="cuda", return_mem=True)
    raw_r = torch.stack([d["raw"] for d in pm])
    raw_0 = torch.stack([d["raw"] for d in pm0])
    post = engine.postprocess(raw_r, ActivationType.NORM_EXP)
    errs = {"raw_render": rel(digest(raw_r, 65536), g[f"{tag}.raw_render"]), "raw_first": rel(digest(raw_0, 65536), g[f"{tag}.raw_first"]),
            "pts3d": rel(digest(post["pts3d"], 32768), g[f"{tag}.pts3d"]), "pts3d_local": rel(digest(post["pts3d_local"], 32768), g[f"{tag}.pts3d_local"]),
            "conf": rel(digest(post["conf"], 32768), g[f"{tag}.conf"]), "raw_view_last": rel(digest(raw_r[-1], 16384), g[f"{tag}.raw_view_last"]),
            "mem0": rel(digest(mem[0][0], 16384), g[f"{tag}.mem0"]), "mem11": rel(digest(mem[0][11], 16384), g[f"{tag}.mem11"])}
    print(tag, dtype, {k: f"{v:.2e}" for k, v in errs.items()})
    assert np.array_equal(mem[1][:, ::193].cpu().numpy(), g[f"{tag}.labels"])
    assert [int(v) for v in mem[2:]] == g[f"{tag}.tail"].tolist()
    for k, v in errs.items():
        assert v < (CHAIN_TOL_LOCAL[dtype] if k == "pts3d_local" else tol), (k, v)


def test_inplace_append_equals_copy_and_old_versions_survive():
    """Memory buffers with spare room (MUSt3R.reserve_memory): update calls append their rows in place; results are bitwise
    those of the reference-style fresh concatenation, older versions of the memory stay valid, and a call on an OLD version
    (a branch) never overwrites the rows of the newer one."""
    set_precision(torch.float16)
    enc, dec = tiny_cuda(7)
    imgs, ts = syn.synthetic_views(6, 32, 48, seed=11)
    x, pos = enc(imgs.cuda(), ts.cuda())
    ts = ts.cuda()

    def chain():
        mems, pms = [], []
        mem, pm = dec(x[None, :2], pos[None, :2], ts[None, :2], None)
        mems.append(mem); pms.append(pm)
        for i in range(2, 5):
            mem, pm = dec(x[None, i:i + 1], pos[None, i:i + 1], ts[None, i:i + 1], mem)
            mems.append(mem); pms.append(pm)
        return mems, pms

    dec.reserve_memory()
    ref_mems, ref_pms = chain()
    dec.reserve_memory(6 * x.shape[1])
    mems, pms = chain()
    base = mems[0][0][0].data_ptr()
    assert all(m[0][0].data_ptr() == base for m in mems), "updates did not append in place"
    for a, b in zip(pms, ref_pms):
        assert torch.equal(a, b)
    for ma, mb in zip(mems, ref_mems):                       # every version, also the old ones, still reads correctly
        assert torch.equal(ma[1], mb[1]) and all(torch.equal(u, v) for u, v in zip(ma[0], mb[0]))
    # branch from an old version: must not disturb the newest one
    newest = [v.clone() for v in mems[-1][0]]
    br, _ = dec(x[None, 5:6], pos[None, 5:6], ts[None, 5:6], mems[1])
    assert br[0][0].data_ptr() != base
    assert all(torch.equal(u, v) for u, v in zip(mems[-1][0], newest))
    ref_br, _ = dec(x[None, 5:6], pos[None, 5:6], ts[None, 5:6], ref_mems[1])
    assert all(torch.equal(u, v) for u, v in zip(br[0], ref_br[0]))
    dec.reserve_memory()


def test_stream_schedule_inplace_memory_equals_copy_path():
    """inference_video_multi_ar (keyframes, rolling window, eviction from the middle of the memory) with the growable
    in-place memory vs the same schedule with fresh tensors per call: bitwise-identical results and final memory."""
    import importlib
    inf = importlib.import_module("must3r_b200.engine.inference")
    set_precision(torch.float16)
    enc, dec = tiny_cuda(7)
    F = 14
    imgs, ts = syn.synthetic_views(F, 32, 48, seed=31)
    views, tss = list(imgs.cuda().unbind(0)), list(ts.unbind(0))
    pp = lambda pm: engine.postprocess(pm, ActivationType.NORM_EXP)  # noqa: E731

    def run():
        return engine.inference_video_multi_ar(enc, dec, views, tss, [2] + [1] * (F - 2), post_process_function=pp, device="cuda",
                                               return_mem=True, local_context_size=4, num_refinements_iterations=1)
    mem_a, out_a = run()
    real = inf._reserve
    inf._reserve = lambda *a, **k: None
    try:
        dec.reserve_memory()
        mem_b, out_b = run()
    finally:
        inf._reserve = real
    assert torch.equal(mem_a[1], mem_b[1]) and all(torch.equal(a, b) for a, b in zip(mem_a[0], mem_b[0]))
    assert all(torch.equal(a[k], b[k]) for a, b in zip(out_a, out_b) for k in a)


def test_two_host_threads_on_two_streams():
    """The reference's SLAM drives the model from a worker thread while another thread owns the main stream (slam.py:533): two
    host threads running decoder chains at the same time on their own CUDA streams (per-stream workspaces, LayerNorm-emit
    counters and attention split scratch) must reproduce the single-threaded results bit for bit (tools/two_threads.py)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "two_threads.py")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "single thread: True" in out.stdout, out.stdout[-1500:] + out.stderr[-1500:]
