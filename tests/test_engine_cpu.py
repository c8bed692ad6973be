"""Host logic of must3r_b200.engine (view grouping, update/refine/evict bookkeeping, chunked render) pinned
against the reference engine's outputs (tests/golden/engine.npz), with the CPU oracle standing in as the model.
fp32 vs fp32: 3e-5 rel-L2."""
import os

import numpy as np
import pytest
import torch

from helpers import load_golden, tiny_oracle, rel
from must3r_b200 import engine, synthetic as syn
from oracle import must3r_oracle as orc

TOL = 3e-5


def _views():
    imgs, tss = [], []
    for i in range(6):
        H, W = (32, 48) if i % 3 != 2 else (48, 32)
        im, ts = syn.synthetic_views(1, H, W, seed=100 + i)
        imgs.append(im[0])
        tss.append(ts[0])
    return imgs, tss


def _check_mem(g, prefix, mem):
    for l in range(3):
        assert rel(mem[0][l], g[f"{prefix}.mem{l}"]) < TOL
    assert np.array_equal(mem[1].numpy(), g[f"{prefix}.labels"])
    assert [int(v) for v in mem[2:]] == g[f"{prefix}.tail"].tolist()


def test_inference_multi_ar_with_refinement():
    g = load_golden("engine.npz")
    enc, dec = tiny_oracle(7)
    imgs, tss = _views()
    ids = [torch.tensor(i) for i in range(6)]
    mem, pm0, pm = engine.inference_multi_ar(enc, dec, imgs, ids, tss, [2, 1, 1], max_bs=2,
                                             post_process_function=orc.postprocess, device="cpu", return_mem=True,
                                             num_refinements_iterations=1)
    assert len(pm0) == 4 and len(pm) == 6
    for i, d in enumerate(pm0):
        for k, v in d.items():
            assert rel(v, g[f"multi_ar.pm0.{i}.{k}"]) < TOL, (i, k)
    for i, d in enumerate(pm):
        for k, v in d.items():
            assert rel(v, g[f"multi_ar.pm.{i}.{k}"]) < TOL, (i, k)
    _check_mem(g, "multi_ar.mem", mem)


def test_inference_multi_ar_precomputed_mem_and_to_render():
    """Render-only call on a memory built earlier, for a subset of the views in a caller-chosen order."""
    g = load_golden("engine.npz")
    enc, dec = tiny_oracle(7)
    imgs, tss = _views()
    ids = [torch.tensor(i) for i in range(6)]
    mem, _, _ = engine.inference_multi_ar(enc, dec, imgs, ids, tss, [2, 1, 1], max_bs=2,
                                          post_process_function=orc.postprocess, device="cpu", return_mem=True,
                                          num_refinements_iterations=1)
    first, pm = engine.inference_multi_ar(enc, dec, imgs, ids, tss, [2, 1, 1], to_render=[5, 0, 2], precomputed_mem=mem,
                                          post_process_function=orc.postprocess, device="cpu")
    assert first is None and len(pm) == 3
    for i, d in enumerate(pm):
        for k, v in d.items():
            assert rel(v, g[f"multi_ar.pm_sel.{i}.{k}"]) < TOL, (i, k)


def test_inference_video_with_refinement_pass():
    """Streaming schedule with a second pass: keyframes are refreshed in place, non-keyframes evicted between passes."""
    g = load_golden("engine.npz")
    enc, dec = tiny_oracle(7)
    imgs, tss = _views()
    mem, pm0 = engine.inference_video_multi_ar(enc, dec, imgs, tss, [2, 1, 1, 1, 1], post_process_function=orc.postprocess,
                                               device="cpu", return_mem=True, local_context_size=3,
                                               num_refinements_iterations=1)
    for i, d in enumerate(pm0):
        for k, v in d.items():
            assert rel(v, g[f"video_ref.pm0.{i}.{k}"]) < TOL, (i, k)
    _check_mem(g, "video_ref.mem", mem)


def test_concat_preds_and_groupby_consecutive():
    a = {"pts3d": torch.zeros(2, 1, 4, 4, 3), "conf": torch.ones(2, 1, 4, 4)}
    b = {"pts3d": torch.ones(2, 3, 4, 4, 3), "conf": torch.zeros(2, 3, 4, 4), "extra": torch.zeros(1)}
    out = engine.concat_preds(a, b)                                  # engine/inference.py:691-695: cat along the view axis
    assert out["pts3d"].shape == (2, 4, 4, 4, 3) and out["conf"].shape == (2, 4, 4, 4) and out["extra"].shape == (1,)
    assert float(out["pts3d"][:, 0].abs().sum()) == 0 and float(out["pts3d"][:, 1:].min()) == 1
    assert engine.groupby_consecutive([5, 1, 2, 3, 9, 8]) == [(1, 3), (5, 5), (8, 9)]
    assert engine.groupby_consecutive([]) == []


def test_inference_video_rolling_window():
    g = load_golden("engine.npz")
    enc, dec = tiny_oracle(7)
    imgs, tss = _views()
    mem, pm0 = engine.inference_video_multi_ar(enc, dec, imgs, tss, [2, 1, 1, 1, 1], post_process_function=orc.postprocess,
                                               device="cpu", return_mem=True, local_context_size=2)
    for i, d in enumerate(pm0):
        for k, v in d.items():
            assert rel(v, g[f"video.pm0.{i}.{k}"]) < TOL, (i, k)
    _check_mem(g, "video.mem", mem)


def test_host_label_shadow_equals_device_masks(monkeypatch):
    """The slice-based memory edits (host label shadow: no device sync, zero-copy tail eviction) must give exactly what the
    reference's boolean-mask edits give: streaming schedule with eviction, refinement passes and two aspect ratios."""
    import importlib
    inf = importlib.import_module("must3r_b200.engine.inference")      # (the package also exports a function `inference`)
    enc, dec = tiny_oracle(7)
    imgs, tss = _views()
    kw = dict(post_process_function=orc.postprocess, device="cpu", return_mem=True, local_context_size=2,
              num_refinements_iterations=1)
    calls = {"slices": 0}
    real_runs = inf._runs

    def counting_runs(mask):
        calls["slices"] += 1
        return real_runs(mask)

    monkeypatch.setattr(inf, "_runs", counting_runs)
    mem_a, pm_a = engine.inference_video_multi_ar(enc, dec, imgs, tss, [2, 1, 1, 1, 1], **kw)
    assert calls["slices"] > 0                                        # the shadow path really ran
    sh = inf._host_labels(mem_a[1])
    assert sh is not None and np.array_equal(sh, mem_a[1][0].numpy())  # shadow == device labels at the end
    assert all(v.untyped_storage().nbytes() == v.numel() * v.element_size() for v in mem_a[0])   # compact storage
    import io
    buf = io.BytesIO()
    torch.save(tuple(mem_a), buf)                                     # a saved memory loads with the default (weights_only) loader
    buf.seek(0)
    back = torch.load(buf)
    assert torch.equal(back[1], mem_a[1]) and np.array_equal(inf._host_labels(back[1]), sh)
    monkeypatch.setattr(inf, "_shadow_after_call", lambda mem_before, new_mem, idx_st, x_st: new_mem)
    mem_b, pm_b = engine.inference_video_multi_ar(enc, dec, imgs, tss, [2, 1, 1, 1, 1], **kw)
    assert inf._host_labels(mem_b[1]) is None
    assert torch.equal(mem_a[1], mem_b[1]) and [int(v) for v in mem_a[2:]] == [int(v) for v in mem_b[2:]]
    for va, vb in zip(mem_a[0], mem_b[0]):
        assert torch.equal(va, vb)
    for da, db in zip(pm_a, pm_b):
        for k in da:
            assert torch.equal(da[k], db[k])
    # offline keyframes + refinement (inference_multi_ar refresh path)
    ids = [torch.tensor(i) for i in range(6)]
    kw2 = dict(max_bs=2, post_process_function=orc.postprocess, device="cpu", return_mem=True, num_refinements_iterations=2)
    mem_d, _, pm_d = engine.inference_multi_ar(enc, dec, imgs, ids, tss, [2, 1, 1], **kw2)
    monkeypatch.undo()
    mem_c, _, pm_c = engine.inference_multi_ar(enc, dec, imgs, ids, tss, [2, 1, 1], **kw2)
    for va, vb in zip(mem_c[0], mem_d[0]):
        assert torch.equal(va, vb)
    for da, db in zip(pm_c, pm_d):
        for k in da:
            assert torch.equal(da[k], db[k])


def test_inference_tensor_path_chunked_render():
    g = load_golden("engine.npz")
    enc, dec = tiny_oracle(7)
    im, ts = syn.synthetic_views(8, 32, 48, seed=200)
    pm0, pm = engine.inference(enc, dec, im.view(2, 4, 3, 32, 48), ts.view(2, 4, 2), [2, 1, 1], max_bs=3)
    assert rel(pm0, g["inference.pm0"]) < TOL and rel(pm, g["inference.pm"]) < TOL
    _, pm = engine.inference(enc, dec, im.view(2, 4, 3, 32, 48), ts.view(2, 4, 2), [2, 1, 1], to_render=[1, 3])
    assert rel(pm, g["inference.pm_sel"]) < TOL


def test_postprocess_compute_cam():
    """postprocess(compute_cam=True) (engine/inference.py:29-47): focal against the reference's estimator output, pose against
    the oracle's float64 Kabsch on the same points, and a synthetic scene whose camera is known."""
    g = load_golden("tiny_model.npz")
    from must3r_b200.engine import camera
    f = camera.estimate_focal_weiszfeld(torch.from_numpy(g["cam.pts_local"]), torch.tensor((24.0, 16.0)))
    assert np.allclose(f.numpy(), g["cam.focal"], rtol=1e-4)
    pm = torch.from_numpy(g["kv.pm_render"])                          # [1, 5, 32, 48, 7]
    out = engine.postprocess(pm, "norm_exp", compute_cam=True)
    assert out["focal"].shape == (1, 5) and out["c2w"].shape == (1, 5, 4, 4)
    assert np.allclose(out["focal"].numpy().ravel(), g["kv.post.focal"], rtol=2e-3, atol=1e-4)
    for v in range(5):
        R, t = orc.rigid_registration(out["pts3d_local"][0, v].reshape(-1, 3), out["pts3d"][0, v].reshape(-1, 3),
                                      out["conf"][0, v].reshape(-1) - 1.0)
        assert torch.allclose(out["c2w"][0, v, :3, :3].double(), R, atol=2e-4)
        assert torch.allclose(out["c2w"][0, v, :3, 3].double(), t, atol=2e-4)
        assert torch.equal(out["c2w"][0, v, 3], torch.tensor([0.0, 0.0, 0.0, 1.0]))
    # known camera: world points = R0 * local + t0, focal 55
    loc = torch.from_numpy(g["cam.pts_local"][1])
    ang = 0.4
    R0 = torch.tensor([[np.cos(ang), 0.0, np.sin(ang)], [0.0, 1.0, 0.0], [-np.sin(ang), 0.0, np.cos(ang)]], dtype=torch.float32)
    t0 = torch.tensor([1.0, 2.0, -0.5])
    res = camera.camera_from_pointmaps({"pts3d": (loc @ R0.T + t0)[None], "pts3d_local": loc[None],
                                        "conf": 1.0 + torch.rand(1, 32, 48)})
    assert abs(float(res["focal"][0]) - 55.0) < 0.1
    assert torch.allclose(res["c2w"][0, :3, :3], R0, atol=1e-4) and torch.allclose(res["c2w"][0, :3, 3], t0, atol=1e-3)
    with pytest.raises(KeyError):
        engine.postprocess(pm[..., :3], "norm_exp", compute_cam=True)


def test_stack_views_groups_and_none_handling():
    ts = torch.tensor([[32, 48], [48, 32], [32, 48], [32, 48], [48, 32]])
    vals = [torch.full((2,), float(i)) for i in range(5)]
    vals_missing = list(vals)
    vals_missing[2] = None
    shapes, idx, v = engine.stack_views(ts, [vals], max_bs=2)
    assert idx == [[0, 2], [3], [1, 4]]
    assert [s.tolist() for s in shapes] == [[[32, 48], [32, 48]], [[32, 48]], [[48, 32], [48, 32]]]
    assert torch.equal(v[0], torch.stack([vals[0], vals[2]]))
    shapes, idx, v = engine.stack_views(ts, [vals_missing])
    assert idx == [[0, 3], [1, 4], [2]] and v[2] is None and v[0].shape[0] == 2


def test_postprocess_cpu_matches_oracle():
    pm = torch.randn(2, 8, 8, 7)
    out = engine.postprocess(pm)
    ref = orc.postprocess(pm)
    for k in ref:
        assert torch.allclose(out[k], ref[k])


def test_memory_pickle_roundtrip_in_reference_layout(tmp_path):
    """engine.save_memory / load_memory write the (memory, keyframe_pointmaps, overlap_tree) pickle of
    must3r/slam/model.py:431-440; a memory made of views into larger buffers (evicted stream, growable arena) is stored
    compactly and can be resumed: continuing the stream from the loaded memory gives the results of the uninterrupted run."""
    import pickle as pkl
    from must3r_b200 import engine
    enc, dec = tiny_oracle(7)
    imgs, ts = syn.synthetic_views(8, 32, 48, seed=51)
    views, tss = list(imgs.unbind(0)), list(ts.unbind(0))
    kw = dict(post_process_function=lambda p: {"raw": p}, device="cpu", return_mem=True, local_context_size=3)
    mem_all, out_all = engine.inference_video_multi_ar(enc, dec, views, tss, [2] + [1] * 6, **kw)
    mem5, _ = engine.inference_video_multi_ar(enc, dec, views[:5], tss[:5], [2] + [1] * 3, **kw)
    path = os.path.join(tmp_path, "mem.pkl")
    engine.save_memory(path, mem5, keyframe_pointmaps={"n": 5})
    raw = pkl.load(open(path, "rb"))
    assert isinstance(raw, tuple) and len(raw) == 3 and raw[1] == {"n": 5} and raw[2] is None
    for v in raw[0][0]:
        assert v.is_contiguous() and v.untyped_storage().nbytes() == v.numel() * v.element_size() and not hasattr(v, "_m3r_arena")
    mem, data, tree = engine.load_memory(path)
    assert torch.equal(mem[1], mem5[1]) and all(torch.equal(a, b) for a, b in zip(mem[0], mem5[0])) and list(mem[2:]) == [int(v) for v in mem5[2:]]
    x, pos = enc(imgs[5:6], ts[5:6])
    m_a, p_a = dec(x[None], pos[None], ts[None, 5:6], tuple(mem))
    m_b, p_b = dec(x[None], pos[None], ts[None, 5:6], tuple(mem5))
    assert torch.equal(p_a, p_b)


def test_rigid_registration_horn_equals_weighted_kabsch():
    """The cuSOLVER-free rotation (Horn's quaternion solution through repeated 4x4 squaring) against weighted Kabsch via SVD:
    exact poses, noisy ones, a batch, near-planar clouds and a configuration whose unconstrained optimum is a reflection."""
    from must3r_b200.engine import camera
    g = torch.Generator().manual_seed(5)
    for case in range(8):
        n = 500
        x = torch.randn(3, n, 3, generator=g) * torch.tensor([1.0, 1.0, 0.02 if case % 3 == 2 else 1.0])
        A = torch.randn(3, 3, 3, generator=g)
        Q, _ = torch.linalg.qr(A)
        Q = Q * torch.sign(torch.linalg.det(Q))[:, None, None]
        if case == 5:
            Q[:, :, 2] *= -1                                           # a reflection: the best PROPER rotation is not Q
        t0 = torch.randn(3, 3, generator=g)
        y = x @ Q.transpose(-1, -2) + t0[:, None, :] + (0.05 * torch.randn(3, n, 3, generator=g) if case % 2 else 0.0)
        w = torch.rand(3, n, generator=g) + 0.1
        Rh, th = camera.rigid_points_registration(x, y, w)
        Rs, ts_ = camera.rigid_points_registration(x, y, w, method="svd")
        assert torch.allclose(torch.linalg.det(Rh), torch.ones(3), atol=1e-5)
        assert torch.allclose(Rh @ Rh.transpose(-1, -2), torch.eye(3).expand(3, 3, 3), atol=1e-5)
        cost = lambda R, t: (w * (x @ R.transpose(-1, -2) + t[:, None] - y).square().sum(-1)).sum(-1)  # noqa: E731
        assert torch.all(cost(Rh, th) <= cost(Rs, ts_) * (1 + 1e-4) + 1e-6), case     # the same minimum ...
        if case != 5 and case % 3 != 2:
            assert torch.allclose(Rh, Rs, atol=2e-4) and torch.allclose(th, ts_, atol=2e-4), case    # ... and the same minimiser when it is unique


class _ArenaOracleDecoder:
    """The oracle decoder behind the growable-memory contract of the CUDA decoder (model/decoder.py MemArena, reserve_memory):
    update calls whose input is the latest version of an arena with room left append their rows in place, everything else
    gets a fresh arena and a copy.  Lets the CPU suite drive the engine's arena-aware memory edits (tail drops, in-place
    compaction, keyframe refresh, released tails) that otherwise only run on the GPU."""

    def __init__(self, dec):
        self.dec, self._reserve_tokens, self._growth = dec, 0, 0.0
        self.inplace_appends = self.fresh_arenas = 0

    def reserve_memory(self, n_tokens=0, growth=0.0):
        self._reserve_tokens, self._growth = int(n_tokens), float(growth)

    def __call__(self, x, pos, true_shape, current_mem=None, render=False):
        from must3r_b200.model.decoder import MemArena
        new_mem, pms = self.dec(x, pos, true_shape, current_mem, render=render)
        if render:
            return current_mem, pms
        Nm = 0 if current_mem is None else current_mem[0][0].shape[1]
        rows = new_mem[0][0].shape[1]
        arena = MemArena.of(current_mem[0]) if Nm > 0 else None
        if arena is None or arena.tail != Nm or rows > arena.cap:
            cap = max(rows, self._reserve_tokens, int(rows * self._growth))
            arena = MemArena([torch.full((v.shape[0], cap, v.shape[2]), float("nan")) for v in new_mem[0]], cap)
            for buf, v in zip(arena.bufs, new_mem[0]):
                buf[:, :Nm] = v[:, :Nm]
            self.fresh_arenas += 1
        else:
            self.inplace_appends += 1
        for buf, v in zip(arena.bufs, new_mem[0]):
            buf[:, Nm:rows] = v[:, Nm:]
        arena.tail = rows
        return (arena.views(rows), new_mem[1]) + tuple(new_mem[2:]), pms


@pytest.mark.parametrize("schedule", ["video", "offline"])
def test_engine_memory_edits_on_growable_buffers(schedule):
    """Streaming schedule (window eviction from the middle of the memory, non-keyframe tail drops, a refinement pass with
    keyframe refresh) and offline schedule (refinement: discarded appended tails) on in-place growing memory buffers: results
    and final memory equal the plain per-call concatenation, and the appends really happen in place."""
    enc, dec = tiny_oracle(7)
    F = 12
    imgs, ts = syn.synthetic_views(F, 32, 48, seed=61)
    views, tss = list(imgs.unbind(0)), list(ts.unbind(0))
    raw = lambda p: {"raw": p}  # noqa: E731
    arena_dec = _ArenaOracleDecoder(dec)
    if schedule == "video":
        kw = dict(post_process_function=raw, device="cpu", return_mem=True, local_context_size=3, num_refinements_iterations=1)
        mem_a, out_a = engine.inference_video_multi_ar(enc, arena_dec, views, tss, [2] + [1] * (F - 2), **kw)
        mem_b, out_b = engine.inference_video_multi_ar(enc, dec, views, tss, [2] + [1] * (F - 2), **kw)
    else:
        ids = [torch.tensor(i) for i in range(F)]
        kw = dict(post_process_function=raw, device="cpu", return_mem=True, num_refinements_iterations=1)
        mem_a, _, out_a = engine.inference_multi_ar(enc, arena_dec, views, ids, tss, [2, 1, 1, 2, 1], **kw)
        mem_b, _, out_b = engine.inference_multi_ar(enc, dec, views, ids, tss, [2, 1, 1, 2, 1], **kw)
    assert torch.equal(mem_a[1], mem_b[1]) and [int(v) for v in mem_a[2:]] == [int(v) for v in mem_b[2:]]
    for a, b in zip(mem_a[0], mem_b[0]):
        assert torch.equal(a, b)
    for a, b in zip(out_a, out_b):
        assert torch.equal(a["raw"], b["raw"])
    # (on CPU the features are encoded step by step, so the engine cannot size the buffers up front: geometric growth 1.5x)
    assert arena_dec.inplace_appends >= 2 * arena_dec.fresh_arenas, (arena_dec.inplace_appends, arena_dec.fresh_arenas)
