"""LayerNorm fused into the producing GEMM (m3r_gemm_args.norm_out), affine-free normalisation (m3r_normalize16), the
narrow / odd tile widths and the grouped GEMM (m3r_gemm_grouped), against fp32 torch math on the same 16-bit inputs.
Reference call sites replaced: nn.LayerNorm at must3r/model/blocks/layers.py:42,46,65,70,71,76; the per-level K|V
projections of the memory append, must3r/model/decoder.py:323-330."""
import os

import pytest
import torch
import torch.nn.functional as F

from must3r_b200 import ops
from must3r_b200.synthetic import rel_l2

pytestmark = pytest.mark.gpu
DT = [torch.float16, torch.bfloat16]
OUT_TOL = {torch.float16: 6e-4, torch.bfloat16: 5e-3, torch.float32: 2e-5}


def rnd(*shape, dtype=torch.float32, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("M,N,K", [(768, 768, 768), (1536, 768, 3072), (196, 768, 768), (100, 128, 64), (392, 768, 1024), (1372, 768, 768),
                                   (768, 768, 3072), (196, 768, 3072), (900, 512, 2048)])   # the last three: split-K CTA pairs
def test_gemm_emits_layernorm(dtype, M, N, K):
    """x = res + a W^T + b (fp32) and norm_out = (x - mean) / sqrt(var + eps) in one launch; rows with a large common
    offset check that the cross-tile (mean, M2) combination is as good as a two-pass LayerNorm."""
    a, w = rnd(M, K, dtype=dtype, seed=1), rnd(N, K, dtype=dtype, seed=2, scale=K ** -0.5)
    bias, res = rnd(N, seed=3), rnd(M, N, seed=4, scale=3.0)
    res[: M // 2] += 40.0                                          # |mean| >> std on half of the rows
    res[:, 5] *= 30.0                                              # an outlier channel
    x_ref = a.float() @ w.float().t() + bias + res
    n_ref = F.layer_norm(x_ref, (N,), eps=1e-6)
    for _ in range(3):                                             # counters must re-arm between launches
        nout = torch.full((M, N), float("nan"), dtype=dtype, device="cuda")
        x = ops.linear(a, w, bias, residual=res, out_dtype=torch.float32, norm_out=nout, norm_eps=1e-6, w_static=True)
        assert rel_l2(x, x_ref) < OUT_TOL[torch.float32]
        assert torch.isfinite(nout.float()).all()
        assert rel_l2(nout, n_ref) < OUT_TOL[dtype]
    # in place on the residual (cproj / fc2 of the render path)
    r2 = res.clone()
    nout = torch.empty((M, N), dtype=dtype, device="cuda")
    ops.linear(a, w, bias, residual=r2, out=r2, norm_out=nout, norm_eps=1e-6)
    assert rel_l2(r2, x_ref) < OUT_TOL[torch.float32] and rel_l2(nout, n_ref) < OUT_TOL[dtype]


def test_gemm_emit_refuses_multi_wave():
    a, w = rnd(4096, 768, dtype=torch.float16), rnd(768, 768, dtype=torch.float16)
    with pytest.raises(RuntimeError, match="one 128x64 tile per SM"):
        ops.linear(a, w, None, out_dtype=torch.float32, norm_out=torch.empty(4096, 768, dtype=torch.float16, device="cuda"))


@pytest.mark.parametrize("dtype", DT)
def test_emit_chain_matches_layernorm_then_gemm(dtype):
    """GEMM1 (emit) -> GEMM2 with folded affine == GEMM1 -> LayerNorm(gamma, beta) -> GEMM2, up to 16-bit rounding."""
    M, D = 768, 768
    a, w1 = rnd(M, D, dtype=dtype, seed=1), rnd(D, D, dtype=dtype, seed=2, scale=D ** -0.5)
    res = rnd(M, D, seed=3)
    gamma, beta = 1.0 + 0.1 * rnd(D, seed=4), 0.1 * rnd(D, seed=5)
    w2, b2 = rnd(3 * D, D, seed=6, scale=D ** -0.5), rnd(3 * D, seed=7)
    x_ref = a.float() @ w1.float().t() + res
    y_ref = F.layer_norm(x_ref, (D,), gamma, beta, 1e-6) @ w2.t() + b2
    h = torch.empty((M, D), dtype=dtype, device="cuda")
    ops.linear(a, w1, None, residual=res, out_dtype=torch.float32, norm_out=h)
    w2f = (w2 * gamma[None, :]).to(dtype)
    b2f = b2 + w2 @ beta
    y = ops.linear(h, w2f, b2f, out_dtype=torch.float32)
    assert rel_l2(y, y_ref) < (1.2e-3 if dtype == torch.float16 else 8e-3)


@pytest.mark.parametrize("dtype", DT)
def test_normalize16_with_periodic_add(dtype):
    P, D, L = 200, 768, 5
    x = rnd(L * P, D, seed=1, scale=2.0) + 3.0
    off = rnd(P, D, seed=2)
    out = ops.normalize16(x, 1e-5, dtype, add=off, add_rows=(L - 1) * P)
    xr = x.clone().view(L, P, D)
    xr[: L - 1] += off
    ref = F.layer_norm(xr.view(L * P, D), (D,), eps=1e-5)
    assert rel_l2(out, ref) < OUT_TOL[dtype]
    out2 = ops.normalize16(x, 1e-6, dtype)
    assert rel_l2(out2, F.layer_norm(x, (D,), eps=1e-6)) < OUT_TOL[dtype]


@pytest.mark.parametrize("bn", [32, 160, 192])
@pytest.mark.parametrize("M,N,K", [(768, 3840, 768), (300, 960, 128), (1000, 1920, 256)])
def test_gemm_new_tile_widths(bn, M, N, K):
    if N % bn:
        pytest.skip("N not a multiple of this tile width")
    dtype = torch.float16
    os.environ["M3R_GEMM_BN"] = str(bn)
    try:
        a, w = rnd(M, K, dtype=dtype, seed=1), rnd(N, K, dtype=dtype, seed=2, scale=K ** -0.5)
        bias = rnd(N, seed=3)
        ref = a.float() @ w.float().t() + bias
        assert rel_l2(ops.linear(a, w, bias, out_dtype=torch.float32), ref) < OUT_TOL[torch.float32]
        assert rel_l2(ops.linear(a, w, bias, act="gelu"), F.gelu(ref)) < OUT_TOL[dtype]
    finally:
        os.environ.pop("M3R_GEMM_BN", None)


def test_merged_qkv_kv_gemm_default_heuristic():
    """The stacked [5D, D] first GEMM of a decoder block at one view (768 x 3840 x 768): whatever tile the heuristic picks."""
    dtype = torch.bfloat16
    a, w = rnd(768, 768, dtype=dtype, seed=1), rnd(3840, 768, dtype=dtype, seed=2, scale=768 ** -0.5)
    ref = a.float() @ w.float().t()
    assert rel_l2(ops.linear(a, w, None, out_dtype=torch.float32, w_static=True), ref) < OUT_TOL[torch.float32]


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("G,M,N,K", [(12, 768, 1536, 768), (3, 196, 256, 128), (5, 100, 1536, 768), (12, 1536, 1536, 768)])
def test_grouped_gemm(dtype, G, M, N, K):
    """G problems of one shape, weights taken at a stride inside a larger stacked matrix (like rows [3D,5D) of each block's
    [5D, D] weight), outputs appended at a row offset inside larger per-group buffers (the memory tensors)."""
    a = rnd(G, M, K, dtype=dtype, seed=1)
    wbig = rnd(G, N + 64, K, dtype=dtype, seed=2, scale=K ** -0.5)
    bbig = rnd(G, N + 64, seed=3)
    w, bias = wbig[:, 64:], bbig[:, 64:]
    bufs = [torch.zeros((M + 50, N), dtype=dtype, device="cuda") for _ in range(G)]
    ops.linear_grouped(a, w, bias, [b[37:37 + M] for b in bufs])
    for g in range(G):
        ref = a[g].float() @ w[g].float().t() + bias[g]
        assert rel_l2(bufs[g][37:37 + M], ref) < OUT_TOL[dtype], g
        assert float(bufs[g][:37].abs().sum()) == 0 and float(bufs[g][37 + M:].abs().sum()) == 0     # nothing outside the rows
