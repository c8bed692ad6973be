"""``curope`` drop-in: routes the UNMODIFIED reference's RoPE through ``m3r_rope_2d`` (libm3r_b200.so).

The reference picks its RoPE at import time: any importable module named ``curope`` that exposes
``rope_2d(tokens[B,N,H,D], positions[B,N,2] int64, base, fwd) -> None`` (in place, strided tokens allowed) is used
(dust3r/croco/models/pos_embed.py:104-106, dust3r/croco/models/curope/curope2d.py:6-9, patched in at
dust3r/dust3r/utils/path_to_croco.py:36-42).  Install it with::

    import sys, must3r_b200.compat.curope as shim
    sys.modules["curope"] = shim          # before the first `import must3r` / `import dust3r`

or copy this file next to the checkout as ``curope.py``.  Same operator contract as curope.cpp:49-69: float base and
fwd (F0, negative for the backward rotation), RuntimeError for bad shapes, fp32 / fp16 / bf16 tokens; unlike the
reference launcher (kernels.cu:102) the kernel runs on the caller's current CUDA stream.  CUDA tensors only - there is
no CPU path here (the reference's curope.cpp:17-47 CPU loop has no counterpart; use the PyTorch fallback on CPU).
"""
import torch

from .. import ops


def rope_2d(tokens, positions, base, fwd):
    ops.rope_2d(tokens, positions, float(base), float(fwd))


class cuRoPE2D_func(torch.autograd.Function):
    """dust3r/croco/models/curope/curope2d.py:12-29 (forward rotates by +F0 in place, backward by -F0)."""

    @staticmethod
    def forward(ctx, tokens, positions, base, F0=1):
        ctx.save_for_backward(positions)
        ctx.saved_base, ctx.saved_F0 = base, F0
        rope_2d(tokens, positions, base, F0)
        ctx.mark_dirty(tokens)
        return tokens

    @staticmethod
    def backward(ctx, grad_res):
        positions, base, F0 = ctx.saved_tensors[0], ctx.saved_base, ctx.saved_F0
        rope_2d(grad_res, positions, base, -F0)
        ctx.mark_dirty(grad_res)
        return grad_res, None, None, None


class cuRoPE2D(torch.nn.Module):
    """dust3r/croco/models/curope/curope2d.py:32-39: tokens [B,H,N,D] rotated in place through a [B,N,H,D] view."""

    def __init__(self, freq=100.0, F0=1.0):
        super().__init__()
        self.base, self.F0 = freq, F0

    def forward(self, tokens, positions):
        cuRoPE2D_func.apply(tokens.transpose(1, 2), positions, self.base, self.F0)
        return tokens
