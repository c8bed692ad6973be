"""``curope`` drop-in: routes the UNMODIFIED reference's RoPE through ``m3r_rope_2d`` (libm3r_b200.so).

The reference picks its RoPE at import time: any importable module named ``curope`` that exposes
``rope_2d(tokens[B,N,H,D], positions[B,N,2] int64, base, fwd) -> None`` (in place, strided tokens allowed) is used
(dust3r/croco/models/pos_embed.py:104-106, dust3r/croco/models/curope/curope2d.py:6-9, patched in at
dust3r/dust3r/utils/path_to_croco.py:36-42).  Install it with::

    import sys, must3r_b200.compat.curope as shim
    sys.modules["curope"] = shim          # before the first `import must3r` / `import dust3r`

or copy this file next to the checkout as ``curope.py``.  Same operator contract as curope.cpp:49-69: float base and
fwd (F0; a negative value rotates backwards), RuntimeError for bad shapes, fp32 / fp16 / bf16 tokens; unlike the
reference launcher (kernels.cu:102) the kernel runs on the caller's current CUDA stream.  CUDA tensors only - there is
no CPU path here (the reference's curope.cpp:17-47 CPU loop has no counterpart; use the PyTorch fallback on CPU).
"""
import torch

from .. import ops


def rope_2d(tokens, positions, base, fwd):
    ops.rope_2d(tokens, positions, float(base), float(fwd))


class cuRoPE2D(torch.nn.Module):
    """Drop-in for the module the reference builds through `get_pos_embed` (must3r/model/blocks/pos_embed.py:7-22): rotates
    q / k of shape [B, heads, N, D] IN PLACE through a [B, N, heads, D] view and hands the same tensor back
    (dust3r/croco/models/curope/curope2d.py:32-39).  Inference only: the reference wraps the call in an autograd Function whose
    backward is the rotation by -F0; this path has no training loop, so gradients are refused loudly instead of being wrong."""

    def __init__(self, freq=100.0, F0=1.0):
        super().__init__()
        self.base, self.F0 = freq, F0

    def forward(self, tokens, positions):
        if tokens.requires_grad and torch.is_grad_enabled():
            raise RuntimeError("must3r_b200.compat.curope is inference-only (run under torch.no_grad())")
        rope_2d(tokens.transpose(1, 2), positions, self.base, self.F0)
        return tokens
