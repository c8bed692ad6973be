"""The dominant kernels of the hot path at their C3 shapes, each launched a few times, for `ncu --set full` captures
(profiles/r02_ncu_*): `python tools/ncu_targets.py emit|merged|grouped|fc1|ca1|carender|sa`."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from must3r_b200 import ops  # noqa: E402

dt = torch.bfloat16
which = sys.argv[1]
reps = 3
torch.manual_seed(0)
if which == "emit":          # proj / cproj form: 768x768x768 + bias + fp32 residual, LayerNorm emitted (72 CTAs)
    a = torch.randn(768, 768, device="cuda").to(dt); w = (torch.randn(768, 768, device="cuda") / 28).to(dt)
    b = torch.randn(768, device="cuda"); r = torch.randn(768, 768, device="cuda"); h = torch.empty(768, 768, device="cuda", dtype=dt)
    for _ in range(reps):
        ops.linear(a, w, b, residual=r, out_dtype=torch.float32, norm_out=h, w_static=True)
elif which == "emitk":       # fc2 form: K = 3072, split-K CTA pairs (144 CTAs)
    a = torch.randn(768, 3072, device="cuda").to(dt); w = (torch.randn(768, 3072, device="cuda") / 55).to(dt)
    b = torch.randn(768, device="cuda"); r = torch.randn(768, 768, device="cuda"); h = torch.empty(768, 768, device="cuda", dtype=dt)
    for _ in range(reps):
        ops.linear(a, w, b, residual=r, out_dtype=torch.float32, norm_out=h, w_static=True)
elif which == "merged":      # first GEMM of a decoder block at one view: [q|k|v|Kc|Vc] 768 x 3840 x 768 (BN = 160, 144 CTAs)
    a = torch.randn(768, 768, device="cuda").to(dt); w = (torch.randn(3840, 768, device="cuda") / 28).to(dt)
    for _ in range(reps):
        ops.linear(a, w, None, w_static=True)
elif which == "grouped":     # memory append: 12 levels x (768 x 1536 x 768) in one launch
    a = torch.randn(12, 768, 768, device="cuda").to(dt); w = (torch.randn(12, 1536, 768, device="cuda") / 28).to(dt)
    outs = [torch.empty(768, 1536, device="cuda", dtype=dt) for _ in range(12)]
    for _ in range(reps):
        ops.linear_grouped(a, w, None, outs)
elif which == "fc1":
    a = torch.randn(768, 768, device="cuda").to(dt); w = (torch.randn(3072, 768, device="cuda") / 28).to(dt)
    for _ in range(reps):
        ops.linear(a, w, None, act="gelu", w_static=True)
elif which in ("ca1", "carender", "sa", "ca100"):
    B, H, Nq, Nk, grp = {"ca1": (1, 12, 768, 7680, 1), "ca100": (1, 12, 768, 76800, 1), "carender": (20, 12, 768, 15360, 20), "sa": (20, 16, 768, 768, 1)}[which]
    D = H * 64
    q = torch.randn(B * Nq, D, device="cuda").to(dt)
    nb = B // grp if grp > 1 else B
    kv = torch.randn(nb * Nk, 2 * D, device="cuda").to(dt)
    for _ in range(reps):
        ops.attention(q, kv[:, :D], kv[:, D:], B=B, H=H, Nq=Nq, Nk0=Nk, kv_group=grp if grp > 1 else 1)
torch.cuda.synchronize()
