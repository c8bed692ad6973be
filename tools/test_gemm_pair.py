"""Correctness + timing of the CTA-pair GEMM (M3R_GEMM_PAIR=1) against the 1-CTA kernel in separate processes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from must3r_b200 import ops
torch.manual_seed(0)
print("pair mode:", os.environ.get("M3R_GEMM_PAIR", "0"))
for dt in (torch.bfloat16, torch.float16):
    for (M, N, K) in [(15360, 3072, 1024), (15360, 1024, 4096), (5000, 768, 768), (15360, 2304, 768)]:
        a = torch.randn(M, K, device="cuda").to(dt); w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(dt)
        bias = torch.randn(N, device="cuda"); res = torch.randn(M, N, device="cuda")
        ref = a.float() @ w.float().t() + bias
        out = ops.linear(a, w, bias, out_dtype=torch.float32)
        e1 = float((out - ref).norm() / ref.norm())
        out = ops.linear(a, w, bias, act="gelu")
        e2 = float((out.float() - torch.nn.functional.gelu(ref)).norm() / torch.nn.functional.gelu(ref).norm())
        x = res.clone(); ops.linear(a, w, bias, residual=x, out=x)
        e3 = float((x - (ref + res)).norm() / (ref + res).norm())
        print(f"{str(dt):15s} {M}x{N}x{K}: fp32-out rel {e1:.2e}  gelu16 rel {e2:.2e}  residual rel {e3:.2e}", flush=True)
