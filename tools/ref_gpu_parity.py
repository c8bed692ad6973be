"""Parity of the CUDA path against the UNMODIFIED reference running on the same GPU (fp32, TF32 off, its RoPE served
by must3r_b200.compat.curope), on the benchmarked schedules, through both engines' `inference_multi_ar`:

    python tools/ref_gpu_parity.py [c2|c3|c4s] ...

c2 = 10 views 224x224, c3 = 20 views 512x384, c4s = 40 views 512x384 (a shortened C4).  Prints rel-L2 of the rendered
raw head output, pts3d, pts3d_local and conf for fp16 and bf16 operands, and the reference's own GPU time.
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline.ref_loader import load_reference  # noqa: E402
from must3r_b200 import engine, synthetic as syn  # noqa: E402
from must3r_b200.model import Dust3rEncoder, MUSt3R, set_precision, ActivationType  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
ref = load_reference(curope_shim=True)
dev = torch.device("cuda", 0)
CFG = {"c2": (10, 224, 224, 224), "c3": (20, 384, 512, 512), "c4s": (40, 384, 512, 512)}


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


@torch.no_grad()
def main():
    for tag in (sys.argv[1:] or ["c2", "c3"]):
        V, H, W, size = CFG[tag]
        imgs, ts = syn.synthetic_views(V, H, W, seed=2)
        views, tss = list(imgs.to(dev).unbind(0)), list(ts.unbind(0))
        ids = [torch.tensor(i) for i in range(V)]
        sched = [2] + [1] * (V - 2)
        renc = ref.Dust3rEncoder(img_size=(size, size)).eval()
        rdec = ref.MUSt3R(img_size=(size, size), feedback_type="single_mlp", memory_mode="kv", landscape_only=False).eval()
        renc.load_state_dict(syn.encoder_state_dict(0)); rdec.load_state_dict(syn.decoder_state_dict(0))
        renc, rdec = renc.to(dev), rdec.to(dev)
        raw = lambda pm: {"raw": pm}  # noqa: E731
        for it in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            _, rpm = ref.engine.inference_multi_ar(renc, rdec, views, ids, [t.to(dev) for t in tss], sched, max_bs=None,
                                                   post_process_function=raw, device=dev)
            torch.cuda.synchronize(); t_ref = time.perf_counter() - t0
        r_raw = torch.stack([d["raw"] for d in rpm]).float()
        r_post = ref.engine.postprocess(r_raw, ref.model.ActivationType.NORM_EXP)
        print(f"[{tag}] reference on GPU (fp32, eager, curope shim -> m3r_rope_2d): {t_ref*1e3:.1f} ms/job = {V/t_ref:.1f} views/s", flush=True)
        del renc, rdec
        enc = Dust3rEncoder(img_size=(size, size)); dec = MUSt3R(img_size=(size, size), feedback_type="single_mlp", memory_mode="kv", landscape_only=False)
        enc.load_state_dict(syn.encoder_state_dict(0)); dec.load_state_dict(syn.decoder_state_dict(0))
        enc, dec = enc.to(dev).eval(), dec.to(dev).eval()
        for dt in (torch.float16, torch.bfloat16):
            set_precision(dt)
            for it in range(2):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                x, pos = engine.encoder_multi_ar(enc, views, ts, device=dev)
                _, pm = engine.inference_multi_ar(enc, dec, views, ids, tss, sched, encoder_precomputed_features=(x, pos),
                                                  post_process_function=raw, device=dev)
                torch.cuda.synchronize(); t_b = time.perf_counter() - t0
            o_raw = torch.stack([d["raw"] for d in pm]).float()
            o_post = engine.postprocess(o_raw, ActivationType.NORM_EXP)
            print(f"[{tag}] {str(dt):15s} {t_b*1e3:8.1f} ms/job  rel-L2 raw {rel(o_raw, r_raw):.2e}  pts3d {rel(o_post['pts3d'], r_post['pts3d']):.2e}  "
                  f"pts3d_local {rel(o_post['pts3d_local'], r_post['pts3d_local']):.2e}  conf {rel(o_post['conf'], r_post['conf']):.2e}  "
                  f"last view raw {rel(o_raw[-1], r_raw[-1]):.2e}", flush=True)
        del enc, dec
        torch.cuda.empty_cache()


main()
