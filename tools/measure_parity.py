"""Print rel-L2 of the CUDA path vs the reference digests / fixtures for both operand formats (run on the GPU box)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import load_golden, digest, rel  # noqa: E402
from must3r_b200 import synthetic as syn  # noqa: E402
from must3r_b200.model import Dust3rEncoder, MUSt3R, set_precision  # noqa: E402

g = load_golden("full_model_digest.npz")
for tag, H, W, size in [("224", 224, 224, 224), ("512", 384, 512, 512)]:
    enc = Dust3rEncoder(img_size=(size, size)); dec = MUSt3R(img_size=(size, size), feedback_type="single_mlp", memory_mode="kv", landscape_only=False)
    enc.load_state_dict(syn.encoder_state_dict(0)); dec.load_state_dict(syn.decoder_state_dict(0))
    enc, dec = enc.cuda().eval(), dec.cuda().eval()
    imgs, ts = syn.synthetic_views(3, H, W, seed=2); imgs, ts = imgs.cuda(), ts.cuda()
    for dt in (torch.float16, torch.bfloat16):
        set_precision(dt)
        x, pos = enc(imgs, ts)
        r = {"enc_x": rel(digest(x), g[f"{tag}.enc_x"])}
        mem, pm = dec(x[None, :2], pos[None, :2], ts[None, :2], None); r["pm_init"] = rel(digest(pm), g[f"{tag}.pm_init"])
        mem, pm = dec(x[None, 2:3], pos[None, 2:3], ts[None, 2:3], mem); r["pm_upd"] = rel(digest(pm), g[f"{tag}.pm_upd"])
        r["mem0"] = rel(digest(mem[0][0]), g[f"{tag}.mem0"]); r["mem11"] = rel(digest(mem[0][11]), g[f"{tag}.mem11"])
        _, pm = dec(x[None], pos[None], ts[None], mem, render=True); r["pm_render"] = rel(digest(pm), g[f"{tag}.pm_render"])
        print(tag, dt, {k: f"{v:.2e}" for k, v in r.items()}, flush=True)
