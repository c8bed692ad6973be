"""Run under torchrun on N GPUs: the sharded schedule over NCCL (engine.sharded.inference_sharded) must reproduce the
schedule composed from single-process decoder calls (SURVEY.md §8e oracle), with bitwise-identical memory on all ranks.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py"""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from must3r_b200 import synthetic as syn  # noqa: E402
from must3r_b200.engine import sharded  # noqa: E402
from must3r_b200.model import Dust3rEncoder, MUSt3R, set_precision  # noqa: E402

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
dist.init_process_group("nccl", device_id=dev)
set_precision(torch.bfloat16)
V, H, W = 3, 224, 224
enc = Dust3rEncoder(img_size=(224, 224), depth=6); dec = MUSt3R(img_size=(224, 224), depth=4, feedback_type="single_mlp", memory_mode="kv", landscape_only=False)
enc.load_state_dict(syn.encoder_state_dict(3, depth=6)); dec.load_state_dict(syn.decoder_state_dict(3, depth=4))
enc, dec = enc.to(dev).eval(), dec.to(dev).eval()
views = [syn.synthetic_views(V, H, W, seed=400 + r) for r in range(world)]
imgs, ts = views[rank][0].to(dev), views[rank][1].to(dev)
mem, outs = sharded.inference_sharded(enc, dec, imgs, ts, device=dev, return_mem=True)
outs = torch.stack(outs)
# composed expectation, computed locally from every rank's views
feats = []
for r in range(world):
    x, pos = enc(views[r][0].to(dev), views[r][1].to(dev)); feats.append((x, pos, views[r][1].to(dev)))
x0, p0, t0 = feats[0]
m, _ = dec(x0[None, :2], p0[None, :2], t0[None, :2], None); m = list(m)
for s in range(V):
    parts = []
    for r in range(world):
        if r == 0 and s < 2: continue
        x, pos, t = feats[r]
        toks, _ = dec.update_tokens(x[None, s:s + 1], pos[None, s:s + 1], t[None, s:s + 1], tuple(m))
        parts.append(toks)
    if not parts: continue
    N = parts[0][0].shape[1]
    m[0] = [torch.cat([m[0][l]] + [p[l] for p in parts], 1) for l in range(len(m[0]))]
    lab = torch.arange(m[2], m[2] + len(parts), device=dev).repeat_interleave(N)[None]
    m[1] = torch.cat([m[1], lab], 1); m[2] = m[3] = m[2] + len(parts); m[4] = m[1].shape[1]
x, pos, t = feats[rank]
_, pm = dec(x[None], pos[None], t[None], tuple(m), render=True)
err = float((outs - pm[0]).norm() / pm[0].norm())
mem_equal = all(torch.equal(mem[0][l], m[0][l]) for l in range(len(m[0]))) and torch.equal(mem[1], m[1])
chk = torch.stack([mm.float().sum() for mm in mem[0]])
allc = [torch.empty_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(torch.equal(allc[0], c) for c in allc)
print(f"rank {rank}/{world}: render rel err vs composed schedule {err:.3e}; memory == composed: {mem_equal}; memory checksum identical on all ranks: {same}; Nmem {mem[1].shape[1]}", flush=True)
ok = err < 1e-5 and mem_equal and same
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
