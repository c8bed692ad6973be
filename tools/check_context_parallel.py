"""Run under torchrun on N GPUs: context-parallel memory cross-attention (engine/context_parallel.py) must reproduce the
single-GPU chain: (1) single decoder calls (update with 1 and 2 views, render) on a sharded memory vs the plain decoder on
the whole memory; (2) the streaming schedule (keyframes, rolling window, evictions, one refinement pass) through
engine.inference_video_multi_ar on every rank vs the same schedule with the plain decoder; the union of the shards must be
the single-GPU memory.  Differences are only the summation order of the softmax across shards (gate 2e-3 rel-L2, fp16).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/check_context_parallel.py"""
import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from must3r_b200 import engine, synthetic as syn  # noqa: E402
from must3r_b200.engine.context_parallel import ContextParallelDecoder  # noqa: E402
from must3r_b200.model import Dust3rEncoder, MUSt3R, set_precision, ActivationType  # noqa: E402

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
dist.init_process_group("nccl", device_id=dev)
set_precision(torch.float16)
ok_all = True


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def case(tag, size, H, W, de, dd, F, window):
    global ok_all
    enc = Dust3rEncoder(img_size=(size, size), depth=de)
    dec = MUSt3R(img_size=(size, size), depth=dd, feedback_type="single_mlp", memory_mode="kv", landscape_only=False)
    enc.load_state_dict(syn.encoder_state_dict(5, depth=de)); dec.load_state_dict(syn.decoder_state_dict(5, depth=dd))
    enc, dec = enc.to(dev).eval(), dec.to(dev).eval()
    N = (H // 16) * (W // 16)
    cpd = ContextParallelDecoder(dec, max_rows_per_call=3 * N)      # the 5-view render below goes in two chunks
    imgs, ts = syn.synthetic_views(F, H, W, seed=77)
    imgs, tsd = imgs.to(dev), ts.to(dev)
    x, pos = enc(imgs, tsd)
    # ---- (1) single calls
    m_ref, _ = dec(x[None, :2], pos[None, :2], tsd[None, :2], None)
    m_cp, _ = cpd(x[None, :2], pos[None, :2], tsd[None, :2], None)
    errs = []
    for i in range(2, 6):
        m_ref, pr = dec(x[None, i:i + 1], pos[None, i:i + 1], tsd[None, i:i + 1], m_ref)
        m_cp, pc = cpd(x[None, i:i + 1], pos[None, i:i + 1], tsd[None, i:i + 1], m_cp)
        errs.append(rel(pc, pr))
    m_ref, pr = dec(x[None, 6:8], pos[None, 6:8], tsd[None, 6:8], m_ref)              # two views in one call (rank 0 holds their mutual keys)
    m_cp, pc = cpd(x[None, 6:8], pos[None, 6:8], tsd[None, 6:8], m_cp)
    errs.append(rel(pc, pr))
    _, pr = dec(x[None, :5], pos[None, :5], tsd[None, :5], m_ref, render=True)
    _, pc = cpd(x[None, :5], pos[None, :5], tsd[None, :5], m_cp, render=True)
    errs.append(rel(pc, pr))
    gv, gl = cpd.gather_memory(m_cp)
    mem_ok = torch.equal(gl, m_ref[1]) and int(m_cp[2]) == int(m_ref[2])
    e_mem = max(rel(a.float(), b.float()) for a, b in zip(gv, m_ref[0]))
    shard = m_cp[1].shape[1] // N
    ok = max(errs) < 2e-3 and mem_ok and e_mem < 2e-3
    print(f"[{tag}] rank {rank}/{world}: single calls max rel err {max(errs):.2e}; labels of the union == single GPU: {mem_ok}; memory rel err {e_mem:.2e}; this shard holds {shard} of {gl.shape[1] // N} views", flush=True)
    ok_all = ok_all and ok
    # ---- (2) streaming schedule through the engine
    views, tss = list(imgs.unbind(0)), list(ts.unbind(0))
    pp = lambda pm: engine.postprocess(pm, ActivationType.NORM_EXP)  # noqa: E731
    kw = dict(encoder_precomputed_features=(list(x.unbind(0)), list(pos.unbind(0))), post_process_function=pp, device=dev, return_mem=True,
              local_context_size=window, num_refinements_iterations=1)
    mem_r, out_r = engine.inference_video_multi_ar(enc, dec, views, tss, [2] + [1] * (F - 2), **kw)
    mem_c, out_c = engine.inference_video_multi_ar(enc, cpd, views, tss, [2] + [1] * (F - 2), **kw)
    e_out = max(rel(a[k], b[k]) for a, b in zip(out_c, out_r) for k in a)
    gv, gl = cpd.gather_memory(mem_c)
    mem_ok = torch.equal(gl, mem_r[1]) and int(mem_c[2]) == int(mem_r[2])
    e_mem = max(rel(a.float(), b.float()) for a, b in zip(gv, mem_r[0])) if mem_ok else float("nan")
    ok = e_out < 3e-3 and mem_ok and e_mem < 3e-3
    print(f"[{tag}] rank {rank}/{world}: stream of {F} frames (window {window}, 1 refinement pass): results max rel err {e_out:.2e}; union of shards == single-GPU memory labels: {mem_ok}; memory rel err {e_mem:.2e}; shard {mem_c[1].shape[1] // N} of {gl.shape[1] // N} views", flush=True)
    ok_all = ok_all and ok
    cpd.arena.close()


case("224 small model", 224, 224, 224, 4, 4, 14, 4)
case("512x384 full model", 512, 384, 512, 24, 12, 12, 5)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
