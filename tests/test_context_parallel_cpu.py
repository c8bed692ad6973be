"""Engine-side bookkeeping of the context-parallel (memory-sharded) stream under gloo, world size 2, on CPU.

The CUDA wrapper (engine/context_parallel.py ContextParallelDecoder) needs GPUs; its sharding POLICY (owner selection, empty
shards, scene-global counters) lives in the base class ShardedMemoryDecoder.  Here a CPU emulation built on the oracle decoder
implements the two hooks - a call on the sharded memory gathers all shards, sorts the rows by label and runs the oracle on the
whole memory - and the engine's streaming schedule (keyframes, rolling window, evictions, a refinement pass with keyframe
refresh across shards) must give the single-process results, with the union of the shards equal to the single-process memory."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F, H, W = 11, 32, 48


def _setup():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import tiny_oracle
    from must3r_b200 import synthetic as syn
    enc, dec = tiny_oracle(7)
    imgs, ts = syn.synthetic_views(F, H, W, seed=41)
    return enc, dec, imgs, ts


def _gather_rows(t, dim=1):
    """all-gather tensors whose size differs along `dim`"""
    world = dist.get_world_size()
    n = torch.tensor([t.shape[dim]], dtype=torch.int64)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n)
    ns = [int(v) for v in ns]
    mx = max(max(ns), 1)
    shape = list(t.shape); shape[dim] = mx
    pad = torch.zeros(shape, dtype=t.dtype)
    pad.narrow(dim, 0, t.shape[dim]).copy_(t)
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o.narrow(dim, 0, c) for o, c in zip(out, ns)], dim)


def _make_emulation(dec):
    from must3r_b200.engine.context_parallel import ShardedMemoryDecoder

    class OracleSharded(ShardedMemoryDecoder):
        def __init__(self):
            super().__init__()
            self.rank, self.world = dist.get_rank(), dist.get_world_size()

        def _first(self, xs, ps, ts):
            return dec(xs, ps, ts, None)

        def _run(self, xs, ps, ts, current_mem, render, owner):
            labels = _gather_rows(current_mem[1])
            order = torch.argsort(labels[0], stable=True)
            full = ([_gather_rows(v)[:, order] for v in current_mem[0]], labels[:, order], current_mem[2], current_mem[3], labels.shape[1])
            new_mem, pms = dec(xs, ps, ts, full, render=render)
            if render:
                return tuple(current_mem), pms
            n_old = full[1].shape[1]
            tot = new_mem[2]
            if owner:
                vals = [torch.cat([a, b[:, n_old:]], 1) for a, b in zip(current_mem[0], new_mem[0])]
                lab = torch.cat([current_mem[1], new_mem[1][:, n_old:]], 1)
                return (vals, lab, tot, tot, lab.shape[1]), pms
            return (list(current_mem[0]), current_mem[1], tot, tot, current_mem[1].shape[1]), pms
    return OracleSharded()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    enc, dec, imgs, ts = _setup()
    from must3r_b200 import engine
    sharded = _make_emulation(dec)
    mem, out = engine.inference_video_multi_ar(enc, sharded, list(imgs.unbind(0)), list(ts.unbind(0)), [2] + [1] * (F - 2),
                                               post_process_function=lambda p: {"raw": p}, device="cpu", return_mem=True,
                                               local_context_size=3, num_refinements_iterations=1)
    labels = _gather_rows(mem[1])
    order = torch.argsort(labels[0], stable=True)
    vals = [_gather_rows(v)[:, order] for v in mem[0]]
    q.put((rank, torch.stack([o["raw"] for o in out]).clone(), [v.clone() for v in vals], labels[:, order].clone(), int(mem[2]), int(mem[1].shape[1])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_stream_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        port = so.getsockname()[1]
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        r, out, vals, labels, nimgs, n_local = q.get(timeout=240)
        got[r] = (out, vals, labels, nimgs, n_local)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    enc, dec, imgs, ts = _setup()
    from must3r_b200 import engine
    mem, out = engine.inference_video_multi_ar(enc, dec, list(imgs.unbind(0)), list(ts.unbind(0)), [2] + [1] * (F - 2),
                                               post_process_function=lambda p: {"raw": p}, device="cpu", return_mem=True,
                                               local_context_size=3, num_refinements_iterations=1)
    want = torch.stack([o["raw"] for o in out])
    for r in range(world):
        o, vals, labels, nimgs, n_local = got[r]
        assert torch.allclose(o, want, atol=2e-5), float((o - want).abs().max())
        assert torch.equal(labels, mem[1]) and nimgs == int(mem[2])          # the union of the shards is the single-process memory
        for a, b in zip(vals, mem[0]):
            assert torch.allclose(a, b, atol=2e-5)
    assert got[0][4] + got[1][4] == mem[1].shape[1] and min(got[0][4], got[1][4]) > 0      # both ranks hold a share
