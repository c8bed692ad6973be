"""Context-parallel memory cross-attention: ONE scene / one stream on several GPUs (the reference has no
multi-GPU inference).

The decoder chain of a stream does not shard by views (frame t needs the memory of frame t-1), but its dominant cost does:
with ~360 memory views every one-view step is > 80 % memory cross-attention.  Here the memory TOKENS are sharded: every rank
runs the same decoder call on the same frame (the small GEMMs and the self-attention are replicated), computes the
attention state of the queries over ITS shard of the K|V rows, the states are exchanged through peer memory and merged
exactly (softmax over the union, like the in-kernel merge of key splits), and only one rank - round-robin over the
update calls - appends the frame's new rows to its shard.  Results are identical on all ranks and equal to the
single-GPU chain up to the summation order of the softmax.

    dec_cp = ContextParallelDecoder(decoder)              # collective: allocates the peer-mapped staging buffers
    engine.inference_video_multi_ar(encoder, dec_cp, ...)  # every rank passes the same frames

The wrapper has the decoder call signature of the reference (must3r/model/decoder.py:158,267); the memory tuple it returns
holds this rank's shard (values, labels) and the scene-global counters, so the engine's label-based edits
(engine/inference.py: _remove_from_mem / _update_in_mem / _restore_label_in_mem) work unchanged on every rank.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from .. import _lib
from ..model import common as cm
from .peer import PeerArena


class ShardedMemoryDecoder:
    """Sharding policy shared by the CUDA wrapper below and by the CPU emulation the tests drive the engine with
    (tests/test_context_parallel_cpu.py): which rank stores a call's new tokens, what an empty shard looks like, how a
    large render is chunked.  Subclasses implement `_first(xs, ps, ts)` (the scene's first call, replicated) and
    `_run(xs, ps, ts, current_mem, render, owner)` (a call on the sharded memory)."""

    rank, world = 0, 1
    max_rows = 1 << 62

    def __init__(self):
        self._calls = 0

    def _next_owner(self, stores: bool) -> bool:
        owner = (self._calls % self.world) == self.rank                # round-robin over the update calls (same on every rank)
        if stores:
            self._calls += 1
        return owner

    @torch.no_grad()
    def __call__(self, x, pos, true_shape, current_mem=None, render=False, return_feats=False):
        assert not return_feats
        as_list = isinstance(x, (list, tuple))
        xs, ps, ts = (list(x), list(pos), list(true_shape)) if as_list else ([x], [pos], [true_shape])
        rows = sum(int(v.shape[0] * v.shape[1] * v.shape[2]) for v in xs)
        if current_mem is None:
            # first call of the scene: no memory yet, replicated on every rank; rank 0 keeps the tokens
            mem, pms = self._first(xs, ps, ts)
            if self.rank != 0:
                vals = [v[:, :0] for v in mem[0]]
                lab = mem[1][:, :0].contiguous()
                lab._m3r_labels_host = torch.zeros((0,), dtype=torch.int64)
                mem = (vals, lab, mem[2], mem[3], 0)
        elif render and rows > self.max_rows and len(xs) == 1 and xs[0].shape[0] == 1:
            # a render pass larger than the staging buffers: views are independent given the memory, so go in chunks
            per_view = int(xs[0].shape[2])
            step = max(1, self.max_rows // per_view)
            parts = []
            for lo in range(0, xs[0].shape[1], step):
                sl = slice(lo, lo + step)
                _, pm = self._run([xs[0][:, sl]], [ps[0][:, sl]], [ts[0][:, sl]], current_mem, True, False)
                parts.append(pm[0])
            mem, pms = tuple(current_mem), [torch.cat(parts, 1)]
        else:
            if rows > self.max_rows:
                raise RuntimeError(f"context-parallel call with {rows} token rows; the staging buffers hold {self.max_rows} "
                                   "(ContextParallelDecoder(max_rows_per_call=...))")
            mem, pms = self._run(xs, ps, ts, current_mem, render, self._next_owner(not render))
        return (mem, pms) if as_list else (mem, pms[0])

    forward = __call__


class ContextParallelDecoder(ShardedMemoryDecoder):
    def __init__(self, decoder, max_rows_per_call: int = 2 * 768):
        super().__init__()
        assert dist.is_initialized() and decoder.memory_mode == "kv"
        self.dec = decoder
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        assert self.world <= 8
        dev = decoder.norm_dec.weight.device
        w = decoder._packed(cm.get_precision())
        self.max_rows = max_rows_per_call
        self.slot_bytes = int(_lib.lib().m3r_decoder_cp_slot_bytes(C.byref(w), max_rows_per_call))
        self.arena = PeerArena(2 * self.world * self.slot_bytes, dev)             # collective
        self._stage = (C.c_void_p * self.world)(*self.arena.ptrs)
        self._flag_slots = (C.c_void_p * self.world)(*[self.arena.ptrs[r] + self.arena.flag_off + 4 * self.rank for r in range(self.world)])

    # attributes the engine / callers read
    def __getattr__(self, name):
        return getattr(self.dec, name)

    def reserve_memory(self, n_tokens: int = 0, growth: float = 0.0):
        # a shard receives ~1/world of the scene's tokens
        self.dec.reserve_memory(-(-int(n_tokens) // self.world) if n_tokens else 0, growth)

    def _first(self, xs, ps, ts):
        return self.dec.forward_list(xs, ps, ts, None, False)

    def _run(self, xs, ps, ts, current_mem, render, owner):
        cp = dict(world=self.world, rank=self.rank, owner=owner, stage_ptrs=self._stage, slot_bytes=self.slot_bytes,
                  flag_slots=self._flag_slots, flags_local=C.c_void_p(self.arena.ptr + self.arena.flag_off), epoch0=self.arena.epoch)
        self.arena.epoch += self.dec.depth                              # every rank consumes the same epochs
        return self.dec.forward_list(xs, ps, ts, current_mem, render, _cp=cp)

    def gather_memory(self, mem):
        """Debug / test helper: the scene's whole memory (rows of all shards, sorted by label) on every rank."""
        vals, labels = mem[0], mem[1]
        counts = [torch.zeros(1, dtype=torch.int64, device=labels.device) for _ in range(self.world)]
        dist.all_gather(counts, torch.tensor([labels.shape[1]], dtype=torch.int64, device=labels.device))
        counts = [int(c.item()) for c in counts]
        mx = max(max(counts), 1)
        pad_l = torch.full((1, mx), -1, dtype=labels.dtype, device=labels.device)
        pad_l[:, :labels.shape[1]] = labels
        all_l = [torch.empty_like(pad_l) for _ in range(self.world)]
        dist.all_gather(all_l, pad_l)
        lab = torch.cat([a[:, :c] for a, c in zip(all_l, counts)], 1)
        order = torch.argsort(lab[0], stable=True)
        out_vals = []
        for v in vals:
            pad = torch.zeros((1, mx, v.shape[2]), dtype=v.dtype, device=v.device)
            pad[:, :v.shape[1]] = v
            allv = [torch.empty_like(pad) for _ in range(self.world)]
            dist.all_gather(allv, pad)
            cat = torch.cat([a[:, :c] for a, c in zip(allv, counts)], 1)
            out_vals.append(cat[:, order])
        return out_vals, lab[:, order]
