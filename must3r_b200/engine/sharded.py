"""Multi-GPU schedule of the hot path (the reference has no multi-GPU inference at all).

One process per GPU (torch.distributed, NCCL over NVLink; gloo in the CPU tests).  Rank r holds view_counts[r]
views of ONE scene (global view order is rank-major):

  1. encoder: each rank encodes its own V views (independent units, no collective);
  2. memory init: rank 0 runs the reference's 2-view initialisation on its views 0,1 (decoder.py:280-285) and the
     resulting K|V tokens reach every rank through the same gather as the other rounds;
  3. memory update, round s = 0..V-1: every rank whose local view s is not yet in memory runs a shard-local
     update of that ONE view against the replicated memory M_{k-1} (peers of the same round do not see each
     other: "shard-local update"), then ONE all-gather per round moves the packed new post-feedback K|V tokens
     [depth, N, 2D] of all ranks into every rank's pre-allocated memory buffer -> M_k identical everywhere,
     labels in rank order;
  4. render: each rank renders its V views against the final replicated memory (no collective).

On CUDA with the must3r_b200 decoder the gather is FUSED into the producing GEMM: the memory buffers live in
peer-visible device memory (`engine/peer.py`, CUDA IPC) and the epilogue of the post-feedback K|V projection (one grouped
GEMM over all decoder levels) stores each 16-bit tile into every rank's buffer over NVLink while the kernel is still
computing other tiles.  Rounds are ordered by a DEVICE-side flag barrier (m3r_peer_signal / m3r_peer_wait: the producers
publish an epoch into every rank's flag array, consumers spin on their own copy): no host synchronisation and no
collective on the data path, so the launch thread enqueues round k+1 while round k runs.  M3R_FUSED_GATHER=0 selects the
NCCL all-gather path, which is also what gloo/CPU runs use.  Ranks may hold different numbers of views (`view_counts`,
e.g. 100 views ceil-split over 8 GPUs): a rank without a view in a round only waits.

With world_size 1 the schedule degenerates to the reference chain (init 2 views, then 1 view per step).
The oracle for world_size > 1 is composed from single-process decoder calls only (tests/test_sharded_cpu.py).
The function is model-agnostic: decoders exposing ``update_tokens`` (the CUDA MUSt3R) avoid materialising the
concatenated memory; any reference-style decoder works through ``mem'[0][l][:, Nm:]``.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional

import torch
import torch.distributed as dist

from .inference import _sync_host_copies, _to_out

_LIVE_ARENAS = []      # arenas backing memory tuples handed back to callers (return_mem=True)
_ARENA_CACHE = {}      # (nbytes, device) -> PeerArena reused by calls that do not hand their memory back


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _new_tokens(decoder, x, pos, ts, mem):
    """-> (list[depth] of [1, n*N, mem_D] new post-feedback tokens, raw pointmaps [1,n,H,W,C])"""
    if hasattr(decoder, "update_tokens"):
        return decoder.update_tokens(x, pos, ts, mem)
    Nm = 0 if mem is None else mem[0][0].shape[1]
    new_mem, pm = decoder(x, pos, ts, mem, render=False)
    return [m[:, Nm:] for m in new_mem[0]], pm


def _all_gather(packed: torch.Tensor, world: int) -> torch.Tensor:
    """[...]: -> [world, ...] (one collective).  all_gather_into_tensor on NCCL; list form elsewhere."""
    out = packed.new_empty((world,) + tuple(packed.shape))
    if packed.is_cuda:
        dist.all_gather_into_tensor(out, packed.contiguous())
    else:
        dist.all_gather(list(out.unbind(0)), packed.contiguous())
    return out


@torch.no_grad()
def inference_sharded(encoder, decoder, imgs: torch.Tensor, true_shape: torch.Tensor,
                      post_process_function: Optional[Callable] = None, device=None, to_host: bool = False,
                      render_bs: Optional[int] = None, return_mem: bool = False, view_counts: Optional[List[int]] = None):
    """imgs [V,3,H,W], true_shape [V,2]: this rank's views of ONE scene (same image size on every rank).
    `view_counts[r]` = number of views held by rank r (default: the same V everywhere); global view order is rank-major.
    Returns the list of this rank's V rendered results (dicts if post_process_function is given)."""
    rank, world = _world()
    device = device or imgs.device
    V = imgs.shape[0]
    counts = list(view_counts) if view_counts is not None else [V] * world
    assert len(counts) == world and counts[rank] == V, "view_counts must list every rank's number of views"
    assert counts[0] >= 1, "rank 0 holds the first views of the scene (memory initialisation)"
    total = sum(counts)
    hw = None
    if not true_shape.is_cuda:                                 # host copy of (H, W): spares a device sync per decoder call
        assert bool((true_shape == true_shape[:1]).all()), "all views of a rank must share one true_shape"
        hw = tuple(int(v) for v in true_shape[0].tolist())
    imgs, ts_dev = imgs.to(device), true_shape.to(device)
    x, pos = encoder(imgs, ts_dev)                             # [V,N,Denc], [V,N,2]
    N = x.shape[1]

    class _TS:                                                 # true_shape[None, a:b] with the host hint attached
        def __getitem__(self, idx):
            t = ts_dev[idx]
            if hw is not None:
                try:
                    t._m3r_hw = hw
                except Exception:  # noqa: BLE001
                    pass
            return t
    true_shape = _TS()

    mem_vals: Optional[List[torch.Tensor]] = None              # depth x [1, cap, mem_D] pre-allocated
    labels = None
    n_mem_views = 0
    cap = total * N

    def current_mem():
        if n_mem_views == 0:
            return None
        Nm = n_mem_views * N
        return ([m[:, :Nm] for m in mem_vals], labels[:, :Nm], n_mem_views, n_mem_views, Nm)

    def append(gathered: torch.Tensor, flags: List[bool], n_each: int):
        """gathered [world, depth, n_each*N, mem_D]; append the participating ranks' tokens in rank order."""
        nonlocal mem_vals, labels, n_mem_views
        depth, mem_D = gathered.shape[1], gathered.shape[3]
        if mem_vals is None:
            mem_vals = [torch.empty((1, cap, mem_D), dtype=gathered.dtype, device=gathered.device) for _ in range(depth)]
            labels = torch.empty((1, cap), dtype=torch.int64, device=gathered.device)
        sel = [r for r in range(world) if flags[r]]
        if not sel:
            return
        Nm = n_mem_views * N
        cnt = len(sel) * n_each * N
        src = gathered if len(sel) == world else gathered[sel]
        for l in range(depth):
            mem_vals[l][0, Nm:Nm + cnt] = src[:, l].reshape(cnt, mem_D)
        lab = torch.arange(n_mem_views, n_mem_views + len(sel) * n_each, device=labels.device)
        labels[0, Nm:Nm + cnt] = lab.repeat_interleave(N)
        n_mem_views += len(sel) * n_each

    n_init = min(2, counts[0])
    rounds = max(counts)

    def round_flags(s):
        """ranks that store view s in round s (rank 0's first n_init views went in with the initialisation)"""
        return [s < counts[r] and not (r == 0 and s < n_init) for r in range(world)]

    fused = (world > 1 and x.is_cuda and hasattr(decoder, "update_tokens_to_peers")
             and getattr(decoder, "memory_mode", "kv") == "kv"      # the GEMM epilogue that stores to peers is the K|V one
             and os.environ.get("M3R_FUSED_GATHER", "1") != "0")
    arena = None
    if fused:
        # ---- fused GEMM -> all-gather: every rank's memory buffers are peer-mapped; producers store into all of them.
        # Ordering is kept on the GPUs: a flag barrier per round (PeerArena.signal / wait), the host never blocks.
        from .peer import PeerArena
        from ..model.common import stream_ptr
        depth, mem_D, dt = decoder.depth, 2 * decoder.embed_dim, decoder.memory_dtype()
        esz = torch.empty((), dtype=dt).element_size()
        akey = (depth * cap * mem_D * esz, str(x.device))
        arena = None if return_mem else _ARENA_CACHE.get(akey)
        if arena is None:
            arena = PeerArena(akey[0], x.device)          # collective: every rank allocates, exports and maps
            if not return_mem:
                _ARENA_CACHE[akey] = arena
        mem_vals = [arena.local[l * cap * mem_D * esz:(l + 1) * cap * mem_D * esz].view(dt).view(1, cap, mem_D) for l in range(depth)]
        labels = torch.empty((1, cap), dtype=torch.int64, device=x.device)
        sp = stream_ptr(x.device)

        def dests(row):       # peer_ptrs[r][l] for new tokens starting at memory row `row`
            return [[arena.ptrs[r] + (l * cap + row) * mem_D * esz for l in range(depth)] for r in range(world)]

        def commit(participants, n_views_added):
            """The participants' rows of this round have been enqueued: they signal, everyone waits for them (on the GPU)."""
            nonlocal n_mem_views
            Nm = n_mem_views * N
            cnt = n_views_added * N
            labels[0, Nm:Nm + cnt] = torch.arange(n_mem_views, n_mem_views + n_views_added, device=labels.device).repeat_interleave(N)
            n_mem_views += n_views_added
            if rank in participants:
                arena.signal(sp)
            else:
                arena.skip_epoch()
            arena.wait(participants, sp)

        # entry barrier: a cached arena may still be read by a slower rank's render of the previous call
        arena.signal(sp)
        arena.wait(range(world), sp)
        if rank == 0:
            decoder.update_tokens_to_peers(x[None, :n_init], pos[None, :n_init], true_shape[None, :n_init], None, dests(0))
        commit([0], n_init)
        for s in range(rounds):
            flags = round_flags(s)
            part = [r for r in range(world) if flags[r]]
            if not part:
                continue
            if flags[rank]:
                slot = sum(flags[:rank])
                decoder.update_tokens_to_peers(x[None, s:s + 1], pos[None, s:s + 1], true_shape[None, s:s + 1], current_mem(),
                                               dests((n_mem_views + slot) * N))
            commit(part, len(part))

    # ---- 2. init on rank 0 (views 0,1); every rank runs the same gather so the memory is replicated
    if not fused:
        if rank == 0:
            toks, _ = _new_tokens(decoder, x[None, :n_init], pos[None, :n_init], true_shape[None, :n_init], None)
            packed = torch.stack([t[0] for t in toks], 0)                                  # [depth, n_init*N, mem_D]
        else:
            packed = None
        if world > 1:
            shape = [0, 0, 0]
            if rank == 0:
                shape = list(packed.shape)
            st = torch.tensor(shape, dtype=torch.int64, device=device)
            dist.broadcast(st, src=0)
            if rank != 0:
                dt = getattr(decoder, "memory_dtype", None)
                dt = dt() if callable(dt) else (dt or torch.float32)
                packed = torch.empty(tuple(int(v) for v in st.tolist()), dtype=dt, device=device)
            dist.broadcast(packed, src=0)
        append(packed[None] if world == 1 else packed[None].expand(world, *packed.shape),
               [r == 0 for r in range(world)], n_init)

        # ---- 3. update rounds: one view per rank per round, one all-gather per round
        for s in range(rounds):
            flags = round_flags(s)
            if not any(flags):
                continue
            # ranks that sit a round out still take part in the collective (with a dummy payload)
            if flags[rank]:
                toks, _ = _new_tokens(decoder, x[None, s:s + 1], pos[None, s:s + 1], true_shape[None, s:s + 1], current_mem())
                packed = torch.stack([t[0] for t in toks], 0)                              # [depth, N, mem_D]
            else:
                packed = torch.zeros((len(mem_vals), N, mem_vals[0].shape[2]), dtype=mem_vals[0].dtype, device=device)
            gathered = _all_gather(packed, world) if world > 1 else packed[None]
            append(gathered, flags, 1)

    # ---- 4. render this rank's views against the replicated memory
    mem = current_mem()
    outs = []
    bs = render_bs or max(V, 1)
    for lo in range(0, V, bs):
        _, pm = decoder(x[None, lo:lo + bs], pos[None, lo:lo + bs], true_shape[None, lo:lo + bs], mem, render=True)
        pm = pm[0]
        if post_process_function is not None:
            res = post_process_function(pm)
            res = {k: (_to_out(v, "cpu") if to_host else v) for k, v in res.items()}
            outs.extend({k: v[j] for k, v in res.items()} for j in range(pm.shape[0]))
        else:
            pm = _to_out(pm, "cpu") if to_host else pm
            outs.extend(pm[j] for j in range(pm.shape[0]))
    _sync_host_copies()
    if arena is not None and return_mem:
        _LIVE_ARENAS.append(arena)           # the returned memory tensors are views of the arena
    return (mem, outs) if return_mem else outs
