"""Saving / loading a scene memory in the reference's pickle layout (must3r/slam/model.py:431-440:
`pkl.dump((memory, keyframe_pointmaps, overlap_tree), f)` / the matching `pkl.load`), for memories produced by must3r_b200.

A memory tuple coming out of the CUDA decoder may be a set of prefix views into growable buffers (model/decoder.py MemArena)
carrying private attributes; `plain_memory` turns it into what the reference's own code would have produced - contiguous
tensors that own exactly their rows, no attributes except the host label shadow (a CPU int64 tensor, which pickles) - so the
file can be read back by either implementation.  16-bit K|V rows are kept as they are (the reference's amp memories are
bf16 / fp16 too); pass `dtype=torch.float32` to widen them for a reference running without autocast."""
from __future__ import annotations

import pickle as pkl
from typing import Any, Optional

import torch


def plain_memory(mem, device: Optional[Any] = None, dtype: Optional[torch.dtype] = None):
    """(values, labels, nimgs, protected_imgs, protected_tokens) -> the same tuple made of plain, compact tensors."""
    vals, labels = mem[0], mem[1]
    out_vals = []
    for v in vals:
        t = v.detach()
        if dtype is not None:
            t = t.to(dtype)
        t = t.to(device) if device is not None else t
        out_vals.append(t.contiguous().clone() if t.data_ptr() == v.data_ptr() else t.contiguous())
    lab = labels.detach().to(device) if device is not None else labels.detach()
    lab = lab.contiguous().clone() if lab.data_ptr() == labels.data_ptr() else lab.contiguous()
    shadow = getattr(labels, "_m3r_labels_host", None)
    if shadow is not None and shadow.shape[0] == lab.shape[1]:
        lab._m3r_labels_host = shadow.clone()
    return (out_vals, lab) + tuple(int(v) for v in mem[2:])


def save_memory(path: str, mem, keyframe_pointmaps=None, overlap_tree=None, dtype: Optional[torch.dtype] = None) -> None:
    """must3r/slam/model.py:431-433 (`write_memory`)"""
    with open(path, "wb") as f:
        pkl.dump((plain_memory(mem, device="cpu", dtype=dtype), keyframe_pointmaps, overlap_tree), f)


def load_memory(path: str, device=None, dtype: Optional[torch.dtype] = None):
    """must3r/slam/model.py:435-440 (`load_memory`): -> (memory, keyframe_pointmaps, overlap_tree)"""
    with open(path, "rb") as f:
        mem, data, tree = pkl.load(f)
    return plain_memory(mem, device=device, dtype=dtype), data, tree
