"""Model API of the reference (must3r/model/__init__.py): load_model, Dust3rEncoder, MUSt3R, CausalMUSt3R,
ActivationType, apply_activation, get_pointmaps_activation, get_dtype - backed by the sm_100a kernels."""
from __future__ import annotations

import re
from functools import partial  # noqa: F401  (constructor strings in checkpoints use it)

import torch
import torch.nn as nn  # noqa: F401

from .common import ActivationType, set_precision, get_precision  # noqa: F401
from .encoder import Dust3rEncoder  # noqa: F401
from .decoder import MUSt3R, CausalMUSt3R, MEMORY_MODES  # noqa: F401


def apply_activation(xyz, activation):
    """must3r/model/blocks/head.py:13-21 (+ must3r/tools/geometry.py:14-18)"""
    if isinstance(activation, str):
        activation = ActivationType(activation)
    if activation == ActivationType.NORM_EXP:
        d = xyz.norm(dim=-1, keepdim=True)
        return xyz / d.clip(min=1e-8) * torch.expm1(d)
    elif activation == ActivationType.LINEAR:
        return xyz
    raise ValueError(f"Unknown activation: {activation}")


def get_pointmaps_activation(decoder, verbose=True):
    """must3r/model/__init__.py:8-15"""
    try:
        pointmaps_activation = decoder.pointmaps_activation
    except Exception:
        pointmaps_activation = ActivationType.NORM_EXP
    if verbose:
        print(f'pointmaps_activation set to {pointmaps_activation}')
    return pointmaps_activation


def get_dtype(amp):
    """must3r/model/__init__.py:18-27"""
    if amp == "fp16":
        return torch.float16
    elif amp == "bf16":
        return torch.bfloat16
    assert not amp
    return torch.float32


def convert_decoder_args(decoder_args):
    """must3r/model/__init__.py:53-63: CausalMUSt3R -> MUSt3R, landscape_only -> False"""
    decoder_args = decoder_args.replace(' ', '')
    for k, v in {'CausalMUSt3R': 'MUSt3R', 'landscape_only=True': "landscape_only=False"}.items():
        decoder_args = decoder_args.replace(k, v)
    if 'landscape_only=False' not in decoder_args:
        decoder_args = decoder_args[:-1] + ",landscape_only=False)"
    return decoder_args


def set_image_size_in_args(model_args, img_size, verbose=True):
    """must3r/model/__init__.py:66-108: rewrite img_size and the adaptive RoPE name ('RoPE100_512:768')."""
    model_args = model_args.replace(' ', '')
    match_size = re.search(r'img_size=\((\d+),(\d+)\)', model_args)
    if not match_size:
        raise ValueError("No image_size tuple found in model args")
    h, w = map(int, match_size.groups())
    assert h == w
    if verbose:
        print(f"image_size {h} -> {img_size}")
    m = re.search(r"pos_embed='([A-Za-z]+)(\d+)\_(\d+):(\d+)'", model_args)
    if m:
        prefix, freq, base_size, new_size = m.groups()
        freq, base_size, new_size = int(freq), int(base_size), int(new_size)
        pos_embed_is_arg = True
    else:
        m = re.search(r"pos_embed='([A-Za-z]+)(\d+)'", model_args)
        if m:
            prefix, freq = m.groups()
            freq = int(freq)
            pos_embed_is_arg = True
        else:
            prefix, freq = "RoPE", 100
            pos_embed_is_arg = False
        base_size = new_size = h
    if verbose:
        print(f"Parsed pos_embed: {prefix}{freq}, base size = {base_size}")
    if img_size != h:
        model_args = model_args.replace(f'img_size=({h},{h})', f'img_size=({img_size},{img_size})')
    if img_size != new_size:
        new_pos_embed = f"{prefix}{freq}_{base_size}:{img_size}"
        if pos_embed_is_arg:
            model_args = re.sub(r"(pos_embed=')(?:[A-Za-z]+\d+(?:_\d+:\d+)?)(')", rf"\1{new_pos_embed}\2", model_args)
        else:
            model_args = model_args[:-1] + ",pos_embed='" + new_pos_embed + "')"
    return model_args


def load_model(chkpt_path, encoder=None, decoder=None, device='cuda', img_size=None, memory_mode=None, verbose=True):
    """must3r/model/__init__.py:30-50: checkpoint {'args': Namespace(encoder=str, decoder=str), 'encoder': sd,
    'decoder': sd}; the constructor strings are evaluated against this package's classes."""
    ckpt = torch.load(chkpt_path, map_location='cpu', weights_only=False)
    encoder_args = encoder or ckpt['args'].encoder
    decoder_args = decoder or convert_decoder_args(ckpt['args'].decoder)
    if img_size is not None:
        encoder_args = set_image_size_in_args(encoder_args, img_size, verbose=verbose)
        decoder_args = set_image_size_in_args(decoder_args, img_size, verbose=verbose)
    scope = {"Dust3rEncoder": Dust3rEncoder, "MUSt3R": MUSt3R, "CausalMUSt3R": CausalMUSt3R, "partial": partial,
             "nn": nn, "torch": torch, "ActivationType": ActivationType}
    enc = eval(encoder_args, scope)
    dec = eval(decoder_args, scope)
    if memory_mode is not None:
        dec.change_memory_mode(memory_mode)
    enc.load_state_dict(ckpt['encoder'], strict=True)
    dec.load_state_dict(ckpt['decoder'], strict=True)
    enc.to(device)
    dec.to(device)
    enc.eval()
    dec.eval()
    return enc, dec
