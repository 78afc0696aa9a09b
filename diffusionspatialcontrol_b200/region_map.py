"""Region-weight map builder on the GPU -- host-side mirror of the reference's
``encode_region_map_sp`` / ``encode_region_map`` (reference
source/modules/encode_region_map_function.py:21-77 and :79-124): same arguments, same return type
(``{L: fp32 [B', L, n_tok]}``), same quirks, but the per-resolution work (mask
binarise, INTER_CUBIC downsample, ``== max``, ``*S``/``-S'``, accumulation into token columns) runs in
two small CUDA kernels (dsc_region_downsample / dsc_region_accumulate) and the maps are returned ON
THE DEVICE, so the attention processor never re-uploads them (the reference copies them H2D on every
one of the 400 attention calls of a generation, attention_modify.py:481).
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check, lib


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _tokenize(tokenizer, phrase: str) -> List[int]:
    return list(
        tokenizer(
            phrase,
            max_length=getattr(tokenizer, "model_max_length", 77),
            truncation=True,
            add_special_tokens=False,
        ).input_ids
    )


def _spans(ids: Optional[Sequence[int]], phrases: List[List[int]]):
    """(region, start, len) for every occurrence, in the reference's application order
    (regions in dict order, occurrences left to right, :59-69)."""
    out = []
    found = [False] * len(phrases)
    if ids is None:
        return out, found
    for r, toks in enumerate(phrases):
        n = len(toks)
        if n == 0:
            continue
        for idx in range(len(ids)):
            if ids[idx : idx + n] == toks:
                out.append((r, idx, n))
                found[r] = True
    return out, found


@_lib.nvtx("region_downsample")
def downsample_regions(maps: torch.Tensor, w_r: int, h_r: int):
    """maps: uint8 [R, Hpx, Wpx] on the device (255 = outside).  Returns (ds uint8 [R, h_r*w_r], any uint32 [R])."""
    if not maps.is_cuda or maps.dtype != torch.uint8 or maps.dim() != 3:
        raise ValueError("maps must be a CUDA uint8 tensor [R, H, W]")
    maps = maps.contiguous()
    R, Hpx, Wpx = maps.shape
    ds = torch.empty((R, h_r * w_r), dtype=torch.uint8, device=maps.device)
    any_set = torch.zeros((max(R, 1),), dtype=torch.int32, device=maps.device)
    with torch.cuda.device(maps.device):
        check(lib.dsc_region_downsample(maps.data_ptr(), R, Hpx, Wpx, w_r, h_r, ds.data_ptr(), any_set.data_ptr(),
                                        _stream(maps.device)))
    return ds, any_set


@_lib.nvtx("region_accumulate")
def accumulate_regions(ds, any_set, weight, mask_outsides, spans, n_tok: int) -> torch.Tensor:
    """W[L_r, n_tok] fp32 on the device from the per-region binary maps and the token spans."""
    device = ds.device
    R, L_r = ds.shape
    W = torch.empty((L_r, n_tok), dtype=torch.float32, device=device)
    n = len(spans)
    sp = torch.tensor(spans if n else [(0, 0, 0)], dtype=torch.int32).t().contiguous().to(device)
    wt = torch.tensor(list(weight) or [0.0], dtype=torch.float64, device=device)
    mo = torch.tensor(list(mask_outsides) or [0.0], dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        check(lib.dsc_region_accumulate(ds.data_ptr(), any_set.data_ptr(), R, L_r, wt.data_ptr(), mo.data_ptr(),
                                        sp[0].data_ptr(), sp[1].data_ptr(), sp[2].data_ptr(), n, n_tok,
                                        W.data_ptr(), _stream(device)))
    return W


def encode_region_map_sp(state, tokenizer, unet, width, height, scale_ratio=8, text_ids=None,
                         do_classifier_free_guidance=True, device=None):
    """Mirror of reference encode_region_map_function.py:21-77 (one prompt, all UNet resolutions)."""
    if text_ids is None:
        return torch.FloatTensor(0)
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("diffusionspatialcontrol_b200 has no CPU path: region maps are built on a CUDA device")
    uncond, cond = text_ids[0], text_ids[1]
    to_list = lambda a: (np.asarray(a.cpu() if isinstance(a, torch.Tensor) else a).reshape(-1).tolist()
                         if isinstance(a, (np.ndarray, torch.Tensor)) else None)
    cond, uncond = to_list(cond), to_list(uncond)
    c = len(cond)

    phrases, maps, weight, outside = [], [], [], []
    if state is not None:
        for k, v in state.items():
            if v["map"] is None:
                continue
            phrases.append(_tokenize(tokenizer, k))
            m = np.asarray(v["map"])
            if m.dtype != np.uint8:
                # the reference tests `map < 255` in the map's OWN dtype (encode_region_map_function.py:49); a uint8 cast
                # would wrap 300 -> 44 ("inside") or -1 -> 255 ("outside"): binarise here, in the original dtype, into the
                # two uint8 values the device-side `< 255` test tells apart
                m = np.where(m < 255, 0, 255).astype(np.uint8)
            maps.append(np.ascontiguousarray(m))
            weight.append(float(v["weight"]))
            outside.append(float(v["mask_outsides"]))
    cond_spans, cond_found = _spans(cond, phrases)
    uncond_spans, uncond_found = _spans(uncond, phrases)
    for r, toks in enumerate(phrases):
        if not (cond_found[r] or uncond_found[r]):
            print(f"tokens {toks} not found in text")

    # upload the raw maps once; regions of one shape go through one batched launch per resolution
    groups: Dict[tuple, List[int]] = {}
    for r, m in enumerate(maps):
        if m.ndim != 2:
            raise ValueError("region maps must be 2-D uint8 arrays")
        groups.setdefault(m.shape, []).append(r)
    dev_groups = [
        (idx, torch.from_numpy(np.stack([maps[r] for r in idx]).astype(np.uint8, copy=False)).to(device))
        for idx in groups.values()
    ]

    w_tensors = {}
    for _ in unet.down_blocks:
        w_r, h_r = int(math.ceil(width / scale_ratio)), int(math.ceil(height / scale_ratio))
        L_r = w_r * h_r
        R = len(maps)
        if R:
            ds = torch.empty((R, L_r), dtype=torch.uint8, device=device)
            any_set = torch.zeros((R,), dtype=torch.int32, device=device)
            for idx, stack in dev_groups:
                d, a = downsample_regions(stack, w_r, h_r)
                ii = torch.tensor(idx, device=device)
                ds[ii] = d
                any_set[ii] = a[: len(idx)]
        else:
            ds = torch.zeros((1, L_r), dtype=torch.uint8, device=device)
            any_set = torch.zeros((1,), dtype=torch.int32, device=device)
        ret_cond = accumulate_regions(ds, any_set, weight, outside, cond_spans, c)[None]
        if do_classifier_free_guidance:
            ret_uncond = (ret_cond if uncond_spans == cond_spans
                          else accumulate_regions(ds, any_set, weight, outside, uncond_spans, c)[None])
            w_tensors[L_r] = torch.cat([ret_uncond, ret_cond])
        else:
            w_tensors[L_r] = ret_cond
        scale_ratio *= 2
    return w_tensors


def encode_region_map(pipe, state, width, height, num_images_per_prompt, text_ids=None, device=None):
    """Mirror of reference encode_region_map_function.py:79-124 (prompt split, CFG concat, repeat),
    including the quirk at :91 -- the negative ids are overwritten with the positive ids."""
    negative_prompt_tokens_id, prompt_tokens_id = text_ids[0], text_ids[1]
    if prompt_tokens_id is None:
        return torch.FloatTensor(0)
    if isinstance(prompt_tokens_id, torch.Tensor):
        prompt_tokens_id = prompt_tokens_id.cpu()
    prompt_tokens_id = np.array(prompt_tokens_id)
    negative_prompt_tokens_id = np.array(prompt_tokens_id) if negative_prompt_tokens_id is not None else None

    number_prompt = prompt_tokens_id.shape[0]
    prompt_tokens_id = np.split(prompt_tokens_id, number_prompt)
    negative_prompt_tokens_id = (
        np.split(negative_prompt_tokens_id, number_prompt) if negative_prompt_tokens_id is not None else None
    )
    lst_prompt_map = []
    if not isinstance(state, list):
        state = [state]
    if len(state) < number_prompt:
        state = [state] + [None] * int(number_prompt - len(state))
    for i in range(number_prompt):
        ids = ([negative_prompt_tokens_id[i], prompt_tokens_id[i]] if negative_prompt_tokens_id is not None
               else [None, prompt_tokens_id[i]])
        lst_prompt_map.append(
            encode_region_map_sp(state[i], pipe.tokenizer, pipe.unet, width, height,
                                 scale_ratio=pipe.vae_scale_factor, text_ids=ids,
                                 do_classifier_free_guidance=pipe.do_classifier_free_guidance, device=device)
        )
    region_state_sp = {}
    for d in lst_prompt_map:
        for key, tensor in d.items():
            region_state_sp[key] = torch.cat((region_state_sp[key], tensor)) if key in region_state_sp else tensor
    # same values and shape as the reference's dict; rows are laid out 80 floats apart (the kernels' fast W layout)
    from .attention import padded_region_map

    return {key: padded_region_map(tensor.repeat(num_images_per_prompt, 1, 1)) for key, tensor in region_state_sp.items()}
