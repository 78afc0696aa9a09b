"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch, run in fp32 on the host) of the
reference's region-masked cross-attention.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this; the product package never does.

Follows:
  scaled_dot_product_attention_regionstate  /root/reference/source/modules/attention_modify.py:74-103
  weight_func (lambda)                      /root/reference/source/app.py:1004
  AttnProcessor2_0.__call__                 /root/reference/source/modules/attention_modify.py:414-503
  AttnProcessor.__call__ (baddbmm variant)  /root/reference/source/modules/attention_modify.py:107-207, :39-70

Pinned: tests/test_oracle_attention.py runs the *unmodified* reference module (oracle/ref_loader.py)
next to this restatement in the build container and requires bit-identical fp32 outputs; the same
vectors are committed under tests/golden/attn_*.npz so the pin travels to the GPU box.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F


def weight_func(w: torch.Tensor, sigma, qk: torch.Tensor) -> torch.Tensor:
    """app.py:1004 -- beta * W with beta = sigma * std over the WHOLE score tensor (unbiased)."""
    return w * sigma * qk.std()


def region_attention(
    query: torch.Tensor,  # [B, H, L, D]
    key: torch.Tensor,  # [B, H, S, D]
    value: torch.Tensor,  # [B, H, S, D]
    region_state: torch.Tensor,  # [B', L, S] fp32
    sigma,
    attn_mask: Optional[torch.Tensor] = None,
    scale: Optional[float] = None,
    weight_fn=weight_func,
) -> torch.Tensor:
    """attention_modify.py:74-103, operation by operation."""
    L, S = query.size(-2), key.size(-2)
    scale_factor = 1 / math.sqrt(query.size(-1)) if scale is None else scale
    attn_bias = torch.zeros(L, S, dtype=query.dtype, device=query.device)
    if attn_mask is not None:
        if attn_mask.dtype == torch.bool:
            attn_mask.masked_fill_(attn_mask.logical_not(), float("-inf"))  # sic (:87)
        else:
            attn_bias += attn_mask
    a = query @ key.transpose(-2, -1) * scale_factor
    a += attn_bias
    B, H, Lq, Sk = a.shape
    a = a.reshape((-1, Lq, Sk))
    cw = weight_fn(region_state, sigma, a)
    repeat_time = a.shape[0] // cw.shape[0]
    a += torch.repeat_interleave(cw, repeats=repeat_time, dim=0)
    a = a.reshape((-1, H, Lq, Sk))
    a = torch.softmax(a, dim=-1)
    return a @ value


def score_std(query: torch.Tensor, key: torch.Tensor, scale: Optional[float] = None) -> torch.Tensor:
    """The scalar the reference's weight_func sees: std of scale*QK^T over everything (:90-95)."""
    scale_factor = 1 / math.sqrt(query.size(-1)) if scale is None else scale
    return (query @ key.transpose(-2, -1) * scale_factor).std()


def processor_forward(attn, hidden_states, encoder_hidden_states=None, region_prompt=None, ip_branch=None,
                      attention_mask=None):
    """attention_modify.py:414-503 for the SD-1.5 case (3-D input, no norms).

    ``attn`` is duck-typed: to_q/to_k/to_v/to_out, heads, residual_connection, rescale_output_factor (and
    prepare_attention_mask when a mask is given: :448-452 -- the mask reaches the region function as a 4-D
    [B, heads, -1, S] tensor, which that function ignores when bool (:86-87) and cannot add when float (:89 raises)).
    """
    residual = hidden_states
    img_sequence_length = hidden_states.shape[1]
    is_xattn = encoder_hidden_states is not None and region_prompt is not None
    batch_size = hidden_states.shape[0]
    if attention_mask is not None:
        sequence_length = (hidden_states if encoder_hidden_states is None else encoder_hidden_states).shape[1]
        attention_mask = attn.prepare_attention_mask(attention_mask, sequence_length, batch_size)
        attention_mask = attention_mask.view(batch_size, attn.heads, -1, attention_mask.shape[-1])
    query = attn.to_q(hidden_states)
    ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
    key, value = attn.to_k(ctx), attn.to_v(ctx)
    inner_dim = key.shape[-1]
    head_dim = inner_dim // attn.heads
    query = query.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
    key = key.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
    value = value.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
    if is_xattn and isinstance(region_prompt["region_state"], dict):
        w = region_prompt["region_state"][img_sequence_length].to(query.device)
        out = region_attention(query, key, value, w, region_prompt["sigma"], attn_mask=attention_mask,
                               weight_fn=region_prompt["weight_func"])
    else:
        out = F.scaled_dot_product_attention(query, key, value, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
    out = out.transpose(1, 2).reshape(batch_size, -1, attn.heads * head_dim).to(query.dtype)
    if ip_branch is not None:  # IP-Adapter image-prompt terms (attention_modify.py:640-682), see oracle/ip_adapter.py
        out = ip_branch(out, query, batch_size, head_dim)
    out = attn.to_out[0](out)
    out = attn.to_out[1](out)
    if getattr(attn, "residual_connection", False):
        out = out + residual
    return out / getattr(attn, "rescale_output_factor", 1.0)


def processor_forward_baddbmm(attn, hidden_states, encoder_hidden_states=None, region_prompt=None, attention_mask=None):
    """attention_modify.py:107-207 (``AttnProcessor``, the ``torch.baddbmm`` variant) with ``get_attention_scores``
    (:39-70) for the SD-1.5 case (3-D input, no norms, no upcasts).  A mask (prepared by ``attn.prepare_attention_mask``,
    :144: [B*heads, 1 or L, S] in the query's dtype) is the ``baddbmm`` input with beta = 1 (:52-63): a true additive mask,
    and the std the weight_func sees is over the masked scores (:166).  Differences from the SDPA-style processor
    that are part of the contract: scores are ``alpha = attn.scale`` times ``Q K^T`` formed by ``torch.baddbmm`` over a
    [B*H, L, S] batch (the std the weight_func sees is over that tensor: same numbers, another shape), softmax + ``bmm``.
    """
    residual = hidden_states
    img_sequence_length = hidden_states.shape[1]
    is_xattn = encoder_hidden_states is not None and region_prompt is not None
    if attention_mask is not None:
        sequence_length = (hidden_states if encoder_hidden_states is None else encoder_hidden_states).shape[1]
        attention_mask = attn.prepare_attention_mask(attention_mask, sequence_length, batch_size=hidden_states.shape[0])
    query = attn.to_q(hidden_states)
    ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
    key, value = attn.to_k(ctx), attn.to_v(ctx)

    def head_to_batch_dim(t):  # diffusers Attention.head_to_batch_dim, out_dim=3
        b, n, c = t.shape
        return t.reshape(b, n, attn.heads, c // attn.heads).permute(0, 2, 1, 3).reshape(b * attn.heads, n, c // attn.heads)

    query, key, value = head_to_batch_dim(query), head_to_batch_dim(key), head_to_batch_dim(value)
    if attention_mask is None:
        empty = torch.empty(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype, device=query.device)
        scores = torch.baddbmm(empty, query, key.transpose(-1, -2), beta=0, alpha=attn.scale).to(query.dtype)  # :39-70
    else:
        scores = torch.baddbmm(attention_mask, query, key.transpose(-1, -2), beta=1, alpha=attn.scale).to(query.dtype)
    if is_xattn and isinstance(region_prompt["region_state"], dict):
        w = region_prompt["region_state"][img_sequence_length].to(query.device)
        cw = region_prompt["weight_func"](w, region_prompt["sigma"], scores)
        scores += torch.repeat_interleave(cw, repeats=scores.shape[0] // cw.shape[0], dim=0)
    probs = scores.softmax(dim=-1).to(query.dtype)
    out = torch.bmm(probs, value)
    bh, n, d = out.shape  # batch_to_head_dim
    out = out.reshape(bh // attn.heads, attn.heads, n, d).permute(0, 2, 1, 3).reshape(bh // attn.heads, n, d * attn.heads)
    out = attn.to_out[0](out)
    out = attn.to_out[1](out)
    if getattr(attn, "residual_connection", False):
        out = out + residual
    return out / getattr(attn, "rescale_output_factor", 1.0)


class OracleAttnProcessor:
    """The restated reference processor behind the diffusers processor protocol (same kwarg names as
    attention_modify.py:414-424), so the CPU baseline can drive the same UNet host module."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0,
                 region_prompt=None, ip_adapter_masks=None):
        return processor_forward(attn, hidden_states, encoder_hidden_states, region_prompt)
