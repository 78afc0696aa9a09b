"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's IP-Adapter attention processor
(/root/reference/source/modules/attention_modify.py:506-700, ``IPAdapterAttnProcessor2_0``) on top of the
restated region-masked attention of oracle/attention.py.  Never imported by the product package.

Pinned: ``tests/test_oracle_ip_adapter.py`` runs the UNMODIFIED reference class (loaded through
oracle/ref_loader.py) against this restatement -- identical outputs, with and without spatial masks.
PARITY UNPINNED for one third-party piece: ``IPAdapterMaskProcessor.downsample`` lives in
``diffusers==0.27.2`` (requirements.txt:3; neither vendored nor installed here, call site
attention_modify.py:671-673); ``ip_mask_downsample`` restates its published algorithm and the pin test
injects this very function into the reference's stubbed ``diffusers.image_processor``.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .attention import processor_forward


def ip_mask_downsample(mask: torch.Tensor, batch_size: int, num_queries: int, value_embed_dim: int) -> torch.Tensor:
    """diffusers.image_processor.IPAdapterMaskProcessor.downsample: [1, H, W] mask -> [B, num_queries, C]."""
    o_h, o_w = mask.shape[1], mask.shape[2]
    ratio = o_w / o_h
    mask_h = int(math.sqrt(num_queries / ratio))
    mask_h = int(mask_h) + int((num_queries % int(mask_h)) != 0)
    mask_w = num_queries // mask_h
    m = F.interpolate(mask.unsqueeze(0), size=(mask_h, mask_w), mode="bicubic").squeeze(0)
    if m.shape[0] < batch_size:
        m = m.repeat(batch_size, 1, 1)
    m = m.view(m.shape[0], -1)
    area = mask_h * mask_w
    if area < num_queries:
        m = F.pad(m, (0, num_queries - m.shape[1]), value=0.0)
    if area > num_queries:
        m = m[:, :num_queries]
    return m.view(m.shape[0], m.shape[1], 1).repeat(1, 1, value_embed_dim)


class OracleIPAdapterProcessor(nn.Module):
    """attention_modify.py:506-700 (constructor :520-546, call :549-700)."""

    def __init__(self, hidden_size, cross_attention_dim=None, num_tokens=(4,), scale=1.0):
        super().__init__()
        if not isinstance(num_tokens, (tuple, list)):
            num_tokens = [num_tokens]
        if not isinstance(scale, list):
            scale = [scale] * len(num_tokens)
        self.num_tokens, self.scale = num_tokens, scale
        self.to_k_ip = nn.ModuleList([nn.Linear(cross_attention_dim, hidden_size, bias=False) for _ in num_tokens])
        self.to_v_ip = nn.ModuleList([nn.Linear(cross_attention_dim, hidden_size, bias=False) for _ in num_tokens])

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0,
                 region_prompt=None, ip_adapter_masks=None):
        if encoder_hidden_states is None:
            return processor_forward(attn, hidden_states, None, region_prompt)
        if isinstance(encoder_hidden_states, tuple):
            encoder_hidden_states, ip_hidden_states = encoder_hidden_states
        else:
            end = encoder_hidden_states.shape[1] - self.num_tokens[0]
            encoder_hidden_states, ip_hidden_states = encoder_hidden_states[:, :end], [encoder_hidden_states[:, end:]]
        masks = ip_adapter_masks if ip_adapter_masks is not None else [None] * len(self.scale)

        def ip_branch(hidden, query, batch_size, head_dim):
            for cur, sc, to_k, to_v, mask in zip(ip_hidden_states, self.scale, self.to_k_ip, self.to_v_ip, masks):
                k = to_k(cur).view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
                v = to_v(cur).view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
                cur = F.scaled_dot_product_attention(query, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
                cur = cur.transpose(1, 2).reshape(batch_size, -1, attn.heads * head_dim).to(query.dtype)
                if mask is not None:
                    cur = cur * ip_mask_downsample(mask, batch_size, cur.shape[1], cur.shape[2]).to(query.dtype)
                hidden = hidden + sc * cur
            return hidden

        return processor_forward(attn, hidden_states, encoder_hidden_states, region_prompt, ip_branch)
