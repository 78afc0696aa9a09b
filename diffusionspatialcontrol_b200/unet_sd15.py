"""SD-1.5-architecture UNet in plain PyTorch -- the HOST for the attention processor.

diffusers is not installed in this image, so the module tree the reference gets from diffusers 0.27
(``UNet2DConditionModel`` as copied in reference source/modules/u_net_condition_modify.py:70-1316,
defaults :175-197 with SD-1.5's cross_attention_dim=768) is restated here at the interface level only:
what matters for the hot path is that every ``Attention`` module calls
``processor(attn, hidden_states, encoder_hidden_states=, attention_mask=, **cross_attention_kwargs)``
with the kwargs filtered by the processor's signature (diffusers ``Attention.forward``), that
``cross_attention_kwargs`` reaches both attn1 and attn2 of every block (:1220-1228, :1250-1257,
:1285-1294), and that ``set_attn_processor`` / ``attn_processors`` behave like :717-749.  Everything in
here is ordinary PyTorch GPU ops ("the rest of the UNet stays on PyTorch", BASELINE north_star);
weights are random-init (no checkpoints are available offline).
"""
from __future__ import annotations

import inspect
import os
import math
from typing import Dict, Optional, Union

import torch
import torch.nn as nn
import torch.nn.functional as F


class DefaultAttnProcessor:
    """Stock scaled-dot-product attention (what diffusers' AttnProcessor2_0 does without regions)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0):
        B = hidden_states.shape[0]
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q, k, v = attn.to_q(hidden_states), attn.to_k(ctx), attn.to_v(ctx)
        hd = k.shape[-1] // attn.heads
        q = q.view(B, -1, attn.heads, hd).transpose(1, 2)
        k = k.view(B, -1, attn.heads, hd).transpose(1, 2)
        v = v.view(B, -1, attn.heads, hd).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(B, -1, attn.heads * hd).to(q.dtype)
        return attn.to_out[1](attn.to_out[0](o))


class Attention(nn.Module):
    """Duck-type of diffusers.models.attention_processor.Attention (the attributes the reference
    processors read: attention_modify.py:428-501)."""

    def __init__(self, query_dim: int, cross_attention_dim: Optional[int] = None, heads: int = 8, dim_head: int = 64):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head**-0.5
        self.scale_qk = True
        self.is_cross_attention = cross_attention_dim is not None
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_v = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        self.spatial_norm = None
        self.group_norm = None
        self.norm_cross = None
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.upcast_attention = False
        self.upcast_softmax = False
        self.processor = DefaultAttnProcessor()
        self._sig_cache = None

    def set_processor(self, processor) -> None:
        self.processor = processor
        self._sig_cache = None

    def prepare_attention_mask(self, attention_mask, target_length, batch_size, out_dim=3):
        if attention_mask is None:
            return None
        if attention_mask.shape[-1] != target_length:
            attention_mask = F.pad(attention_mask, (0, target_length), value=0.0)
        if out_dim == 3 and attention_mask.shape[0] < batch_size * self.heads:
            attention_mask = attention_mask.repeat_interleave(self.heads, dim=0)
        elif out_dim == 4:
            attention_mask = attention_mask.unsqueeze(1).repeat_interleave(self.heads, dim=1)
        return attention_mask

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        # diffusers 0.27 Attention.forward: silently drop kwargs the processor's __call__ does not name
        if self._sig_cache is None or self._sig_cache[0] is not self.processor:
            params = set(inspect.signature(self.processor.__call__).parameters.keys())
            self._sig_cache = (self.processor, params)
        params = self._sig_cache[1]
        kept = {k: w for k, w in cross_attention_kwargs.items() if k in params}
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kept)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    # DSC_GEGLU_SPLIT=0 restores the single projection + chunk (A/B runs)
    SPLIT = os.environ.get("DSC_GEGLU_SPLIT", "1") != "0"

    def forward(self, x):
        if self.SPLIT and x.is_cuda:
            # two projections on row slices of the same weight (same parameters, same state-dict keys): value and gate come
            # out CONTIGUOUS, so gelu and the product run as vectorised elementwise kernels; chunk() of one [.., 2 * inner]
            # output hands strided halves to both (non-vectorised generic kernels: 4 ms of a 32 ms UNet step at batch 16,
            # profiles/r2_unet_host_groupnorm_nhwc.log).  Plain PyTorch ops: the UNet host is not part of the hot path.
            inner = self.proj.out_features // 2
            w, b = self.proj.weight, self.proj.bias
            gate = F.gelu(F.linear(x, w[inner:], None if b is None else b[inner:]))
            return F.linear(x, w[:inner], None if b is None else b[:inner]).mul_(gate)
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, encoder_hidden_states=None, cross_attention_kwargs=None):
        kw = cross_attention_kwargs or {}
        x = x + self.attn1(self.norm1(x), encoder_hidden_states=None, **kw)
        x = x + self.attn2(self.norm2(x), encoder_hidden_states=encoder_hidden_states, **kw)
        x = x + self.ff(self.norm3(x))
        return x


class GroupNorm(nn.GroupNorm):
    """``nn.GroupNorm`` (same parameters, same state-dict keys) that keeps a channels_last CUDA activation in channels_last.

    ``F.group_norm`` on CUDA first makes its input NCHW-contiguous and returns NCHW; with channels_last convolutions on both
    sides that is two permuting copies of the activation per normalisation (in round 1: 709 such copy launches, 35 % of the
    device time of a UNet step, profiles/r1_ncu_launch_list_bench_summary.csv).  Here the statistics are two reductions over
    the NHWC view (fp32 accumulation and fp32 results straight from the fp16 data), and normalisation + affine is one fused
    multiply-add with per-(sample, channel) scale / shift -- all coalesced, nothing is re-laid out.  Plain PyTorch ops: the
    UNet host is not part of the hot path (BASELINE north_star: "the rest of the UNet stays on PyTorch GPU ops").
    Any other input takes ``F.group_norm``."""

    ONE_PASS_STATS = os.environ.get("DSC_GN_ONE_PASS", "1") != "0"  # A/B switch
    # activations of at least this many bytes (default: all) take their statistics per CHANNEL first (two reductions over
    # the rows of the [N, H*W, C] view: long coalesced rows, fp32 results) and fold the channels of a group afterwards on
    # [N, C] -- the one-pass Welford reduction over (H*W, C/G) of the [N, H*W, G, C/G] view runs at a tenth of the HBM rate
    # on the large activations (95 / 183 / 287 us for 40 / 80 / 120 MB against 66 / 71 / 104 us) and returns 16-bit results;
    # UNet step at batch 16 in a graph: 28.11 ms (Welford everywhere) -> 26.69 ms (per channel everywhere) -> 25.64 ms (10
    # instead of 16 small kernels per normalisation), cosine 0.999999 (profiles/r2_unet_host_gn_stats_ab.jsonl,
    # r2_unet_host_gn_mode_ab.log)
    PER_CHANNEL_BYTES = int(os.environ.get("DSC_GN_PER_CHANNEL_BYTES", "0"))

    def _affine32(self):
        """fp32 [1, G, C/G] views of weight / bias, converted once per parameter version (not once per call)."""
        key = (self.weight._version, self.bias._version, self.weight.data_ptr(), self.weight.device)
        if getattr(self, "_a32_key", None) != key:
            G = self.num_groups
            self._a32 = (self.weight.detach().float().view(1, G, -1), self.bias.detach().float().view(1, G, -1))
            self._a32_key = key
        return self._a32

    def _eps_t(self, device):
        t = getattr(self, "_eps_tensor", None)
        if t is None or t.device != device:
            t = self._eps_tensor = torch.full((1,), float(self.eps), dtype=torch.float32, device=device)
        return t

    def forward(self, x: torch.Tensor, silu: bool = False) -> torch.Tensor:
        if (x.is_cuda and x.dim() == 4 and x.dtype != torch.float32 and not x.is_contiguous()
                and x.is_contiguous(memory_format=torch.channels_last)):
            N, C, H, W = x.shape
            G = self.num_groups
            Cg = C // G
            xl = x.permute(0, 2, 3, 1)                      # [N, H, W, C] view of the same memory, contiguous
            xv = xl.reshape(N, H * W, G, Cg)
            if x.numel() * x.element_size() >= self.PER_CHANNEL_BYTES:
                # the [N, G]-sized arithmetic behind the two big reductions is launch-bound (61 normalisations per UNet
                # step): 10 small kernels instead of 16 -- 1/n folded into the fused multiply-adds, fp32 copies of weight /
                # bias kept, scale / shift written straight into their 16-bit tensors
                xc = xl.reshape(N, H * W, C)
                inv_n = 1.0 / float(H * W * Cg)
                a = xc.sum(dim=1, dtype=torch.float32).view(N, G, Cg).sum(-1, keepdim=True)           # sum x      [N, G, 1]
                q = torch.linalg.vector_norm(
                    torch.linalg.vector_norm(xc, dim=1, dtype=torch.float32).view(N, G, Cg), dim=-1, keepdim=True)  # sqrt(sum x^2)
                u = torch.addcmul(q * q, a, a, value=-inv_n).clamp_min_(0.0)                          # n * var
                rstd = torch.rsqrt(torch.add(self._eps_t(x.device), u, alpha=inv_n))                  # [N, G, 1]
                w32, b32 = self._affine32()
                ss = torch.empty((2, N, G, Cg), dtype=x.dtype, device=x.device)
                torch.mul(rstd, w32, out=ss[0])                                                       # scale
                torch.addcmul(b32, a * rstd, w32, value=-inv_n, out=ss[1])                            # shift = b - mean * scale
                y = torch.addcmul(ss[1].view(N, 1, 1, C), xl, ss[0].view(N, 1, 1, C))
                if silu:
                    y = F.silu(y, inplace=True)
                return y.permute(0, 3, 1, 2)
            elif self.ONE_PASS_STATS:
                # ONE reduction over the activation (Welford, fp32 accumulation inside the kernel; the results come back in
                # the activation's 16-bit type: the mean is off by <= 2^-11 |mean|, below the activation's own rounding
                # for any |mean| / std a UNet produces) instead of a sum and a norm pass
                var, mean = torch.var_mean(xv, dim=(1, 3), correction=0)
                mean = mean.float()
                rstd = torch.rsqrt(var.float() + self.eps)
            else:
                n = float(H * W * Cg)
                mean = xv.sum(dim=(1, 3), dtype=torch.float32) / n                              # [N, G]
                ex2 = torch.linalg.vector_norm(xv, dim=(1, 3), dtype=torch.float32).square() / n
                rstd = torch.rsqrt((ex2 - mean * mean).clamp_min(0.0) + self.eps)
            scale = rstd[:, :, None] * self.weight.float().view(1, G, Cg)                   # [N, G, Cg]
            shift = self.bias.float().view(1, G, Cg) - mean[:, :, None] * scale
            y = torch.addcmul(shift.to(x.dtype).view(N, 1, 1, C), xl, scale.to(x.dtype).view(N, 1, 1, C))
            if silu:
                y = F.silu(y, inplace=True)
            return y.permute(0, 3, 1, 2)                    # NCHW shape, channels_last strides
        y = super().forward(x)
        return F.silu(y) if silu else y


def conv1x1(conv: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """A 1x1 convolution of a channels_last CUDA activation as ONE GEMM over its [N*H*W, C] view with the bias in the GEMM's
    epilogue (cuBLAS) -- the cuDNN path adds the bias with a separate broadcast kernel (32 + 10 such launches per UNet
    step).  Same parameters, same result shape and memory format; anything else takes the convolution."""
    if (x.is_cuda and x.dim() == 4 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last)
            and conv.kernel_size == (1, 1) and conv.stride == (1, 1) and conv.padding == (0, 0) and conv.groups == 1):
        w = conv.weight.reshape(conv.out_channels, conv.in_channels)
        return F.linear(x.permute(0, 2, 3, 1), w, conv.bias).permute(0, 3, 1, 2)
    return conv(x)


class Transformer2DModel(nn.Module):
    def __init__(self, channels, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.norm = GroupNorm(32, channels, eps=1e-6)
        self.proj_in = nn.Conv2d(channels, channels, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(channels, heads, dim_head, cross_attention_dim)])
        self.proj_out = nn.Conv2d(channels, channels, 1)

    def forward(self, x, encoder_hidden_states=None, cross_attention_kwargs=None):
        B, C, H, W = x.shape
        res = x
        h = conv1x1(self.proj_in, self.norm(x))
        h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
        for blk in self.transformer_blocks:
            h = blk(h, encoder_hidden_states, cross_attention_kwargs)
        h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
        return conv1x1(self.proj_out, h) + res


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_ch=1280):
        super().__init__()
        self.norm1 = GroupNorm(32, cin, eps=1e-5)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_ch, cout)
        self.norm2 = GroupNorm(32, cout, eps=1e-5)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        # conv1's bias rides on the time-embedding add: one per-(sample, channel) broadcast pass over h instead of two
        c1 = self.conv1
        h = F.conv2d(self.norm1(x, silu=True), c1.weight, None, c1.stride, c1.padding, c1.dilation, c1.groups)
        t = self.time_emb_proj(F.silu(temb))
        h = h + (t if c1.bias is None else t + c1.bias)[:, :, None, None]
        h = self.conv2(self.norm2(h, silu=True))
        if self.conv_shortcut is not None:
            x = conv1x1(self.conv_shortcut, x)
        return x + h


class Downsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cin, cout, heads, cross_dim, has_attn, add_down, layers=2):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout) for i in range(layers)])
        self.attentions = (
            nn.ModuleList([Transformer2DModel(cout, heads, cout // heads, cross_dim) for _ in range(layers)])
            if has_attn else None
        )
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, temb, ehs, kw):
        outs = ()
        for i, res in enumerate(self.resnets):
            x = res(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ehs, kw)
            outs += (x,)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs += (x,)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, ch, heads, cross_dim):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch), ResnetBlock2D(ch, ch)])
        self.attentions = nn.ModuleList([Transformer2DModel(ch, heads, ch // heads, cross_dim)])

    def forward(self, x, temb, ehs, kw):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ehs, kw)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cin, cout, prev_out, heads, cross_dim, has_attn, add_up, layers=3):
        super().__init__()
        res = []
        for i in range(layers):
            skip = cin if i == layers - 1 else cout
            rin = prev_out if i == 0 else cout
            res.append(ResnetBlock2D(rin + skip, cout))
        self.resnets = nn.ModuleList(res)
        self.attentions = (
            nn.ModuleList([Transformer2DModel(cout, heads, cout // heads, cross_dim) for _ in range(layers)])
            if has_attn else None
        )
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips, temb, ehs, kw):
        for i, res in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = res(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ehs, kw)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


def timestep_embedding(t: torch.Tensor, dim: int = 320) -> torch.Tensor:
    """Sinusoidal, flip_sin_to_cos=True, freq_shift=0 (reference u_net_condition_modify.py:261-275)."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class UNetSD15(nn.Module):
    """SD-1.5 layout: 16 cross-attention layers (L = 4096/1024/256/64 at 512x512, 8 heads x 40/80/160)."""

    def __init__(self, block_out_channels=(320, 640, 1280, 1280), heads=8, cross_attention_dim=768,
                 in_channels=4, out_channels=4):
        super().__init__()
        ch = block_out_channels
        self.in_channels = in_channels
        self.conv_in = nn.Conv2d(in_channels, ch[0], 3, padding=1)
        self.time_embedding = nn.ModuleDict({"linear_1": nn.Linear(ch[0], ch[0] * 4), "linear_2": nn.Linear(ch[0] * 4, ch[0] * 4)})
        temb = ch[0] * 4
        self.down_blocks = nn.ModuleList()
        cout = ch[0]
        for i, c in enumerate(ch):
            cin, cout = cout, c
            self.down_blocks.append(DownBlock(cin, cout, heads, cross_attention_dim, has_attn=i < len(ch) - 1,
                                              add_down=i < len(ch) - 1))
        self.mid_block = MidBlock(ch[-1], heads, cross_attention_dim)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(ch))
        cout = rev[0]
        for i, c in enumerate(rev):
            prev_out, cout = cout, c
            cin = rev[min(i + 1, len(ch) - 1)]
            self.up_blocks.append(UpBlock(cin, cout, prev_out, heads, cross_attention_dim, has_attn=i > 0,
                                          add_up=i < len(ch) - 1))
        self.conv_norm_out = GroupNorm(32, ch[0], eps=1e-5)
        self.conv_out = nn.Conv2d(ch[0], out_channels, 3, padding=1)
        for m in self.modules():  # ResnetBlock time projections were sized for 1280; rebuild if temb differs
            if isinstance(m, ResnetBlock2D) and m.time_emb_proj.in_features != temb:
                m.time_emb_proj = nn.Linear(temb, m.time_emb_proj.out_features)

    # ---- processor plumbing (reference u_net_condition_modify.py:692-749) -------------------------
    @property
    def attn_processors(self) -> Dict[str, object]:
        return {f"{name}.processor": m.processor for name, m in self.named_modules() if isinstance(m, Attention)}

    def set_attn_processor(self, processor: Union[object, Dict[str, object]]) -> None:
        count = len(self.attn_processors)
        if isinstance(processor, dict) and len(processor) != count:
            raise ValueError(
                f"A dict of processors was passed, but the number of processors {len(processor)} does not match the"
                f" number of attention layers: {count}. Please make sure to pass {count} processor classes."
            )
        for name, m in self.named_modules():
            if isinstance(m, Attention):
                m.set_processor(processor[f"{name}.processor"] if isinstance(processor, dict) else processor)

    def forward(self, sample, timestep, encoder_hidden_states, cross_attention_kwargs=None):
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.float32, device=sample.device)
        timestep = timestep.reshape(-1).expand(sample.shape[0])
        t_emb = timestep_embedding(timestep, self.conv_in.out_channels).to(sample.dtype)
        temb = self.time_embedding["linear_2"](F.silu(self.time_embedding["linear_1"](t_emb)))
        kw = cross_attention_kwargs
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, temb, encoder_hidden_states, kw)
            skips.extend(outs)
        x = self.mid_block(x, temb, encoder_hidden_states, kw)
        for blk in self.up_blocks:
            x = blk(x, skips, temb, encoder_hidden_states, kw)
        return self.conv_out(self.conv_norm_out(x, silu=True))


def cross_attention_shapes(height: int = 512, width: int = 512):
    """(L, D) of the 16 cross-attention layers in execution order."""
    h, w = math.ceil(height / 8), math.ceil(width / 8)
    lv = [(h * w, 40), (math.ceil(h / 2) * math.ceil(w / 2), 80), (math.ceil(h / 4) * math.ceil(w / 4), 160),
          (math.ceil(h / 8) * math.ceil(w / 8), 160)]
    return [lv[0]] * 2 + [lv[1]] * 2 + [lv[2]] * 2 + [lv[3]] + [lv[2]] * 3 + [lv[1]] * 3 + [lv[0]] * 3
