"""Drop-in attention processor: install with ``unet.set_attn_processor(RegionAttnProcessor())`` exactly
where the reference installs its own (reference source/app.py:479-481).

Mirrors ``AttnProcessor2_0.__call__`` of the reference (source/modules/attention_modify.py:414-503):
same signature (the kwarg names matter -- diffusers drops cross_attention_kwargs the processor does
not name), same dispatch rule (region path iff cross-attention AND region_prompt given AND
region_state is a dict, :437-442,:479), same projections / head split / merge / out-proj / residual.
Only the body of ``scaled_dot_product_attention_regionstate`` (:74-103) is replaced by the two CUDA
passes in libdsc_b200.so.  Self-attention and region-less calls go to
``F.scaled_dot_product_attention`` like the reference (:483-485).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, Optional

import torch
import torch.nn.functional as F

from .attention import (PreparedKV, compact_region_map, padded_region_map, prepare_kv, prepared_supported, region_attention,
                        region_attention_prepared)


def _is_reference_weight_func(fn: Callable) -> bool:
    """The kernels hard-wire ``w * sigma * qk.std()`` (reference app.py:1004); probe the callable."""
    try:
        qk = torch.tensor([0.0, 1.0, 3.0, -2.0])
        w = torch.tensor([[[1.0, -0.5]]])
        got = fn(w, torch.tensor(2.0), qk)
        want = w * 2.0 * qk.std()
        return bool(torch.allclose(got, want, rtol=1e-6, atol=0))
    except Exception:
        return False


class RegionAttnProcessor:
    r"""B200-native replacement for the reference's ``AttnProcessor2_0``.

    cache_kv: reuse ``to_k/to_v(encoder_hidden_states)`` -- and the K / V^T image the tcgen05 kernels multiply from
    (``prepare_kv``) -- while the same text-embedding tensor (same object, same ``_version``) is passed (SURVEY 8f-1;
    numerically identical: the reference recomputes the same projections on each of the 25 steps,
    attention_modify.py:465-466).  With ``cache_kv=False`` the projections and the image are rebuilt on every call.
    ``register_static_kv`` is the same for a captured CUDA graph: projections computed outside the graph into static
    buffers that the replays read.
    """

    def __init__(self, cache_kv: bool = True, max_cached_maps: int = 16, skip_zero_maps: bool = True):
        if not hasattr(F, "scaled_dot_product_attention"):
            raise ImportError("RegionAttnProcessor requires PyTorch 2.0")
        self.cache_kv = cache_kv
        self._w_cache: "OrderedDict[tuple, tuple]" = OrderedDict()
        self._w_zero: dict = {}
        self._w_compact: dict = {}
        self._static_maps: dict = {}
        self.skip_zero_maps = skip_zero_maps
        self._max_cached_maps = max_cached_maps
        self._checked_funcs: dict = {}
        self._sigma_cache: Optional[tuple] = None
        self._kv_cache: dict = {}
        self._kv_static: dict = {}
        self._img_cache: dict = {}
        # Statistics workspace handed to every region call (None: one zero-filled buffer per (device, stream), allocated
        # on first use).  A CUDA-graph capture must set it to a buffer allocated OUTSIDE the capture: the per-stream
        # lookup would allocate inside the capture (graph-private pool, a memset node in every replay, and one buffer
        # shared by every graph captured on that stream).
        self.workspace: Optional[torch.Tensor] = None

    # -- small caches (all keyed so that a changed tensor is never served stale) ----------------
    def register_static_map(self, w: torch.Tensor, compact=None, zero: bool = False) -> None:
        """A region map that lives in a STATIC device buffer whose contents are rewritten in place between calls (CUDA-graph
        replay): it is used exactly as given -- no re-layout copy, no content inspection -- together with the compact
        form ``(Wc, cols)`` (also a static buffer; the column list is baked into a captured graph) and the "all zero" verdict
        stated here.  ``w`` must already be in the padded device layout (``padded_region_map``)."""
        if not w.is_cuda or w.dtype != torch.float32 or padded_region_map(w) is not w:
            raise ValueError("a static region map must be a padded fp32 device tensor (see padded_region_map)")
        self._static_maps[w.data_ptr()] = (w, compact, bool(zero))

    def _static_entry(self, w: torch.Tensor):
        hit = self._static_maps.get(w.data_ptr()) if self._static_maps else None
        return hit if hit is not None and hit[0] is w else None

    def _device_map(self, w: torch.Tensor, device: torch.device) -> torch.Tensor:
        if self._static_entry(w) is not None:
            return w
        key = (w.data_ptr(), w._version, tuple(w.shape), w.dtype, str(device))
        hit = self._w_cache.get(key)
        if hit is not None and hit[0] is w:
            self._w_cache.move_to_end(key)
            return hit[1]
        dev = padded_region_map(w.to(device=device, dtype=torch.float32, non_blocking=False))
        # Regions switched off still reach this path with an all-zero map (reference encode_region_map_function.py:33-36,
        # :74): beta * 0 adds nothing, so such a map is recognised ONCE here (one readback per map and
        # generation, never inside a CUDA-graph capture) and its calls take plain SDPA -- no std pass at all.
        # The same readback yields the compact form (only the key columns that carry weights: the tokens of the region
        # phrases), which the tcgen05 pass 2 streams instead of the dense map (80 instead of 308 bytes per query row).
        zero, compact = False, None
        if dev.is_cuda and not torch.cuda.is_current_stream_capturing():
            compact = compact_region_map(dev)
            zero = self.skip_zero_maps and compact is not None and len(compact[1]) == 0
        self._w_zero[key] = zero
        self._w_compact[key] = compact
        self._w_cache[key] = (w, dev)  # keeps `w` alive, so data_ptr cannot be recycled under the key
        while len(self._w_cache) > self._max_cached_maps:
            old, _ = self._w_cache.popitem(last=False)
            self._w_zero.pop(old, None)
            self._w_compact.pop(old, None)
        return dev

    def _map_compact(self, w: torch.Tensor, device: torch.device):
        st = self._static_entry(w)
        if st is not None:
            return st[1]
        return self._w_compact.get((w.data_ptr(), w._version, tuple(w.shape), w.dtype, str(device)))

    def _map_is_zero(self, w: torch.Tensor, device: torch.device) -> bool:
        st = self._static_entry(w)
        if st is not None:
            return st[2]
        return self._w_zero.get((w.data_ptr(), w._version, tuple(w.shape), w.dtype, str(device)), False)

    def _check_weight_func(self, fn: Callable) -> None:
        ok = self._checked_funcs.get(id(fn))
        if ok is None or ok[0] is not fn:
            ok = (fn, _is_reference_weight_func(fn))
            if len(self._checked_funcs) > 64:
                self._checked_funcs.clear()
            self._checked_funcs[id(fn)] = ok
        if not ok[1]:
            raise NotImplementedError(
                "RegionAttnProcessor implements weight_func = w * sigma * qk.std() (reference app.py:1004); "
                "the callable passed in region_prompt['weight_func'] computes something else"
            )

    def _sigma_arg(self, sigma, device: torch.device):
        if isinstance(sigma, torch.Tensor) and sigma.is_cuda:
            if sigma.dtype == torch.float32 and sigma.device == device:
                return sigma
            c = self._sigma_cache
            if c is not None and c[0] is sigma and c[1] == sigma._version:
                return c[2]
            conv = sigma.detach().to(device=device, dtype=torch.float32)  # one tiny cast per step, no host sync
            self._sigma_cache = (sigma, sigma._version, conv)
            return conv
        return float(sigma)

    def register_static_kv(self, attn, ehs: torch.Tensor, key: torch.Tensor, value: torch.Tensor, images=None) -> None:
        """K / V projections of ``attn`` for the text embeddings that live in the STATIC buffer ``ehs`` (rewritten in
        place between CUDA-graph replays), themselves in static buffers the caller refreshes whenever it rewrites
        ``ehs``: calls with exactly this ``ehs`` use them instead of running ``to_k`` / ``to_v`` (so a captured graph
        holds no K/V GEMMs).  ``images``: optional ``{cols tuple: PreparedKV}`` of static K / V^T images, one per
        active-column list of the region maps the layer is used with."""
        self._kv_static[id(attn)] = (attn, ehs, key, value, dict(images or {}))

    def _project_kv(self, attn, ehs: torch.Tensor, args):
        st = self._kv_static.get(id(attn)) if self._kv_static else None
        if st is not None and st[0] is attn and st[1] is ehs:
            return st[2], st[3]
        if not self.cache_kv:
            return attn.to_k(ehs, *args), attn.to_v(ehs, *args)
        key = id(attn)
        hit = self._kv_cache.get(key)
        if hit is not None and hit[0] is ehs and hit[1] == ehs._version:
            return hit[2], hit[3]
        k, v = attn.to_k(ehs, *args), attn.to_v(ehs, *args)
        self._kv_cache[key] = (ehs, ehs._version, k, v)
        return k, v

    def _kv_image(self, attn, ehs: torch.Tensor, key: torch.Tensor, value: torch.Tensor, cols) -> PreparedKV:
        """The K / V^T image of this layer for the active-column list ``cols``: static (CUDA graph), cached (same text
        embeddings as last time) or rebuilt.  ``key`` / ``value`` are the [B, H, S, D] views of the projections."""
        cols = tuple(int(c) for c in cols)
        st = self._kv_static.get(id(attn)) if self._kv_static else None
        if st is not None and st[0] is attn and st[1] is ehs and cols in st[4]:
            return st[4][cols]
        if not self.cache_kv:
            return prepare_kv(key, value, cols)
        ck = (id(attn), cols)
        hit = self._img_cache.get(ck)
        if hit is not None and hit[0] is ehs and hit[1] == ehs._version and hit[2].B == key.shape[0] and hit[2].dtype == key.dtype:
            return hit[2]
        img = prepare_kv(key, value, cols, out=None if hit is None or hit[2].B != key.shape[0] or hit[2].dtype != key.dtype
                         else hit[2].image)
        if len(self._img_cache) > 256:
            self._img_cache.clear()
        self._img_cache[ck] = (ehs, ehs._version, img)
        return img

    def clear_caches(self) -> None:
        self._w_cache.clear()
        self._w_zero.clear()
        self._w_compact.clear()
        self._kv_cache.clear()
        self._img_cache.clear()
        self._sigma_cache = None

    def _score_scale(self, attn, head_dim: int) -> Optional[float]:
        """Factor on Q K^T on the region path.  The SDPA-style reference processor calls its function without ``scale``
        (attention_modify.py:479-481), i.e. 1/sqrt(head_dim) (:77) whatever ``attn.scale`` says: ``None``."""
        return None

    # -- processor protocol ---------------------------------------------------------------------
    def __call__(
        self,
        attn,
        hidden_states: torch.Tensor,
        encoder_hidden_states: Optional[torch.Tensor] = None,
        attention_mask: Optional[torch.Tensor] = None,
        temb: Optional[torch.Tensor] = None,
        scale: float = 1.0,
        region_prompt=None,
        ip_adapter_masks=None,
    ) -> torch.Tensor:
        return self._forward(attn, hidden_states, encoder_hidden_states, attention_mask, temb, region_prompt)

    def _region_mask(self, attention_mask, query):
        """The mask as the reference's SDPA-style region function treats it (attention_modify.py:84-89), probed on the
        unmodified module (tests/test_oracle_attention.py): a BOOL mask is never applied (:86-87 only rewrites the mask
        tensor), and a float mask is added IN PLACE into a [L, S] bias (:89) -- which raises for the 4-D
        [B, heads, -1, S] tensor this processor builds (:452), whatever its values.  Same behaviour here: None for a bool
        mask, the reference's RuntimeError for a float one.  (The baddbmm processor adds its mask: see the subclass.)"""
        if attention_mask is None or attention_mask.dtype == torch.bool:
            return None
        L, S = query.shape[-2], attention_mask.shape[-1]
        raise RuntimeError(f"output with shape [{L}, {S}] doesn't match the broadcast shape {list(attention_mask.shape[:2]) + [L, S]}"
                           " (the reference adds a float attention mask in place into its [L, S] bias, attention_modify.py:89;"
                           " use the baddbmm processor or region_attention(attn_mask=...) for an additive mask)")

    def _forward(self, attn, hidden_states, encoder_hidden_states, attention_mask, temb, region_prompt, ip_branch=None):
        """The processor body (reference attention_modify.py:425-503).  ``ip_branch(hidden, query, batch, head_dim)``,
        when given, adds the IP-Adapter image-prompt terms before the output projection (:640-682)."""
        residual = hidden_states
        img_sequence_length = hidden_states.shape[1]
        if getattr(attn, "spatial_norm", None) is not None:
            hidden_states = attn.spatial_norm(hidden_states, temb)

        input_ndim = hidden_states.ndim
        if input_ndim == 4:
            batch_size, channel, height, width = hidden_states.shape
            hidden_states = hidden_states.view(batch_size, channel, height * width).transpose(1, 2)

        is_xattn = False
        region_state = weight_func = sigma = None
        if encoder_hidden_states is not None and region_prompt is not None:
            is_xattn = True
            region_state = region_prompt["region_state"]
            weight_func = region_prompt["weight_func"]
            sigma = region_prompt["sigma"]

        batch_size, sequence_length, _ = (
            hidden_states.shape if encoder_hidden_states is None else encoder_hidden_states.shape
        )
        if attention_mask is not None:
            attention_mask = attn.prepare_attention_mask(attention_mask, sequence_length, batch_size)
            attention_mask = attention_mask.view(batch_size, attn.heads, -1, attention_mask.shape[-1])

        if getattr(attn, "group_norm", None) is not None:
            hidden_states = attn.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)

        args = ()  # diffusers >= 0.26 with the PEFT backend: Linear takes no scale argument (reference :457)
        query = attn.to_q(hidden_states, *args)

        cross = encoder_hidden_states is not None
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states
        elif getattr(attn, "norm_cross", None):
            encoder_hidden_states = attn.norm_encoder_hidden_states(encoder_hidden_states)

        if cross and is_xattn:
            key, value = self._project_kv(attn, encoder_hidden_states, args)
        else:
            key = attn.to_k(encoder_hidden_states, *args)
            value = attn.to_v(encoder_hidden_states, *args)

        inner_dim = key.shape[-1]
        head_dim = inner_dim // attn.heads
        query = query.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        key = key.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        value = value.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)

        if is_xattn and isinstance(region_state, dict):
            self._check_weight_func(weight_func)
            w = region_state[img_sequence_length]  # KeyError for an unknown resolution, like the reference (:481)
            w_dev = self._device_map(w, query.device)
            attention_mask = self._region_mask(attention_mask, query)  # what this processor's region branch does with a mask
            if self._map_is_zero(w, query.device):
                hidden_states = F.scaled_dot_product_attention(
                    query, key, value, attn_mask=attention_mask, dropout_p=0.0, is_causal=False,
                    scale=self._score_scale(attn, head_dim))
            else:
                compact = self._map_compact(w, query.device)
                if (attention_mask is None and compact is not None and query.dtype in (torch.float16, torch.bfloat16)
                        and prepared_supported(attn.heads, head_dim, key.shape[2], len(compact[1]))):
                    # SD-1.5 layers with 40-wide heads: tcgen05 kernels over a K / V^T image that is laid out once per
                    # generation (K, V are projections of the text embeddings: constant over the 25 steps)
                    kv = self._kv_image(attn, encoder_hidden_states, key, value, compact[1])
                    hidden_states = region_attention_prepared(
                        query, kv, compact, self._sigma_arg(sigma, query.device), workspace=self.workspace,
                        scale=self._score_scale(attn, head_dim))
                else:
                    hidden_states = region_attention(
                        query, key, value, w_dev,
                        self._sigma_arg(sigma, query.device),
                        attn_mask=attention_mask,
                        scale=self._score_scale(attn, head_dim),
                        workspace=self.workspace,
                        compact=compact,
                    )
        else:
            hidden_states = F.scaled_dot_product_attention(
                query, key, value, attn_mask=attention_mask, dropout_p=0.0, is_causal=False
            )

        hidden_states = hidden_states.transpose(1, 2).reshape(batch_size, -1, attn.heads * head_dim)
        hidden_states = hidden_states.to(query.dtype)
        if ip_branch is not None:
            hidden_states = ip_branch(hidden_states, query, batch_size, head_dim)

        hidden_states = attn.to_out[0](hidden_states, *args)
        hidden_states = attn.to_out[1](hidden_states)

        if input_ndim == 4:
            hidden_states = hidden_states.transpose(-1, -2).reshape(batch_size, channel, height, width)
        if getattr(attn, "residual_connection", False):
            hidden_states = hidden_states + residual
        hidden_states = hidden_states / getattr(attn, "rescale_output_factor", 1.0)
        return hidden_states


class RegionAttnProcessorBaddbmm(RegionAttnProcessor):
    """Drop-in for the reference's ``AttnProcessor`` (source/modules/attention_modify.py:107-207), the
    ``torch.baddbmm`` variant the reference falls back to when ``F.scaled_dot_product_attention`` is missing
    (source/app.py:479-481).  Its region branch (:164-175 via ``get_attention_scores`` :39-70) computes exactly the
    same scores, std, bias, softmax and PV as the SDPA-style function (bit-identical on CPU fp32), so
    the same two CUDA passes serve it.  The one difference that is part of the contract: its scores are
    ``alpha = attn.scale`` times Q K^T (``get_attention_scores`` :58-64), not 1/sqrt(head_dim); the two coincide for every
    SD-1.5 layer (``attn.scale = dim_head ** -0.5``) and differ for a module built with another scale."""

    def _score_scale(self, attn, head_dim: int) -> Optional[float]:
        return float(attn.scale)

    def _region_mask(self, attention_mask, query):
        """``get_attention_scores`` (:39-70) takes the prepared mask as the ``baddbmm`` input with beta = 1: a real additive
        mask M, a = M + scale Q K^T, and the weight_func's std is over that (:166).  A mask of another dtype than the query
        (a bool mask included) makes ``torch.baddbmm`` raise in the reference; here too."""
        if attention_mask is None:
            return None
        if attention_mask.dtype != query.dtype:
            raise RuntimeError(f"expected the attention mask in the query's dtype {query.dtype}, got {attention_mask.dtype} "
                               "(torch.baddbmm input, attention_modify.py:57-63)")
        return attention_mask


def ip_mask_downsample(mask: torch.Tensor, batch_size: int, num_queries: int, value_embed_dim: int) -> torch.Tensor:
    """Restatement of ``diffusers.image_processor.IPAdapterMaskProcessor.downsample`` (diffusers==0.27.2, third party,
    not vendored by the reference; called at attention_modify.py:671-673): bicubic resize of the [1, H, W] mask to the
    layer's latent grid (aspect ratio kept), flattened, repeated over the batch and the value channels."""
    import math

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
    if area < num_queries:  # aspect ratios that do not tile the grid exactly: pad with zeros / truncate, as upstream
        m = F.pad(m, (0, num_queries - m.shape[1]), value=0.0)
    if area > num_queries:
        m = m[:, :num_queries]
    return m.view(m.shape[0], m.shape[1], 1).repeat(1, 1, value_embed_dim)


class RegionIPAdapterAttnProcessor(torch.nn.Module):
    """Drop-in for the reference's ``IPAdapterAttnProcessor2_0`` (source/modules/attention_modify.py:506-700; installed
    by source/modules/ip_adapter.py:292): the text branch is the region-masked cross-attention of ``RegionAttnProcessor``
    (same CUDA path), each image-prompt branch is a plain attention of the same queries over the adapter's
    ``to_k_ip`` / ``to_v_ip`` projections, optionally gated by a spatial mask, added with its scale (:640-682).
    Same constructor, parameter names (state dicts load unchanged) and call signature."""

    _core_cls = RegionAttnProcessor

    def __init__(self, hidden_size, cross_attention_dim=None, num_tokens=(4,), scale=1.0, cache_kv: bool = False):
        super().__init__()
        self.hidden_size = hidden_size
        self.cross_attention_dim = cross_attention_dim
        if not isinstance(num_tokens, (tuple, list)):
            num_tokens = [num_tokens]
        self.num_tokens = num_tokens
        if not isinstance(scale, list):
            scale = [scale] * len(num_tokens)
        if len(scale) != len(num_tokens):
            raise ValueError("`scale` should be a list of integers with the same length as `num_tokens`.")
        self.scale = scale
        self.to_k_ip = torch.nn.ModuleList(
            [torch.nn.Linear(cross_attention_dim, hidden_size, bias=False) for _ in range(len(num_tokens))])
        self.to_v_ip = torch.nn.ModuleList(
            [torch.nn.Linear(cross_attention_dim, hidden_size, bias=False) for _ in range(len(num_tokens))])
        self._core = self._core_cls(cache_kv=cache_kv)

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0,
                 region_prompt=None, ip_adapter_masks=None):
        ip_hidden_states = None
        if encoder_hidden_states is not None:
            if isinstance(encoder_hidden_states, tuple):
                encoder_hidden_states, ip_hidden_states = encoder_hidden_states
            else:  # deprecated single-tensor form (:575-585): the last num_tokens[0] rows are the image tokens
                end_pos = encoder_hidden_states.shape[1] - self.num_tokens[0]
                encoder_hidden_states, ip_hidden_states = (
                    encoder_hidden_states[:, :end_pos, :], [encoder_hidden_states[:, end_pos:, :]])
        if ip_hidden_states is None:  # self-attention call: nothing to add
            return self._core._forward(attn, hidden_states, encoder_hidden_states, attention_mask, temb, region_prompt)

        if ip_adapter_masks is not None:
            if not isinstance(ip_adapter_masks, torch.Tensor) or ip_adapter_masks.ndim != 4:
                raise ValueError(" ip_adapter_mask should be a tensor with shape [num_ip_adapter, 1, height, width]."
                                 " Please use `IPAdapterMaskProcessor` to preprocess your mask")
            if len(ip_adapter_masks) != len(self.scale):
                raise ValueError(f"Number of ip_adapter_masks ({len(ip_adapter_masks)}) must match number of IP-Adapters "
                                 f"({len(self.scale)})")
        else:
            ip_adapter_masks = [None] * len(self.scale)

        def ip_branch(hidden, query, batch_size, head_dim):
            for cur, sc, to_k_ip, to_v_ip, mask in zip(ip_hidden_states, self.scale, self.to_k_ip, self.to_v_ip,
                                                       ip_adapter_masks):
                ip_key = to_k_ip(cur).view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
                ip_value = to_v_ip(cur).view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
                cur = F.scaled_dot_product_attention(query, ip_key, ip_value, attn_mask=None, dropout_p=0.0, is_causal=False)
                cur = cur.transpose(1, 2).reshape(batch_size, -1, attn.heads * head_dim).to(query.dtype)
                if mask is not None:
                    m = ip_mask_downsample(mask, batch_size, cur.shape[1], cur.shape[2])
                    cur = cur * m.to(dtype=query.dtype, device=query.device)
                hidden = hidden + sc * cur
            return hidden

        return self._core._forward(attn, hidden_states, encoder_hidden_states, attention_mask, temb, region_prompt, ip_branch)


class RegionIPAdapterAttnProcessorBaddbmm(RegionIPAdapterAttnProcessor):
    """Drop-in for the reference's ``IPAdapterAttnProcessor`` (source/modules/attention_modify.py:210-411), the
    ``torch.baddbmm`` twin of ``IPAdapterAttnProcessor2_0`` used when ``F.scaled_dot_product_attention`` is missing
    (source/modules/ip_adapter.py:292).  Same scores, std, bias, softmax and P V in the text branch and the same
    image-prompt terms, so the same CUDA path serves it (text-branch scores scaled by ``attn.scale`` like
    ``RegionAttnProcessorBaddbmm``)."""

    _core_cls = RegionAttnProcessorBaddbmm
