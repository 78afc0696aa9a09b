"""Functional entry points of the hot path: thin Python over the C ABI (include/dsc_b200.h).

``region_attention`` has the call shape of the reference's
``scaled_dot_product_attention_regionstate(query, key, value, ..., region_state=, sigma=)``
(reference source/modules/attention_modify.py:74-103) with ``weight_func`` fixed to the reference's
``w * sigma * qk.std()`` (reference source/app.py:1004).  PyTorch is used for device memory and the
current stream only; all arithmetic happens in libdsc_b200.so.  There is no CPU path.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Union

import torch

from . import _lib
from ._lib import check, lib

_I64x4 = ctypes.c_int64 * 4
_I64x3 = ctypes.c_int64 * 3
_DTYPES = {torch.float16: _lib.DTYPE_F16, torch.bfloat16: _lib.DTYPE_BF16}
_WORKSPACES: dict = {}


def workspace_bytes(B: int = 1, H: int = 1, L: int = 1, D: int = 40, S: int = 77) -> int:
    n = ctypes.c_size_t(0)
    check(lib.dsc_xattn_workspace_bytes(B, H, L, D, S, ctypes.byref(n)))
    return int(n.value)


def get_workspace(device: torch.device, nbytes: int = 0) -> torch.Tensor:
    """One zero-initialised workspace per (device, stream); calls on a stream are serialised, so all
    layers can share it.  Allocated through PyTorch's caching allocator (CUDA-graph friendly).  Long prompts
    (more than 80 keys) need room for their per-chunk outputs: the buffer grows to the largest request."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(workspace_bytes(), nbytes), dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


def _checked_workspace(workspace: Optional[torch.Tensor], device: torch.device, need: int) -> torch.Tensor:
    """The kernels always write the 64-byte header and their per-CTA partials: a caller-supplied workspace is checked
    against ``dsc_xattn_workspace_bytes`` for EVERY call (not only long prompts), plus dtype / device / alignment.  It
    must have been zero-filled once after allocation (the ticket word returns to 0 after each call)."""
    if workspace is None:
        return get_workspace(device, need)
    if workspace.dtype != torch.uint8 or not workspace.is_cuda or workspace.device != device or not workspace.is_contiguous():
        raise ValueError(f"workspace must be a contiguous uint8 tensor on {device}")
    if workspace.numel() < need:
        raise ValueError(f"workspace too small: {workspace.numel()} < {need} bytes (dsc_xattn_workspace_bytes)")
    if workspace.data_ptr() % 16 != 0:
        raise ValueError("workspace must be 16-byte aligned")
    return workspace


def _sigma_args(sigma, device: torch.device):
    """(device pointer | None, host value, keep-alive).  A device sigma is passed by address (no host sync).  A host
    sigma is passed by value -- which a CUDA-graph capture would freeze into the kernel parameters, so every replay
    would reuse the capture-time sigma: refused while the stream is capturing."""
    if isinstance(sigma, torch.Tensor) and sigma.is_cuda:
        if sigma.dtype != torch.float32 or sigma.device != device:
            sigma = sigma.to(device=device, dtype=torch.float32)
        keep = sigma.reshape(-1)
        return keep.data_ptr(), 0.0, keep
    if torch.cuda.is_current_stream_capturing():
        raise RuntimeError("sigma must be a CUDA tensor while a CUDA graph is being captured: a host value would be "
                           "frozen into the captured kernel parameters and every replay would reuse it")
    return None, float(sigma), None


def _as_bhxd(x: torch.Tensor, name: str) -> torch.Tensor:
    """Return x ([B,H,rows,D]) in the layout the kernels stream: a view of [B, rows, H*D]."""
    if x.dim() != 4:
        raise ValueError(f"{name} must be [B, H, rows, D], got {tuple(x.shape)}")
    B, H, R, D = x.shape
    s = x.stride()
    ok = s[3] == 1 and s[1] == D and s[2] % 8 == 0 and s[0] % 8 == 0 and x.data_ptr() % 16 == 0
    if not ok:  # e.g. a contiguous [B,H,rows,D] tensor: re-lay it out once (plumbing, not arithmetic)
        x = x.transpose(1, 2).contiguous().transpose(1, 2)
    return x


def _stream_ptr(device: torch.device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_inputs(query, key):
    if not query.is_cuda:
        raise RuntimeError("diffusionspatialcontrol_b200 has no CPU path: tensors must live on a CUDA device")
    if query.dtype not in _DTYPES:
        raise TypeError(f"query dtype {query.dtype} not supported (float16 / bfloat16)")
    if key.dtype != query.dtype:
        raise TypeError("query/key/value must share one dtype")


@_lib.nvtx("xattn_stats")
def score_stats(query: torch.Tensor, key: torch.Tensor, scale: Optional[float] = None,
                workspace: Optional[torch.Tensor] = None, attn_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Pass 1 alone.  Returns the workspace (uint8); bytes 8..12 hold the fp32 unbiased std of
    ``scale * Q K^T (+ attn_mask)`` over the whole call, see ``read_stats``.  Asynchronous.  ``attn_mask``: additive float
    mask broadcastable to [B, H, L, S] (handed over as the dense fp32 tensor ``dsc_xattn_stats`` takes)."""
    _check_inputs(query, key)
    q, k = _as_bhxd(query, "query"), _as_bhxd(key, "key")
    B, H, L, D = q.shape
    S = k.shape[2]
    scale = 1.0 / math.sqrt(D) if scale is None else float(scale)
    ws = _checked_workspace(workspace, q.device, workspace_bytes(B, H, L, D, S))
    dense = None
    if attn_mask is not None:
        dense = _mask_arg(attn_mask, B, H, L, S, q.device)[0].expand(B, H, L, S).contiguous()
    with torch.cuda.device(q.device):
        check(lib.dsc_xattn_stats(q.data_ptr(), k.data_ptr(), _I64x4(*q.stride()), _I64x4(*k.stride()),
                                  None if dense is None else dense.data_ptr(),
                                  B, H, L, D, S, scale, _DTYPES[q.dtype], ws.data_ptr(), _stream_ptr(q.device)))
    return ws


def read_stats(workspace: torch.Tensor) -> dict:
    """Synchronising debug/test helper: decode the dsc_xattn_stats_t header."""
    raw = workspace[:48].cpu().numpy().tobytes()
    import struct

    ticket, n_part, std, mean, s, ss, n = struct.unpack("<IIffddd", raw[:40])
    return {"ticket": ticket, "n_partials": n_part, "std": std, "mean": mean, "sum": s, "sumsq": ss, "n": n}


MAX_KEYS = 80  # DSC_MAX_KEYS: keys per kernel pass
MAX_KEYS_TOTAL = 480  # DSC_MAX_KEYS_TOTAL: long prompts (77*k tokens) run as chunks of 80 keys


def _padded_pitch(S: int) -> int:
    return MAX_KEYS * ((S + MAX_KEYS - 1) // MAX_KEYS)


def _region_layout_ok(W: torch.Tensor) -> bool:
    """[B', L, S] fp32 with unit key stride and batch stride L * pitch; row pitch in [S, 80] (dense or padded), or,
    for long prompts, any multiple of 4 floats with a 16-byte aligned base."""
    Bw, L, S = W.shape
    pitch = W.stride(1)
    if W.stride(2) != 1 or pitch < S or (Bw > 1 and W.stride(0) != L * pitch):
        return False
    if S <= MAX_KEYS:
        return pitch <= MAX_KEYS and W.data_ptr() % 4 == 0
    return pitch % 4 == 0 and W.data_ptr() % 16 == 0


def padded_region_map(W: torch.Tensor) -> torch.Tensor:
    """The same [B', L, S] values in the fast device layout: rows 80 floats apart (a multiple of 80 for long
    prompts), i.e. 16-byte aligned rows that the kernels fetch with TMA boxes and read as 128-bit words.  Returns a
    [B', L, S] VIEW of a zero-padded buffer, so it still indexes, compares and prints like the reference's tensor."""
    Bw, L, S = W.shape
    if S > MAX_KEYS_TOTAL:
        raise NotImplementedError(f"at most {MAX_KEYS_TOTAL} keys are supported, got {S}")
    pitch = _padded_pitch(S)
    if W.stride(1) == pitch and _region_layout_ok(W):
        return W
    buf = torch.zeros((Bw, L, pitch), dtype=torch.float32, device=W.device)
    buf[:, :, :S] = W
    return buf[:, :, :S]


MAX_COMPACT_COLS = 16  # DSC_MAX_COMPACT_COLS
COMPACT_PITCH = 20     # DSC_COMPACT_PITCH
AUTO_COMPACT = False   # tests: derive the compact form inside region_attention (costs a device->host readback per call)


def compact_region_map(W: torch.Tensor):
    """Compact form of a region map: with region prompts only the few tokens of the region phrases carry weights
    (reference encode_region_map_function.py:57-63), so W[B', L, S] is non-zero in a handful of key columns.  Returns
    ``(Wc, cols)`` -- fp32 [B', L, 20] holding those columns (ascending, zero-padded to 16 + 4) and their indices -- or
    ``None`` when more than 16 columns are in use.  One device->host readback: call it once per map (the processor
    does, at upload), never per attention call."""
    Bw, L, S = W.shape
    cols = torch.nonzero((W != 0).reshape(-1, S).any(dim=0)).flatten().tolist()
    if len(cols) > MAX_COMPACT_COLS:
        return None
    Wc = torch.zeros((Bw, L, COMPACT_PITCH), dtype=torch.float32, device=W.device)
    if cols:
        Wc[:, :, : len(cols)] = W[:, :, cols]
    return Wc, cols


@_lib.nvtx("xattn_call")
def region_attention(
    query: torch.Tensor,  # [B, H, L, D]
    key: torch.Tensor,  # [B, H, S, D]
    value: torch.Tensor,  # [B, H, S, D]
    region_state: torch.Tensor,  # fp32 [B', L, S] on the same device
    sigma: Union[float, torch.Tensor],
    attn_mask: Optional[torch.Tensor] = None,
    scale: Optional[float] = None,
    workspace: Optional[torch.Tensor] = None,
    compact=None,  # optional (Wc, cols) from compact_region_map(region_state)
) -> torch.Tensor:
    """softmax(a + sigma*std(a)*W) V with a = scale*QK^T (+ M)  ->  [B, H, L, D] (a view of a fresh [B, L, H*D]).

    ``attn_mask``: optional ADDITIVE float mask M, any shape that broadcasts to [B, H, L, S] (``[S]``, ``[L, S]``,
    ``[B, H, 1, S]`` ...).  The std is taken over the masked scores, as the reference's weight_func sees them
    (attention_modify.py:90-95; baddbmm variant :39-70).  A BOOL mask is ignored: the reference function never applies
    one (:86-87 rewrites the mask tensor and adds nothing) -- its output with a bool mask equals its output without.
    This function accepts every broadcastable float mask; which masks reach it is the processors' business (the reference's
    SDPA-style processor raises for its own 4-D mask, the baddbmm one adds it: see ``RegionAttnProcessor._region_mask``).
    Masked calls run as two launches on the mma.sync kernels (``dsc_xattn_call_masked``)."""
    _check_inputs(query, key)
    if attn_mask is not None and attn_mask.dtype == torch.bool:
        attn_mask = None  # reference :86-87: never added to the bias
    q, k, v = _as_bhxd(query, "query"), _as_bhxd(key, "key"), _as_bhxd(value, "value")
    B, H, L, D = q.shape
    S = k.shape[2]
    W = region_state
    if W.dim() != 3 or W.shape[1] != L or W.shape[2] != S:
        raise ValueError(f"region_state must be [B', L={L}, S={S}], got {tuple(W.shape)}")
    if (B * H) % W.shape[0] != 0:
        raise ValueError(f"region_state batch {W.shape[0]} does not divide B*H={B * H}")  # reference: shape error at :97
    if B % W.shape[0] != 0:
        raise NotImplementedError(f"region_state batch {W.shape[0]} must divide the attention batch {B}")
    if W.dtype != torch.float32 or W.device != q.device or not _region_layout_ok(W):
        W = padded_region_map(W.to(device=q.device, dtype=torch.float32))
    scale = 1.0 / math.sqrt(D) if scale is None else float(scale)
    ws = _checked_workspace(workspace, q.device, workspace_bytes(B, H, L, D, S))
    out = torch.empty((B, L, H * D), dtype=q.dtype, device=q.device)
    sigma_ptr, sigma_host, _sigma_keepalive = _sigma_args(sigma, q.device)  # never .item() a device sigma

    dt = _DTYPES[q.dtype]
    st = _stream_ptr(q.device)
    qs, ks, vs = _I64x4(*q.stride()), _I64x4(*k.stride()), _I64x4(*v.stride())
    with torch.cuda.device(q.device):
        if compact is None and AUTO_COMPACT:
            compact = compact_region_map(W)
        wc_ptr, n_act, cols_arr = None, 0, None
        if compact is not None and len(compact[1]) > 0:
            Wc, cols = compact
            if Wc.shape != (W.shape[0], L, COMPACT_PITCH) or Wc.dtype != torch.float32 or not Wc.is_contiguous() \
                    or Wc.device != q.device:
                raise ValueError("compact region map must be a contiguous fp32 [B', L, 20] tensor on the query's device")
            wc_ptr, n_act = Wc.data_ptr(), len(cols)
            cols_arr = (ctypes.c_int32 * n_act)(*cols)
        if attn_mask is not None:
            m, m_str = _mask_arg(attn_mask, B, H, L, S, q.device)
            check(lib.dsc_xattn_call_masked(q.data_ptr(), k.data_ptr(), v.data_ptr(), qs, ks, vs, W.data_ptr(), W.shape[0],
                                            W.stride(1), m.data_ptr(), _I64x3(*m_str), sigma_ptr, sigma_host, ws.data_ptr(),
                                            out.data_ptr(), _I64x3(*out.stride()), B, H, L, D, S, scale, dt, st))
            return out.view(B, L, H, D).transpose(1, 2)
        check(lib.dsc_xattn_call_cw(q.data_ptr(), k.data_ptr(), v.data_ptr(), qs, ks, vs, W.data_ptr(), W.shape[0],
                                    W.stride(1), wc_ptr, n_act, cols_arr, sigma_ptr, sigma_host, ws.data_ptr(),
                                    out.data_ptr(), _I64x3(*out.stride()), B, H, L, D, S, scale, dt, st))
    return out.view(B, L, H, D).transpose(1, 2)


def _mask_arg(attn_mask: torch.Tensor, B: int, H: int, L: int, S: int, device):
    """fp32 device copy of an additive mask that broadcasts to [B, H, L, S] + its (B, H, L) element strides (0 = broadcast)."""
    if not attn_mask.is_floating_point():
        raise TypeError(f"attn_mask must be a float (additive) or bool tensor, got {attn_mask.dtype}")
    if attn_mask.dim() > 4:
        raise ValueError(f"attn_mask must broadcast to [B, H, L, S], got {tuple(attn_mask.shape)}")
    m = attn_mask.reshape((1,) * (4 - attn_mask.dim()) + tuple(attn_mask.shape))
    for have, want, name in zip(m.shape, (B, H, L, S), "BHLS"):
        if have not in (1, want):
            raise ValueError(f"attn_mask {tuple(attn_mask.shape)} does not broadcast to [B={B}, H={H}, L={L}, S={S}] (dim {name})")
    if m.shape[3] == 1:
        m = m.expand(m.shape[0], m.shape[1], m.shape[2], S)
    m = m.to(device=device, dtype=torch.float32).contiguous()
    strides = [0 if m.shape[d] == 1 else m.stride(d) for d in range(3)]
    return m, strides


# ---- prepared K / V (the fast form for the SD-1.5 shapes) ---------------------------------------------------------------
class PreparedKV:
    """K and V of one cross-attention layer laid out ONCE as the shared-memory image the tcgen05 kernels multiply from
    (``dsc_xattn_prepare_kv``).  They are projections of the text embeddings and do not change during a generation
    (reference attention_modify.py:465-466 recomputes the same ``to_k`` / ``to_v`` on each of the 25 steps).  The image
    bakes in the active-column list of the compact region map it is used with."""

    __slots__ = ("image", "cols", "B", "H", "D", "S", "dtype")

    def __init__(self, image, cols, B, H, D, S, dtype):
        self.image, self.cols, self.B, self.H, self.D, self.S, self.dtype = image, tuple(cols), B, H, D, S, dtype


def prepared_supported(H: int, D: int, S: int, n_active: int) -> bool:
    return 1 <= n_active <= MAX_COMPACT_COLS and bool(lib.dsc_xattn_prepared_supported(H, D, S))


def kv_image_bytes(B: int, H: int, D: int, S: int) -> int:
    n = ctypes.c_size_t(0)
    check(lib.dsc_xattn_kv_image_bytes(B, H, D, S, ctypes.byref(n)))
    return int(n.value)


@_lib.nvtx("xattn_prepare_kv")
def prepare_kv(key: torch.Tensor, value: torch.Tensor, cols, out: Optional[torch.Tensor] = None) -> PreparedKV:
    """key, value: [B, H, S, D] (views of the [B, S, H*D] projections); cols: ascending active key columns of the compact
    region map (``compact_region_map(W)[1]``).  ``out``: optional uint8 device buffer to (re)fill in place (static buffers
    of a captured CUDA graph).  Asynchronous on the current stream."""
    _check_inputs(key, value)
    k, v = _as_bhxd(key, "key"), _as_bhxd(value, "value")
    B, H, S, D = k.shape
    cols = [int(c) for c in cols]
    nbytes = kv_image_bytes(B, H, D, S)
    if out is None:
        out = torch.empty(nbytes, dtype=torch.uint8, device=k.device)
    elif out.dtype != torch.uint8 or out.device != k.device or out.numel() < nbytes or not out.is_contiguous():
        raise ValueError(f"kv image buffer must be a contiguous uint8 tensor of >= {nbytes} bytes on {k.device}")
    cols_arr = (ctypes.c_int32 * max(1, len(cols)))(*cols)
    with torch.cuda.device(k.device):
        check(lib.dsc_xattn_prepare_kv(k.data_ptr(), v.data_ptr(), _I64x4(*k.stride()), _I64x4(*v.stride()), len(cols),
                                       cols_arr, B, H, D, S, _DTYPES[k.dtype], out.data_ptr(), _stream_ptr(k.device)))
    return PreparedKV(out, cols, B, H, D, S, k.dtype)


@_lib.nvtx("xattn_call_prepared")
def region_attention_prepared(
    query: torch.Tensor,  # [B, H, L, D]
    kv: PreparedKV,
    compact,  # (Wc fp32 [B', L, 20], cols) from compact_region_map; cols must equal kv.cols
    sigma: Union[float, torch.Tensor],
    scale: Optional[float] = None,
    workspace: Optional[torch.Tensor] = None,
    passes: int = _lib.PASS_BOTH,
    out: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """``region_attention`` on a prepared K / V image: same result, K / V^T fetched with one bulk copy per CTA."""
    if not query.is_cuda or query.dtype not in _DTYPES:
        raise TypeError("query must be a float16 / bfloat16 CUDA tensor")
    q = _as_bhxd(query, "query")
    B, H, L, D = q.shape
    Wc, cols = compact
    if tuple(int(c) for c in cols) != kv.cols:
        raise ValueError("the K/V image was prepared for another active-column list")
    if (B, H, D) != (kv.B, kv.H, kv.D) or q.dtype != kv.dtype or kv.image.device != q.device:
        raise ValueError("query does not match the prepared K/V image (batch / heads / head dim / dtype / device)")
    if Wc.dim() != 3 or Wc.shape[1:] != (L, COMPACT_PITCH) or Wc.dtype != torch.float32 or not Wc.is_contiguous() \
            or Wc.device != q.device or B % Wc.shape[0] != 0:
        raise ValueError("compact region map must be a contiguous fp32 [B', L, 20] tensor on the query's device, B' | B")
    scale = 1.0 / math.sqrt(D) if scale is None else float(scale)
    ws = _checked_workspace(workspace, q.device, workspace_bytes(B, H, L, D, kv.S))
    if out is None:
        out = torch.empty((B, L, H * D), dtype=q.dtype, device=q.device)
    sigma_ptr, sigma_host, _keep = _sigma_args(sigma, q.device)
    with torch.cuda.device(q.device):
        check(lib.dsc_xattn_call_prepared(q.data_ptr(), _I64x4(*q.stride()), kv.image.data_ptr(), Wc.data_ptr(), Wc.shape[0],
                                          len(kv.cols), sigma_ptr, sigma_host, ws.data_ptr(), out.data_ptr(),
                                          _I64x3(*out.stride()), B, H, L, D, kv.S, scale, _DTYPES[q.dtype], passes,
                                          _stream_ptr(q.device)))
    return out.view(B, L, H, D).transpose(1, 2)
