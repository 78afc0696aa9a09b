"""Shared builders for the parity tests (synthetic inputs of SURVEY.md 8d)."""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

PROMPT_IDS = [49406, 320, 1611, 4919, 525, 518, 2465] + [49407] * 70
NEG_IDS = [49406] + [49407] * 76
VOCAB = {"A girl": [320, 1611], "bridge": [2465], "sitting": [4919], "the": [518], "dog": [1929], "on the": [525, 518]}
weight_func = lambda w, sigma, qk: w * sigma * qk.std()  # noqa: E731  reference app.py:1004


class StubTokenizer:
    model_max_length = 77

    def __call__(self, text, **kw):
        return SimpleNamespace(input_ids=list(VOCAB[text]))


def rect_map(h, w, r0, r1, c0, c1):
    m = np.full((h, w), 255, np.uint8)
    m[r0 : r1 + 1, c0 : c1 + 1] = 0
    return m


def ellipse_map(h, w, cy, cx, ry, rx, val=37):
    yy, xx = np.mgrid[0:h, 0:w]
    m = np.full((h, w), 255, np.uint8)
    m[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1] = val
    return m


def two_rect_state(h=512, w=512):
    """The two (overlapping) rectangles of Figure/.../Test_case 1/Region map.png, scaled to (h, w)."""
    sy, sx = h / 512, w / 512
    r = lambda a, s: int(round(a * s))
    return {
        "A girl": {"map": rect_map(h, w, r(178, sy), r(325, sy), r(50, sx), r(304, sx)), "weight": 0.5, "mask_outsides": 0.0},
        "bridge": {"map": rect_map(h, w, r(317, sy), r(470, sy), r(52, sx), r(299, sx)), "weight": 0.7, "mask_outsides": 0.0},
    }


def state_from_golden(z):
    state = {}
    for i, ph in enumerate(z["phrases"].tolist()):
        m = z[f"map_{i}"]
        state[ph] = {"map": None if m.size == 0 else m, "weight": float(z[f"weight_{i}"]),
                     "mask_outsides": float(z[f"outside_{i}"])}
    return state


def synthetic_w(Bw, L, S, dtype=torch.float32):
    W = torch.zeros(Bw, L, S, dtype=dtype)
    W[:, : L // 2, 1 : min(3, S)] = 0.5
    W[:, L // 3 :, 6 % S] += 0.7
    W[:, L // 4 : L // 2, 3 % S] = -0.25
    return W


def make_qkv(B, H, L, D, S, seed, dtype=torch.float16, device="cpu", sink=2.0):
    """[B, rows, H*D] projections (low-precision representable) + their [B,H,rows,D] views."""
    g = torch.Generator().manual_seed(seed)
    q = (torch.randn(B, L, H * D, generator=g) * 1.3).to(dtype)
    k = (torch.randn(B, S, H * D, generator=g) * 1.1).to(dtype)
    k[:, 0] += sink
    v = torch.randn(B, S, H * D, generator=g).to(dtype)
    q, k, v = q.to(device), k.to(device), v.to(device)
    view = lambda t: t.view(B, -1, H, D).transpose(1, 2)
    return view(q), view(k), view(v)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())



def load_numpy_weights(attn, seed):
    """Same as scripts/gen_golden.py::load_numpy_weights: fp16-representable weights from numpy's legacy generator
    (bit-stable across versions), so the proc_* fixtures store only the seed."""
    rng = np.random.RandomState(seed)
    with torch.no_grad():
        for _, prm in sorted(attn.named_parameters()):
            w = (rng.standard_normal(tuple(prm.shape)) * (1.0 / np.sqrt(prm.shape[-1]))).astype(np.float16)
            prm.copy_(torch.from_numpy(w.astype(np.float32)))


class AttnModule(torch.nn.Module):
    """Duck-typed diffusers ``Attention`` with everything both reference processors touch
    (attention_modify.py:107-207 needs head_to_batch_dim / batch_to_head_dim / prepare_attention_mask / scale)."""

    upcast_attention = False
    upcast_softmax = False

    def __init__(self, C, H, D, ctx=768):
        super().__init__()
        nn = torch.nn
        self.heads, self.scale = H, D**-0.5
        self.to_q, self.to_k, self.to_v = nn.Linear(C, H * D, bias=False), nn.Linear(ctx, H * D, bias=False), nn.Linear(ctx, H * D, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(H * D, C), nn.Dropout(0.0)])
        self.spatial_norm = self.group_norm = self.norm_cross = None
        self.residual_connection, self.rescale_output_factor = False, 1.0

    def head_to_batch_dim(self, t, out_dim=3):
        b, n, c = t.shape
        t = t.reshape(b, n, self.heads, c // self.heads).permute(0, 2, 1, 3)
        return t.reshape(b * self.heads, n, c // self.heads) if out_dim == 3 else t

    def batch_to_head_dim(self, t):
        bh, n, d = t.shape
        return t.reshape(bh // self.heads, self.heads, n, d).permute(0, 2, 1, 3).reshape(bh // self.heads, n, d * self.heads)

    def prepare_attention_mask(self, m, *_a, **_k):
        return m

    def get_attention_scores(self, query, key, attention_mask=None):
        """diffusers ``Attention.get_attention_scores`` (third party; the reference's non-region branch calls it, :188):
        softmax of ``scale * Q K^T``."""
        empty = torch.empty(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype, device=query.device)
        scores = torch.baddbmm(empty, query, key.transpose(-1, -2), beta=0, alpha=self.scale)
        return scores.softmax(dim=-1).to(query.dtype)


def baddbmm_fixture(path):
    """(attn fp32 module, hs, ctx, region_prompt, reference output) of a tests/golden/proc_baddbmm_*.npz fixture."""
    z = np.load(path)
    H, D = int(z["heads"]), int(z["head_dim"])
    hs, ctx = torch.from_numpy(z["hs"]).float(), torch.from_numpy(z["ctx"]).float()
    attn = AttnModule(hs.shape[-1], H, D)
    load_numpy_weights(attn, int(z["weight_seed"]))
    rp = {"region_state": {hs.shape[1]: torch.from_numpy(z["W"]).clone()}, "sigma": torch.tensor(float(z["sigma"])),
          "weight_func": weight_func}
    return attn, hs, ctx, rp, torch.from_numpy(z["out"])


def baddbmm_mask_fixture(path):
    """baddbmm_fixture + the additive attention mask ([B*heads, 1, S] fp32) of a tests/golden/procm_baddbmm_*.npz fixture."""
    return (*baddbmm_fixture(path), torch.from_numpy(np.load(path)["mask"]).float())
