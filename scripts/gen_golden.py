"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE modules (build container only).

Run:  python scripts/gen_golden.py
Needs /root/reference (read-only).  The fixtures are small, committed, and travel to the GPU box where
the reference tree does not exist.  Everything is seeded; re-running reproduces the files bit for bit.

  attn_*.npz     inputs (fp16-representable, stored as float16) + fp32 output of the reference's
                 scaled_dot_product_attention_regionstate (attention_modify.py:74-103) with the
                 reference weight_func (app.py:1004), and the std it saw
  procm_baddbmm_*.npz the same with an additive attention mask ([B*heads, 1, S], the baddbmm input with beta = 1)
  proc_baddbmm_*.npz  module weights + inputs (float16) + fp32 output of the reference's ``AttnProcessor`` (the
                 torch.baddbmm variant, attention_modify.py:107-207) on the region path
  region_*.npz   inputs (uint8 maps, strengths, token ids) + the fp32 maps returned by the reference's
                 encode_region_map (encode_region_map_function.py:79-124, which calls cv2.resize)
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
weight_func = lambda w, sigma, qk: w * sigma * qk.std()  # noqa: E731  (reference app.py:1004)


def rect_map(h, w, r0, r1, c0, c1):
    m = np.full((h, w), 255, np.uint8)
    m[r0 : r1 + 1, c0 : c1 + 1] = 0
    return m


def ellipse_map(h, w, cy, cx, ry, rx):
    yy, xx = np.mgrid[0:h, 0:w]
    m = np.full((h, w), 255, np.uint8)
    m[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1] = 37
    return m


PROMPT_IDS = [49406, 320, 1611, 4919, 525, 518, 2465] + [49407] * 70  # "a girl sitting on the bridge" placeholders
VOCAB = {"A girl": [320, 1611], "bridge": [2465], "sitting": [4919], "the": [518], "dog": [1929], "on the": [525, 518]}


class StubTokenizer:
    model_max_length = 77

    def __call__(self, text, **kw):
        return SimpleNamespace(input_ids=list(VOCAB[text]))


def region_cases():
    # config 1/2: the two rectangles of Test_case 1
    yield "region_512_2rect", 512, 512, 2, {
        "A girl": {"map": rect_map(512, 512, 178, 325, 50, 304), "weight": 0.5, "mask_outsides": 0.0},
        "bridge": {"map": rect_map(512, 512, 317, 470, 52, 299), "weight": 0.7, "mask_outsides": 0.0},
    }
    # config 3 flavour: 4 regions, S' > 0, ellipses, a region that vanishes at coarse levels, S = 0, a missing phrase
    yield "region_768_4reg", 768, 768, 1, {
        "A girl": {"map": rect_map(768, 768, 30, 400, 40, 350), "weight": 0.5, "mask_outsides": 0.2},
        "bridge": {"map": ellipse_map(768, 768, 500, 520, 160, 210), "weight": 0.7, "mask_outsides": 0.0},
        "sitting": {"map": rect_map(768, 768, 700, 705, 700, 706), "weight": 0.4, "mask_outsides": 0.1},
        "on the": {"map": ellipse_map(768, 768, 200, 600, 120, 90), "weight": 1.0, "mask_outsides": 0.3},
        "the": {"map": rect_map(768, 768, 0, 767, 0, 383), "weight": 0.0, "mask_outsides": 0.25},
        "dog": {"map": rect_map(768, 768, 10, 20, 10, 20), "weight": 1.5, "mask_outsides": 0.5},
    }
    # non-square, all four levels still integer scale
    yield "region_512x768_overlap", 768, 512, 1, {
        "A girl": {"map": ellipse_map(512, 768, 256, 300, 200, 250), "weight": 1.25, "mask_outsides": 0.0},
        "bridge": {"map": ellipse_map(512, 768, 300, 420, 150, 300), "weight": 0.3, "mask_outsides": 0.05},
        "the": {"map": None, "weight": 2.0, "mask_outsides": 1.0},
    }


def gen_region():
    ref = ref_loader.encode_region_map_function()
    pipe = SimpleNamespace(tokenizer=StubTokenizer(), unet=SimpleNamespace(down_blocks=[0] * 4), vae_scale_factor=8,
                           do_classifier_free_guidance=True)
    ids = np.array([PROMPT_IDS])
    neg = np.array([[49406] + [49407] * 76])
    for name, width, height, n_img, state in region_cases():
        out = ref.encode_region_map(pipe, state, width, height, n_img, text_ids=[neg, ids])
        save = {"width": width, "height": height, "n_img": n_img, "ids": ids, "neg": neg,
                "phrases": np.array(list(state.keys()))}
        for i, (k, v) in enumerate(state.items()):
            save[f"map_{i}"] = v["map"] if v["map"] is not None else np.zeros((0, 0), np.uint8)
            save[f"weight_{i}"] = float(v["weight"])
            save[f"outside_{i}"] = float(v["mask_outsides"])
        for L, t in out.items():
            save[f"W_{L}"] = t.numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
        print(name, {L: tuple(t.shape) for L, t in out.items()})


def attn_cases():
    # name, B, H, L, D, S, Bw, sigma
    yield "attn_L64_D160", 1, 8, 64, 160, 77, 1, 14.6146
    yield "attn_L152_D80_tail", 2, 4, 144 + 8, 80, 77, 1, 3.9105  # L not a multiple of 16, Bw=1, one head group
    yield "attn_L256_D40", 2, 8, 256, 40, 77, 2, 0.0292
    yield "attn_L64_D64_S40", 4, 5, 64, 64, 40, 2, 7.0944


def gen_attn():
    ref = ref_loader.attention_modify()
    for seed, (name, B, H, L, D, S, Bw, sigma) in enumerate(attn_cases()):
        g = torch.Generator().manual_seed(100 + seed)
        q = (torch.randn(B, L, H * D, generator=g) * 1.3).half()
        k = (torch.randn(B, S, H * D, generator=g) * 1.1).half()
        k[:, 0] += 2.0  # a BOS-like sink column so the score mean is not ~0
        v = torch.randn(B, S, H * D, generator=g).half()
        W = torch.zeros(Bw, L, S)
        W[:, : L // 2, 1:3] = 0.5
        W[:, L // 3 :, 6 % S] += 0.7
        W[:, L // 4 : L // 2, 3] = -0.25
        W = W + 0.0
        view = lambda t: t.float().view(B, -1, H, D).transpose(1, 2)
        out = ref.scaled_dot_product_attention_regionstate(
            view(q), view(k), view(v), weight_func=weight_func, region_state=W.clone(), sigma=torch.tensor(sigma))
        std = (view(q) @ view(k).transpose(-2, -1) * (D**-0.5)).std()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), q=q.numpy(), k=k.numpy(), v=v.numpy(), W=W.numpy(),
                            sigma=np.float32(sigma), heads=H, out=out.transpose(1, 2).reshape(B, L, H * D).numpy(),
                            std=std.numpy())
        print(name, tuple(out.shape), float(std))


def load_numpy_weights(attn, seed):
    """fp16-representable weights from numpy's legacy generator (bit-stable across versions): the fixtures store only the
    seed.  tests/helpers.py holds the same function."""
    rng = np.random.RandomState(seed)
    with torch.no_grad():
        for _, prm in sorted(attn.named_parameters()):
            w = (rng.standard_normal(tuple(prm.shape)) * (1.0 / np.sqrt(prm.shape[-1]))).astype(np.float16)
            prm.copy_(torch.from_numpy(w.astype(np.float32)))


def gen_proc_baddbmm():
    """proc_baddbmm_*.npz: the reference's ``AttnProcessor`` (baddbmm variant, attention_modify.py:107-207) run on a small
    SD-1.5-shaped attention module: fp16-representable weights / inputs, fp32 output."""
    import torch.nn as nn

    ref = ref_loader.attention_modify()

    class Attn(nn.Module):
        upcast_attention = upcast_softmax = False

        def __init__(self, C, H, D):
            super().__init__()
            self.heads, self.scale = H, D**-0.5
            self.to_q, self.to_k, self.to_v = nn.Linear(C, H * D, bias=False), nn.Linear(768, H * D, bias=False), nn.Linear(768, H * D, bias=False)
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

    # procm_*: the same with an additive attention mask in the form diffusers' prepare_attention_mask hands over
    # ([B*heads, 1, S], :144): the baddbmm input with beta = 1 (:52-63) -- key biases per (batch, head), finite values
    for name, C, H, D, B, L, sigma, masked in (("proc_baddbmm_L256_D40", 320, 8, 40, 2, 256, 5.0, False),
                                               ("proc_baddbmm_L64_D160", 1280, 8, 160, 2, 64, 11.0, False),
                                               ("procm_baddbmm_L144_D80", 640, 8, 80, 2, 144, 7.0, True)):
        attn = Attn(C, H, D)
        seed = len(name) + L
        load_numpy_weights(attn, seed)
        g = torch.Generator().manual_seed(L)
        hs = torch.randn(B, L, C, generator=g).half().float()
        ctx = torch.randn(B, 77, 768, generator=g).half().float()
        W = torch.zeros(B, L, 77)
        W[:, : L // 2, 1:3] = 0.5
        W[:, L // 3 :, 6] += 0.7
        W[:, L // 4 : L // 2, 3] = -0.25
        rp = {"region_state": {L: W.clone()}, "sigma": torch.tensor(sigma), "weight_func": weight_func}
        extra = {}
        mask = None
        if masked:
            mask = torch.randn(B * H, 1, 77, generator=g) * 0.75
            mask[:, :, 60:] -= 4.0  # "padding" keys pushed down, still finite
            mask = mask.half().float()
            extra["mask"] = mask.half().numpy()
        with torch.no_grad():
            out = ref.AttnProcessor()(attn, hs, encoder_hidden_states=ctx, attention_mask=mask, region_prompt=rp)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), hs=hs.half().numpy(), ctx=ctx.half().numpy(), W=W.numpy(),
                            sigma=np.float32(sigma), heads=H, head_dim=D, weight_seed=seed, out=out.numpy(), **extra)
        print(name, tuple(out.shape))


if __name__ == "__main__":
    if not ref_loader.reference_available():
        sys.exit("reference tree not available: golden fixtures can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    gen_region()
    gen_attn()
    gen_proc_baddbmm()
