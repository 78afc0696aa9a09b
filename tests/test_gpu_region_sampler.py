"""Companion kernels: region-map builder (bit-exact) and the fused DPM++ 2M step."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import region_map as orm
from oracle import sampler as osm

from .helpers import NEG_IDS, PROMPT_IDS, StubTokenizer, ellipse_map, rect_map, state_from_golden, two_rect_state

pytestmark = pytest.mark.gpu
TOK = lambda phrase: StubTokenizer()(phrase).input_ids  # noqa: E731


def _pipe():
    from types import SimpleNamespace

    return SimpleNamespace(tokenizer=StubTokenizer(), unet=SimpleNamespace(down_blocks=[0] * 4), vae_scale_factor=8,
                           do_classifier_free_guidance=True)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "region_*.npz"))))
def test_region_maps_bit_exact_vs_reference_golden(path):
    from diffusionspatialcontrol_b200 import encode_region_map

    z = np.load(path, allow_pickle=False)
    out = encode_region_map(_pipe(), state_from_golden(z), int(z["width"]), int(z["height"]), int(z["n_img"]),
                            text_ids=[z["neg"], z["ids"]])
    keys = sorted(int(k[2:]) for k in z.files if k.startswith("W_"))
    assert sorted(out.keys()) == keys
    for L in keys:
        assert out[L].is_cuda and out[L].dtype == torch.float32
        assert torch.equal(out[L].cpu(), torch.from_numpy(z[f"W_{L}"])), L


@pytest.mark.parametrize("H,W", [(512, 512), (768, 768), (512, 768), (1088, 1920), (500, 500), (520, 776), (64, 64)])
def test_downsample_bit_exact_vs_oracle_and_cv2(H, W):
    """Random / tie-heavy / blob masks at every level; integer scales additionally against cv2 itself."""
    cv2 = pytest.importorskip("cv2")
    from diffusionspatialcontrol_b200.region_map import downsample_regions

    rng = np.random.default_rng(H + W)
    yy, xx = np.mgrid[0:H, 0:W]
    masks01 = np.stack([
        (rng.random((H, W)) < 0.5), (rng.random((H, W)) < 0.1), ((yy + xx) % 2 == 1), ((yy // 3 + xx // 5) % 2 == 1),
        ((yy - H * 0.4) / (H * 0.3)) ** 2 + ((xx - W * 0.55) / (W * 0.25)) ** 2 < 1, np.zeros((H, W), bool), np.ones((H, W), bool),
    ]).astype(np.uint8)
    maps = np.where(masks01 == 1, rng.integers(0, 255, masks01.shape), 255).astype(np.uint8)  # any value < 255 is "inside"
    dev = torch.from_numpy(maps).cuda()
    for sr in (8, 16, 32, 64):
        w_r, h_r = -(-W // sr), -(-H // sr)
        ds, any_set = downsample_regions(dev, w_r, h_r)
        ds = ds.cpu().numpy().reshape(len(maps), h_r, w_r)
        for r in range(len(maps)):
            want = orm.cubic_resize_binary(masks01[r], w_r, h_r)
            assert np.array_equal(ds[r], want), (sr, r)
            assert int(any_set[r]) == int(want.max())
            if W % sr == 0 and H % sr == 0:
                assert np.array_equal(ds[r], cv2.resize(masks01[r], (w_r, h_r), interpolation=cv2.INTER_CUBIC))
    same, _ = downsample_regions(dev, W, H)  # identity size: cv2 copies
    assert np.array_equal(same.cpu().numpy().reshape(masks01.shape), masks01)


def test_region_builder_quirks_and_live_oracle():
    from diffusionspatialcontrol_b200 import encode_region_map

    ids, neg = np.array([PROMPT_IDS]), np.array([NEG_IDS])
    state = two_rect_state(512, 512)
    state["on the"] = {"map": ellipse_map(512, 512, 100, 400, 60, 80), "weight": 1.1, "mask_outsides": 0.4}
    state["the"] = {"map": rect_map(512, 512, 90, 130, 380, 420), "weight": 0.9, "mask_outsides": 0.15}
    state["bridge"]["map"] = rect_map(256, 384, 100, 101, 100, 102)  # another source size, vanishes at coarse levels
    state["sitting"] = {"map": rect_map(512, 512, 0, 255, 0, 255), "weight": 0.0, "mask_outsides": 0.3}
    state["dog"] = {"map": rect_map(512, 512, 0, 9, 0, 9), "weight": 1.0, "mask_outsides": 1.0}  # not in the prompt
    got = encode_region_map(_pipe(), state, 512, 512, 3, text_ids=[neg, ids])
    want = orm.encode_region_map(state, TOK, 512, 512, 3, text_ids=[neg, ids])
    assert got.keys() == want.keys()
    for L in want:
        assert torch.equal(got[L].cpu(), want[L]), L
    off = encode_region_map(_pipe(), None, 512, 512, 2, text_ids=[neg, ids])
    assert all(t.shape == (4, L, 77) and not t.any() for L, t in off.items())


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_fused_step_matches_sampler_oracle(dtype):
    """25 fused steps against the fp64 k-diffusion restatement on a toy eps model."""
    from diffusionspatialcontrol_b200.sampler import KarrasSchedule, dpmpp2m_step

    torch.manual_seed(0)
    n, shape, g = 3, (3, 4, 16, 16), 7.5
    sched = KarrasSchedule(25)
    sig = sched.sigma_list()
    noise = torch.randn(shape, dtype=torch.float64)
    A = torch.randn(16, 16, dtype=torch.float64) * 0.05

    def eps_fn(x_in, sigma):  # x_in: [2n,...] already scaled by c_in
        e = torch.tanh(x_in @ A)
        e[n:] = e[n:] * 1.1 + 0.05
        return e

    want = osm.txt2img_latents(lambda xi, s: eps_fn(xi, s), noise, steps=25, guidance=g)
    x = (noise * (sig[0] ** 2 + 1) ** 0.5).float().cuda().contiguous()
    den_prev = torch.zeros_like(x)
    unet_in = torch.cat([x, x]).mul(sched.c_in(0)).to(dtype).contiguous()
    nxt = torch.empty_like(unet_in)
    for i in range(25):
        eps = eps_fn(unet_in.double().cpu(), sig[i]).to(dtype).cuda().contiguous()
        dpmpp2m_step(x, eps, den_prev, None if i == 24 else nxt, sig[i - 1] if i else 0.0, sig[i], sig[i + 1], g, first=(i == 0))
        unet_in, nxt = nxt, unet_in
    cos = torch.nn.functional.cosine_similarity(x.double().cpu().flatten(), want.flatten(), dim=0)
    tol = 2e-2 if dtype == torch.float16 else 1e-1
    assert cos > 0.9999 and (x.double().cpu() - want).abs().max() <= tol * want.abs().max()


def test_fused_step_exact_arithmetic_single_step():
    from diffusionspatialcontrol_b200.sampler import dpmpp2m_step
    import math

    torch.manual_seed(1)
    n = 1000
    x = torch.randn(n, device="cuda")
    den_prev = torch.randn(n, device="cuda")
    eps = torch.randn(2 * n, device="cuda").half()
    x0, d0 = x.double().clone(), den_prev.double().clone()
    nxt = torch.empty(2 * n, device="cuda", dtype=torch.float16)
    sp, s, sn, g = 5.0, 3.0, 2.0, 7.5
    dpmpp2m_step(x, eps, den_prev, nxt, sp, s, sn, g, first=False)
    e = eps.double()
    ee = e[:n] + g * (e[n:] - e[:n])
    den = x0 - s * ee
    h, hl = math.log(s / sn), math.log(sp / s)
    r = hl / h
    dd = (1 + 1 / (2 * r)) * den - (1 / (2 * r)) * d0
    want = (sn / s) * x0 - math.expm1(-h) * dd
    assert torch.allclose(x.double(), want, rtol=2e-5, atol=2e-5)
    assert torch.allclose(den_prev.double(), den, rtol=2e-6, atol=2e-6)
    assert torch.allclose(nxt[:n].double(), want / math.sqrt(sn * sn + 1), rtol=2e-3, atol=2e-3) and torch.equal(nxt[:n], nxt[n:])
