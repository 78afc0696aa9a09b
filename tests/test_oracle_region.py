"""Pins for the region-map oracle: cv2 itself (present in the image), the unmodified reference module
(build container only) and the golden fixtures produced by it."""
import glob
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import region_map as orm

from .helpers import NEG_IDS, PROMPT_IDS, StubTokenizer, ellipse_map, rect_map, state_from_golden, two_rect_state

cv2 = pytest.importorskip("cv2")
needs_ref = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")
TOK = lambda phrase: StubTokenizer()(phrase).input_ids  # noqa: E731


def _masks(H, W, rng):
    yy, xx = np.mgrid[0:H, 0:W]
    yield (rng.random((H, W)) < 0.5).astype(np.uint8)
    yield (rng.random((H, W)) < 0.1).astype(np.uint8)
    yield (((yy - H * 0.4) / (H * 0.3)) ** 2 + ((xx - W * 0.55) / (W * 0.25)) ** 2 < 1).astype(np.uint8)
    yield ((yy + xx) % 2).astype(np.uint8)  # exact 0.5 ties everywhere
    yield ((yy // 3 + xx // 5) % 2).astype(np.uint8)
    yield np.ones((H, W), np.uint8)
    yield np.zeros((H, W), np.uint8)


@pytest.mark.parametrize("H,W", [(512, 512), (768, 768), (512, 768), (1088, 1920), (64, 64), (128, 192)])
def test_cubic_rule_equals_cv2_for_integer_scales(H, W):
    rng = np.random.default_rng(H * 7 + W)
    for sr in (8, 16, 32, 64):
        if H % sr or W % sr:
            continue
        w_r, h_r = W // sr, H // sr
        for m in _masks(H, W, rng):
            a = cv2.resize(m, (w_r, h_r), interpolation=cv2.INTER_CUBIC)
            b = orm.cubic_resize_binary(m, w_r, h_r)
            assert np.array_equal(a, b), (H, W, sr)


def test_integer_scale_taps_are_minus3_19_19_minus3_over_32():
    ofs, coef = orm.cubic_tables(512, 64)
    assert np.array_equal(coef, np.tile(np.float32([-3, 19, 19, -3]) / 32, (64, 1)))
    assert np.array_equal(ofs, np.arange(64) * 8 + 3)


def test_non_integer_scale_differs_from_cv2_only_at_ties():
    """Stretch sizes (not multiples of 64): only pixels whose exact sum is within 1e-5 of 0.5 may differ."""
    rng = np.random.default_rng(5)
    H, W = 520, 776
    for sr in (16, 32, 64):
        w_r, h_r = -(-W // sr), -(-H // sr)
        m = (rng.random((H, W)) < 0.5).astype(np.uint8)
        a = cv2.resize(m, (w_r, h_r), interpolation=cv2.INTER_CUBIC)
        b = orm.cubic_resize_binary(m, w_r, h_r)
        assert (a != b).mean() < 2e-3


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "region_*.npz"))))
def test_oracle_matches_golden(path, capsys):
    z = np.load(path, allow_pickle=False)
    state = state_from_golden(z)
    out = orm.encode_region_map(state, TOK, int(z["width"]), int(z["height"]), int(z["n_img"]), text_ids=[z["neg"], z["ids"]])
    keys = sorted(int(k[2:]) for k in z.files if k.startswith("W_"))
    assert sorted(out.keys()) == keys
    for L in keys:
        assert torch.equal(out[L], torch.from_numpy(z[f"W_{L}"])), L


@needs_ref
def test_oracle_matches_reference_module_live():
    ref = ref_loader.encode_region_map_function()
    pipe = SimpleNamespace(tokenizer=StubTokenizer(), unet=SimpleNamespace(down_blocks=[0] * 4), vae_scale_factor=8,
                           do_classifier_free_guidance=True)
    ids, neg = np.array([PROMPT_IDS]), np.array([NEG_IDS])
    state = two_rect_state(512, 512)
    state["on the"] = {"map": ellipse_map(512, 512, 100, 400, 60, 80), "weight": 1.1, "mask_outsides": 0.4}
    state["the"] = {"map": rect_map(512, 512, 90, 130, 380, 420), "weight": 0.9, "mask_outsides": 0.15}  # overlaps "on the"
    a = ref.encode_region_map(pipe, state, 512, 512, 3, text_ids=[neg, ids])
    b = orm.encode_region_map(state, TOK, 512, 512, 3, text_ids=[neg, ids])
    assert a.keys() == b.keys()
    for L in a:
        assert torch.equal(a[L], b[L]) and a[L].shape == (6, L, 77)
        assert torch.equal(a[L][0], a[L][1])  # quirk: uncond half == cond half (ids overwritten at :91)


def test_quirks():
    ids, neg = np.array([PROMPT_IDS]), np.array([NEG_IDS])
    # a region that vanishes at coarse levels gets weight S everywhere there (== max with max 0)
    tiny = {"bridge": {"map": rect_map(512, 512, 100, 102, 100, 102), "weight": 0.8, "mask_outsides": 0.3}}
    out = orm.encode_region_map(tiny, TOK, 512, 512, 1, text_ids=[neg, ids])
    assert torch.all(out[64][0, :, 6] == np.float32(0.8))
    # weight 0 turns every pixel into -S'
    zero = {"bridge": {"map": rect_map(512, 512, 0, 255, 0, 255), "weight": 0.0, "mask_outsides": 0.3}}
    out = orm.encode_region_map(zero, TOK, 512, 512, 1, text_ids=[neg, ids])
    assert torch.all(out[4096][1, :, 6] == np.float32(-0.3))
    # regions off: still a dict of all-zero maps
    out = orm.encode_region_map(None, TOK, 512, 512, 2, text_ids=[neg, ids])
    assert all(t.shape == (4, L, 77) and not t.any() for L, t in out.items())
