"""Host-side logic that needs no GPU: processor dispatch rules, kwarg filtering of the UNet host, token
span search, schedules, and the refusal to run the region path on the CPU."""
import numpy as np
import pytest
import torch

from oracle import attention as oa
from oracle import sampler as osm

from .helpers import PROMPT_IDS, synthetic_w, weight_func


def test_processor_stock_paths_run_anywhere_and_region_path_refuses_cpu():
    from diffusionspatialcontrol_b200 import RegionAttnProcessor
    from diffusionspatialcontrol_b200.unet_sd15 import Attention

    torch.manual_seed(0)
    proc = RegionAttnProcessor()
    cross = Attention(64, 96, heads=2, dim_head=40)
    self_ = Attention(64, None, heads=2, dim_head=40)
    hs, ctx = torch.randn(2, 16, 64), torch.randn(2, 77, 96)
    rp = {"region_state": {16: synthetic_w(2, 16, 77)}, "sigma": torch.tensor(2.0), "weight_func": weight_func}
    with torch.no_grad():
        # self-attention ignores region_prompt (reference :437-438); cross-attention without region_prompt and with a
        # non-dict region_state (regions unavailable, reference :22-23/:479) take the stock SDPA path
        for attn, kw in ((self_, dict(region_prompt=rp)), (cross, dict(encoder_hidden_states=ctx)),
                         (cross, dict(encoder_hidden_states=ctx, region_prompt={**rp, "region_state": torch.FloatTensor(0)}))):
            got = proc(attn, hs, **kw)
            want = oa.processor_forward(attn, hs, kw.get("encoder_hidden_states"), kw.get("region_prompt"))
            assert torch.allclose(got, want, atol=1e-6)
        with pytest.raises(RuntimeError, match="no CPU path"):
            proc(cross, hs, encoder_hidden_states=ctx, region_prompt=rp)
        with pytest.raises(NotImplementedError):
            proc(cross, hs, encoder_hidden_states=ctx, region_prompt={**rp, "weight_func": lambda w, s, qk: w * s * qk.var()})
        with pytest.raises(KeyError):
            proc(cross, hs, encoder_hidden_states=ctx, region_prompt={**rp, "region_state": {999: rp["region_state"][16]}})


def test_processor_signature_names_match_the_reference():
    import inspect

    from diffusionspatialcontrol_b200 import RegionAttnProcessor

    params = list(inspect.signature(RegionAttnProcessor.__call__).parameters)
    assert params == ["self", "attn", "hidden_states", "encoder_hidden_states", "attention_mask", "temb", "scale",
                      "region_prompt", "ip_adapter_masks"]  # attention_modify.py:414-424


def test_unet_host_plumbing():
    from diffusionspatialcontrol_b200.unet_sd15 import UNetSD15, cross_attention_shapes

    torch.manual_seed(0)
    net = UNetSD15(block_out_channels=(32, 64, 128, 128), heads=2, cross_attention_dim=48).eval()
    assert len(net.attn_processors) == 32
    seen = []

    class Spy:
        def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0,
                     region_prompt=None, ip_adapter_masks=None):
            seen.append((encoder_hidden_states is not None, hidden_states.shape[1], region_prompt))
            from diffusionspatialcontrol_b200.unet_sd15 import DefaultAttnProcessor

            return DefaultAttnProcessor()(attn, hidden_states, encoder_hidden_states)

    net.set_attn_processor(Spy())
    with torch.no_grad():
        y = net(torch.randn(2, 4, 32, 32), torch.tensor(10.0), torch.randn(2, 77, 48),
                cross_attention_kwargs={"region_prompt": "RP", "not_a_processor_kwarg": 1})
    assert y.shape == (2, 4, 32, 32) and len(seen) == 32
    assert all(rp == "RP" for _, _, rp in seen)  # reaches attn1 and attn2 of every block; unknown kwargs are dropped
    cross = [L for is_x, L, _ in seen if is_x]
    assert cross == [L // 4 for L, _ in cross_attention_shapes(512, 512)]  # 32x32 latent = 256x256 image
    with pytest.raises(ValueError):
        net.set_attn_processor({"a": Spy()})
    net.set_attn_processor({k: Spy() for k in net.attn_processors})
    assert [(L, D) for L, D in cross_attention_shapes()] [:3] == [(4096, 40), (4096, 40), (1024, 80)]


def test_span_search_matches_reference_rule():
    from diffusionspatialcontrol_b200.region_map import _spans

    ids = PROMPT_IDS[:10] + [320, 1611] + PROMPT_IDS[12:]
    spans, found = _spans(ids, [[320, 1611], [2465], [1929], []])
    assert spans == [(0, 1, 2), (0, 10, 2), (1, 6, 1)] and found == [True, True, False, False]
    assert _spans(None, [[1]]) == ([], [False])


def test_schedule_matches_sampler_oracle():
    from diffusionspatialcontrol_b200.sampler import KarrasSchedule

    s = KarrasSchedule(25)
    train = osm.sd15_train_sigmas()
    want = osm.get_sigmas_karras(25, train[0].item(), train[-1].item())
    assert torch.equal(s.sigmas, want)
    assert torch.equal(s.timesteps, osm.sigma_to_t(want[:-1], train.log()))
    assert abs(s.c_in(0) - 1 / (14.6146469**2 + 1) ** 0.5) < 1e-6


def test_sampler_and_region_entry_points_refuse_cpu():
    from diffusionspatialcontrol_b200 import encode_region_map_sp
    from diffusionspatialcontrol_b200.sampler import dpmpp2m_step

    with pytest.raises(RuntimeError, match="no CPU path"):
        dpmpp2m_step(torch.zeros(4), torch.zeros(8).half(), torch.zeros(4), None, 0, 2.0, 1.0, 7.5, True)
    from types import SimpleNamespace

    with pytest.raises(RuntimeError, match="no CPU path"):
        encode_region_map_sp(None, None, SimpleNamespace(down_blocks=[0]), 64, 64, text_ids=[np.array([1]), np.array([1])],
                             device="cpu")
    assert encode_region_map_sp(None, None, None, 64, 64, text_ids=None).numel() == 0  # reference :22-23


def test_region_map_layouts_keep_the_reference_values():
    """padded_region_map / compact_region_map are pure re-layouts of the reference's [B', L, n_tok] tensor (no GPU needed)."""
    import torch

    from diffusionspatialcontrol_b200 import compact_region_map, padded_region_map
    from diffusionspatialcontrol_b200.attention import _region_layout_ok

    from .helpers import synthetic_w

    W = synthetic_w(2, 96, 77)
    P = padded_region_map(W)
    assert P.shape == W.shape and P.stride() == (96 * 80, 80, 1) and torch.equal(P, W) and _region_layout_ok(P)
    assert padded_region_map(P) is P  # already in the fast layout: returned as is
    Wc, cols = compact_region_map(W)
    assert cols == [1, 2, 3, 6] and Wc.shape == (2, 96, 20) and Wc.is_contiguous()
    assert torch.equal(Wc[:, :, :4], W[:, :, cols]) and not Wc[:, :, 4:].any()
    zc, zcols = compact_region_map(torch.zeros(1, 8, 77))
    assert zcols == [] and not zc.any()  # regions off: nothing weighted
    many = torch.zeros(1, 8, 77)
    many[0, 0, :17] = 1.0
    assert compact_region_map(many) is None  # more than 16 weighted columns: no compact form
    long = synthetic_w(1, 32, 154)
    PL = padded_region_map(long)
    assert PL.stride(1) == 160 and torch.equal(PL, long) and _region_layout_ok(PL)


def test_static_region_maps_must_be_padded_device_tensors():
    import torch

    from diffusionspatialcontrol_b200 import RegionAttnProcessor

    proc = RegionAttnProcessor()
    with pytest.raises(ValueError):
        proc.register_static_map(torch.zeros(1, 8, 77))  # CPU / dense: would be copied at first use, i.e. frozen in a graph


def test_nvtx_ranges_are_off_by_default_and_wrap_when_enabled(monkeypatch):
    """DSC_NVTX=1 wraps the kernel entry points in NVTX ranges; by default the functions are returned as they are."""
    from diffusionspatialcontrol_b200 import _lib

    def f(x):
        return x + 1

    monkeypatch.setattr(_lib, "NVTX", False)
    assert _lib.nvtx("k")(f) is f
    monkeypatch.setattr(_lib, "NVTX", True)
    g = _lib.nvtx("k")(f)
    assert g is not f and g.__wrapped__ is f and g.__name__ == "f"
