"""Additive attention masks on the region path (north_star: a = Q K^T + M; attention_modify.py:84-91, baddbmm variant
:39-70): the CUDA path (dsc_xattn_call_masked / dsc_xattn_stats with a mask, mma.sync kernels) against the fp32 oracle,
whose mask handling is pinned bit for bit to the unmodified reference module in tests/test_oracle_attention.py, and against
the golden output of the reference's baddbmm processor with a mask (tests/golden/procm_baddbmm_*.npz).

Tolerances as everywhere (BASELINE north_star): rel-L2 <= 2e-3 per attention output, std relative error <= 1e-5."""
import glob
import os

import pytest
import torch

from oracle import attention as oa

from .helpers import AttnModule, baddbmm_mask_fixture, make_qkv, rel_l2, synthetic_w, weight_func

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL, STD_TOL = 2e-3, 1e-5


def _masked_reference(q, k, v, W, sigma, M, scale=None):
    """softmax(a + sigma * std(a) * W) V with a = scale Q K^T + M in fp32: what both reference variants compute once a
    mask is added (the function restated by oracle.attention.region_attention for masks that broadcast into [L, S], the
    baddbmm processor for [B*H, 1 or L, S]); written out here so that every broadcast form has one expectation."""
    q, k, v = q.float(), k.float(), v.float()
    B, H, L, D = q.shape
    a = q @ k.transpose(-2, -1) * (D ** -0.5 if scale is None else scale) + M.float()
    cw = oa.weight_func(W.float(), sigma, a)
    a = a + torch.repeat_interleave(cw, repeats=B // W.shape[0], dim=0).unsqueeze(1)
    return torch.softmax(a, dim=-1) @ v, a


def _mask(shape, seed, device="cuda"):
    g = torch.Generator().manual_seed(seed)
    m = torch.randn(shape, generator=g) * 0.75
    m[..., 60:] -= 4.0
    return m.to(device)


SHAPES = [(2, 8, 256, 40, 77), (2, 4, 152, 80, 77), (2, 8, 64, 160, 77), (4, 5, 100, 64, 40), (1, 2, 37, 128, 77)]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,H,L,D,S", SHAPES)
@pytest.mark.parametrize("form", ["S", "LS", "BH1S", "BHLS", "B1LS", "1H1S"])
def test_masked_call_matches_oracle(B, H, L, D, S, form, dtype):
    import diffusionspatialcontrol_b200 as dsc

    shape = {"S": (S,), "LS": (L, S), "BH1S": (B, H, 1, S), "BHLS": (B, H, L, S), "B1LS": (B, 1, L, S), "1H1S": (1, H, 1, S)}[form]
    q, k, v = make_qkv(B, H, L, D, S, seed=31, dtype=dtype, device="cuda")
    W = synthetic_w(B // 2 if B > 1 else 1, L, S).cuda()
    M = _mask(shape, seed=L + len(form))
    sigma = torch.tensor(5.0, device="cuda")
    got = dsc.region_attention(q, k, v, W, sigma, attn_mask=M)
    want, _ = _masked_reference(q, k, v, W, sigma, M)
    tol = TOL if dtype == torch.float16 else 2 * TOL  # bf16 P has 8 mantissa bits: same allowance as tests/test_gpu_attention.py
    assert rel_l2(got.float(), want) <= tol
    unmasked = dsc.region_attention(q, k, v, W, sigma)
    assert rel_l2(unmasked.float(), want) > 10 * tol  # the mask is really applied
    if form in ("S", "LS"):  # ... and the oracle function (pinned to the reference) says the same for what it accepts
        orc = oa.region_attention(q.float(), k.float(), v.float(), W.clone(), sigma, attn_mask=M.clone())
        assert rel_l2(got.float(), orc) <= tol


@pytest.mark.parametrize("B,H,L,D,S", SHAPES[:4] + [(2, 8, 64, 40, 154)])
def test_masked_std_is_the_std_of_the_masked_scores(B, H, L, D, S):
    from diffusionspatialcontrol_b200 import attention as att

    q, k, v = make_qkv(B, H, L, D, S, seed=5, device="cuda")
    for shape in ((B, H, 1, S), (B, H, L, S)):
        M = _mask(shape, seed=S)
        M[..., -3:] = -10000.0  # the usual "large negative" padding mask: finite, dominates the std
        ws = att.score_stats(q, k, attn_mask=M, workspace=torch.zeros(att.workspace_bytes(B, H, L, D, S), dtype=torch.uint8, device="cuda"))
        a = (q.double() @ k.double().transpose(-2, -1)) * D ** -0.5 + M.double()
        st = att.read_stats(ws)
        assert st["n"] == a.numel()
        assert abs(st["std"] - float(a.std())) <= STD_TOL * float(a.std())
        assert abs(st["mean"] - float(a.mean())) <= 1e-5 * abs(float(a.mean())) + 1e-6


def test_long_prompt_with_mask():
    """S = 154 (two 77-token windows, prompt_parser.py:161-194): key chunks of 80, mask columns follow the chunk."""
    import diffusionspatialcontrol_b200 as dsc

    B, H, L, D, S = 2, 8, 144, 40, 154
    q, k, v = make_qkv(B, H, L, D, S, seed=9, device="cuda")
    W = synthetic_w(B, L, S).cuda()
    W[:, :, 100] = 0.3
    sigma = torch.tensor(3.0, device="cuda")
    for shape in ((S,), (B, H, 1, S), (B, H, L, S)):
        g = torch.Generator().manual_seed(3)
        M = (torch.randn(shape, generator=g) * 0.75).cuda()
        M[..., 130:] -= 5.0
        got = dsc.region_attention(q, k, v, W, sigma, attn_mask=M)
        want, _ = _masked_reference(q, k, v, W, sigma, M)
        assert rel_l2(got.float(), want) <= TOL


def test_bool_mask_is_ignored_and_inf_mask_is_nan_like_the_reference():
    import diffusionspatialcontrol_b200 as dsc

    B, H, L, D, S = 2, 8, 64, 40, 77
    q, k, v = make_qkv(B, H, L, D, S, seed=2, device="cuda")
    W = synthetic_w(B, L, S).cuda()
    sigma = torch.tensor(4.0, device="cuda")
    mb = torch.ones(B, H, 1, S, dtype=torch.bool, device="cuda")
    mb[..., 40:] = False
    assert torch.equal(dsc.region_attention(q, k, v, W, sigma, attn_mask=mb), dsc.region_attention(q, k, v, W, sigma))
    assert bool(mb[..., :40].all()) and not bool(mb[..., 40:].any())  # (the caller's mask is left alone)
    M = torch.zeros(S, device="cuda")
    M[70:] = float("-inf")
    want, _ = _masked_reference(q, k, v, W, sigma, M)
    got = dsc.region_attention(q, k, v, W, sigma, attn_mask=M)
    assert torch.isnan(want).all() and torch.isnan(got.float()).all()  # std(a) is NaN: the beta term poisons every row


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "procm_baddbmm_*.npz"))))
def test_baddbmm_processor_with_mask_matches_reference_class_output(path):
    """``RegionAttnProcessorBaddbmm`` (fp16, CUDA kernels) with an attention mask against the fp32 output of the UNMODIFIED
    reference ``AttnProcessor`` (attention_modify.py:107-207; scripts/gen_golden.py) and against the oracle's restatement."""
    from diffusionspatialcontrol_b200.attention_processor import RegionAttnProcessorBaddbmm

    attn32, hs, ctx, rp, want, mask = baddbmm_mask_fixture(path)
    attn16 = AttnModule(hs.shape[-1], attn32.heads, round(attn32.scale ** -2))
    attn16.load_state_dict(attn32.state_dict())
    attn16 = attn16.cuda().half()
    rp_dev = {**rp, "sigma": rp["sigma"].cuda()}
    with torch.no_grad():
        got = RegionAttnProcessorBaddbmm()(attn16, hs.cuda().half(), encoder_hidden_states=ctx.cuda().half(),
                                           attention_mask=mask.cuda().half(), region_prompt=rp_dev)
        plain = RegionAttnProcessorBaddbmm()(attn16, hs.cuda().half(), encoder_hidden_states=ctx.cuda().half(), region_prompt=rp_dev)
        orc = oa.processor_forward_baddbmm(attn32.cuda(), hs.cuda(), ctx.cuda(), {**rp, "region_state": {
            L: w.cuda() for L, w in rp["region_state"].items()}}, attention_mask=mask.cuda())
    assert rel_l2(got.float(), want) <= 4e-3  # (fp16 projections are PyTorch's: gate as for the unmasked processors)
    assert rel_l2(got.float(), orc) <= 4e-3
    assert rel_l2(plain.float(), want) > 4e-2


def test_processors_treat_masks_as_the_reference_classes_do():
    """SDPA-style processor: bool mask ignored, float mask -> the reference's RuntimeError (attention_modify.py:86-89 with
    the 4-D mask of :452); baddbmm processor: mask of another dtype than the query -> RuntimeError (torch.baddbmm input);
    a zero region map with a mask: plain attention WITH the mask (W = 0 adds nothing, the mask stays), baddbmm only."""
    from diffusionspatialcontrol_b200 import RegionAttnProcessor
    from diffusionspatialcontrol_b200.attention_processor import RegionAttnProcessorBaddbmm

    torch.manual_seed(6)
    C, H, D, L, B = 320, 8, 40, 96, 2
    attn = AttnModule(C, H, D).cuda().half()
    hs, ctx = torch.randn(B, L, C, device="cuda").half(), torch.randn(B, 77, 768, device="cuda").half()
    rp = {"region_state": {L: synthetic_w(B, L, 77)}, "sigma": torch.tensor(3.0, device="cuda"), "weight_func": weight_func}
    mb = torch.ones(B * H, 1, 77, dtype=torch.bool, device="cuda")
    mb[..., 30:] = False
    mf = _mask((B * H, 1, 77), seed=1).half()
    with torch.no_grad():
        base = RegionAttnProcessor()(attn, hs, encoder_hidden_states=ctx, region_prompt=rp)
        assert torch.equal(RegionAttnProcessor()(attn, hs, encoder_hidden_states=ctx, attention_mask=mb, region_prompt=rp), base)
        with pytest.raises(RuntimeError, match="broadcast shape"):
            RegionAttnProcessor()(attn, hs, encoder_hidden_states=ctx, attention_mask=mf, region_prompt=rp)
        with pytest.raises(RuntimeError):
            RegionAttnProcessorBaddbmm()(attn, hs, encoder_hidden_states=ctx, attention_mask=mb, region_prompt=rp)
        rp0 = {**rp, "region_state": {L: torch.zeros(B, L, 77)}}
        got0 = RegionAttnProcessorBaddbmm()(attn, hs, encoder_hidden_states=ctx, attention_mask=mf, region_prompt=rp0)
        a32 = AttnModule(C, H, D).cuda()
        a32.load_state_dict({k: v.float() for k, v in attn.state_dict().items()})
        want0 = oa.processor_forward_baddbmm(a32, hs.float(), ctx.float(), {**rp0, "region_state": {L: torch.zeros(B, L, 77, device="cuda")}},
                                             attention_mask=mf.float())
        assert rel_l2(got0.float(), want0) <= 4e-3
