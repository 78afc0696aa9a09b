"""Parity of the CUDA attention path (through the C ABI) against the oracle -- the first gate.

Tolerances (BASELINE north_star): per-layer attention output ||o - o_ref||_2 / ||o_ref||_2 <= 2e-3 for
fp16/bf16 kernels versus the fp32 reference on identical (low-precision-representable) inputs;
std(a) relative error <= 1e-5 versus an fp64 evaluation.
"""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import attention as oa

from .helpers import make_qkv, rel_l2, synthetic_w, weight_func

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["tc5", "tc5-padded", "mma", "mma-padded", "fused", "fused-padded", "tc5fused", "tc5fused-padded",
                                      "tc5-compact"])
def impl(request, monkeypatch, dsc_config):
    """Every test runs against both kernel families: tcgen05/TMEM (default where implemented: D=40, 80)
    and the legacy mma.sync path (all head dims; also the cross-check of the first) -- and with the region map in
    both device layouts: dense [B', L, 77] as the reference builds it, and the padded fast layout (rows 80 floats
    apart) that the processor's cache and encode_region_map produce."""
    family, _, layout = request.param.partition("-")
    # "fused": the single-launch mma.sync kernel (both passes, Q resident on chip) wherever the problem fits, two-pass
    # mma.sync elsewhere; "tc5fused": the single-launch two-phase tcgen05 kernel (D = 40 / 80), two-pass elsewhere;
    # "mma" / "tc5": always two launches
    dsc_config("xattn_impl", {"fused": "mma", "tc5fused": "tc5"}.get(family, family))
    dsc_config("no_fused", "0" if family in ("fused", "tc5fused") else "1")
    dsc_config("tc5_fused", "1" if family == "tc5fused" else "0")
    if layout == "compact":  # tcgen05 pass 2 fed with the compact region map (weighted key columns only, keys permuted)
        from diffusionspatialcontrol_b200 import attention as att

        monkeypatch.setattr(att, "AUTO_COMPACT", True)
    if layout == "padded":
        from diffusionspatialcontrol_b200 import attention as att

        dense_ok = att._region_layout_ok
        monkeypatch.setattr(att, "_region_layout_ok", lambda W: W.stride(1) == att.MAX_KEYS and dense_ok(W))
    return request.param


TOL = 2e-3
STD_TOL = 1e-5


def _dsc():
    import diffusionspatialcontrol_b200 as dsc
    from diffusionspatialcontrol_b200 import attention as att

    return dsc, att


def _oracle(q, k, v, W, sigma, device="cuda"):
    """fp32 oracle evaluated on `device` (plain torch ops; TF32 off)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    sig = sigma.float().to(device) if isinstance(sigma, torch.Tensor) else sigma
    return oa.region_attention(q.float().to(device), k.float().to(device), v.float().to(device),
                               W.float().to(device).clone(), sig)


def _std64(q, k):
    a = (q.double() @ k.double().transpose(-2, -1)) * (q.shape[-1] ** -0.5)
    return float(a.std()), float(a.sum()), float((a * a).sum())


SD15 = [(4096, 40), (1024, 80), (256, 160), (64, 160), (9216, 40), (2304, 80), (576, 160), (144, 160)]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("L,D", SD15)
def test_sd15_layer_shapes_match_oracle(L, D, dtype):
    dsc, att = _dsc()
    B, H, S = 2, 8, 77
    q, k, v = make_qkv(B, H, L, D, S, seed=L + D, dtype=dtype, device="cuda")
    W = synthetic_w(B, L, S).cuda()
    for sigma in (14.6146, 0.0292):
        out = dsc.region_attention(q, k, v, W, sigma)
        ref = _oracle(q, k, v, W, sigma)
        assert out.shape == ref.shape == (B, H, L, D)
        err = rel_l2(out.float(), ref)
        assert err <= TOL, f"L={L} D={D} {dtype} sigma={sigma}: rel-L2 {err:.3e}"
    st = att.read_stats(att.get_workspace(q.device))
    std, s, ss = _std64(q, k)
    assert st["ticket"] == 0 and st["n"] == B * H * L * S
    assert abs(st["std"] - std) / std <= STD_TOL, (st, std)
    assert abs(st["sumsq"] - ss) / ss <= 1e-5 and abs(st["sum"] - s) <= 1e-5 * (ss * st["n"]) ** 0.5


@pytest.mark.parametrize(
    "B,H,L,D,S,Bw",
    [
        (1, 8, 16, 40, 77, 1),      # a single slice
        (2, 8, 100, 40, 77, 2),     # tail slice of 4 rows (W tile still 16-byte sized)
        (2, 8, 37, 80, 77, 1),      # tail of 5 rows: W tile falls back to plain loads
        (3, 8, 200, 160, 77, 3),    # odd batch
        (4, 8, 64, 40, 77, 1),      # one map row shared by the whole batch
        (4, 8, 64, 40, 77, 2),      # B / Bw = 2 (batch-major repeat_interleave)
        (2, 5, 72, 64, 77, 2),      # SD-2.x style heads: 5 x 64
        (2, 3, 48, 160, 77, 2),     # partial last head group (G = 2, H = 3)
        (2, 12, 48, 40, 77, 2),     # two head groups at D = 40, second one partial
        (2, 8, 64, 80, 80, 2),      # S = 80: no padded keys
        (2, 8, 64, 80, 40, 2),      # short prompts
        (2, 8, 64, 128, 1, 2),      # degenerate single key
        (16, 8, 1024, 80, 77, 16),  # config 2 batch, mid resolution
    ],
)
def test_edge_shapes_match_oracle(B, H, L, D, S, Bw):
    dsc, _ = _dsc()
    q, k, v = make_qkv(B, H, L, D, S, seed=B * 1000 + L + D + S, device="cuda")
    W = synthetic_w(Bw, L, S).cuda()
    out = dsc.region_attention(q, k, v, W, 5.0)
    ref = _oracle(q, k, v, W, 5.0)
    err = rel_l2(out.float(), ref)
    assert err <= TOL, f"rel-L2 {err:.3e}"
    assert torch.isfinite(out.float()).all()


def test_full_size_config2_layer_matches_oracle():
    """BASELINE configs[1]: batch 8 + CFG -> attention batch 16 at the largest layer (L=4096, D=40)."""
    dsc, att = _dsc()
    B, H, L, D, S = 16, 8, 4096, 40, 77
    q, k, v = make_qkv(B, H, L, D, S, seed=2, device="cuda")
    W = synthetic_w(B, L, S).cuda()
    out = dsc.region_attention(q, k, v, W, 7.0944)
    ref = _oracle(q, k, v, W, 7.0944)
    assert rel_l2(out.float(), ref) <= TOL
    # size-independent property: every softmax row sums to 1 => with V = ones the output is ones
    ones = torch.ones_like(v)
    o1 = dsc.region_attention(q, k, ones, W, 7.0944)
    assert torch.allclose(o1.float(), torch.ones_like(o1.float()), atol=2e-3)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "attn_*.npz"))))
def test_golden_vectors_from_the_reference(path):
    dsc, att = _dsc()
    z = np.load(path)
    H = int(z["heads"])
    q, k, v = (torch.from_numpy(z[n]).cuda() for n in "qkv")
    B, L, HD = q.shape
    D = HD // H
    view = lambda t: t.view(B, -1, H, D).transpose(1, 2)
    out = dsc.region_attention(view(q), view(k), view(v), torch.from_numpy(z["W"]).cuda(), float(z["sigma"]))
    got = out.transpose(1, 2).reshape(B, L, HD).float().cpu()
    assert rel_l2(got, torch.from_numpy(z["out"])) <= TOL
    st = att.read_stats(att.get_workspace(q.device))
    assert abs(st["std"] - float(z["std"])) / float(z["std"]) <= 1e-5


def test_sigma_sources_agree():
    """sigma as python float, CPU 0-dim tensor, fp32 / fp16 CUDA 0-dim tensor (k-diffusion path)."""
    dsc, _ = _dsc()
    q, k, v = make_qkv(2, 8, 256, 40, 77, seed=5, device="cuda")
    W = synthetic_w(2, 256, 77).cuda()
    s16 = torch.tensor(3.5, dtype=torch.float16)  # exactly representable
    base = dsc.region_attention(q, k, v, W, 3.5)
    for sig in (torch.tensor(3.5), torch.tensor(3.5, device="cuda"), s16.cuda(), torch.tensor([3.5], device="cuda")[0]):
        assert torch.equal(dsc.region_attention(q, k, v, W, sig), base)


def test_zero_map_or_zero_sigma_is_plain_sdpa():
    dsc, _ = _dsc()
    q, k, v = make_qkv(2, 8, 1024, 80, 77, seed=9, device="cuda")
    W = synthetic_w(2, 1024, 77).cuda()
    want = F.scaled_dot_product_attention(q.float(), k.float(), v.float())
    a = dsc.region_attention(q, k, v, torch.zeros_like(W), 9.0)
    b = dsc.region_attention(q, k, v, W, 0.0)
    assert rel_l2(a.float(), want) <= TOL and rel_l2(b.float(), want) <= TOL
    # the two routes differ only in how the logit is rounded (beta*(s*scale/beta + 0) vs s*scale + 0*W)
    assert rel_l2(a.float(), b.float()) <= 2e-4


def test_deterministic_and_workspace_reusable():
    dsc, att = _dsc()
    q, k, v = make_qkv(4, 8, 1024, 80, 77, seed=4, device="cuda")
    W = synthetic_w(4, 1024, 77).cuda()
    outs = [dsc.region_attention(q, k, v, W, 2.0).clone() for _ in range(3)]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert att.read_stats(att.get_workspace(q.device))["ticket"] == 0


def test_batch_coupling_matches_reference_semantics():
    """std spans batch x heads x queries x keys of ONE call: scaling one sample changes the others."""
    dsc, _ = _dsc()
    q, k, v = make_qkv(2, 8, 256, 40, 77, seed=12, device="cuda")
    W = synthetic_w(1, 256, 77).cuda()
    q2 = q.clone()
    q2[1] *= 3.0
    a = dsc.region_attention(q, k, v, W, 10.0)
    b = dsc.region_attention(q2, k, v, W, 10.0)
    assert not torch.allclose(a[0].float(), b[0].float(), atol=1e-3)
    assert rel_l2(b.float(), _oracle(q2, k, v, W, 10.0)) <= TOL


def test_gram_identity_for_the_stats_pass():
    """With no mask, sum a^2 = scale^2 <Q^T Q, K^T K>_F per (batch, head) -- independent check of pass 1."""
    _, att = _dsc()
    B, H, L, D, S = 2, 8, 1024, 80, 77
    q, k, _ = make_qkv(B, H, L, D, S, seed=21, device="cuda")
    ws = att.score_stats(q, k)
    st = att.read_stats(ws)
    qd, kd = q.double(), k.double()
    gram = ((qd.transpose(-2, -1) @ qd) * (kd.transpose(-2, -1) @ kd)).sum() / D
    ssum = (qd.sum(-2) * kd.sum(-2)).sum() / D**0.5
    assert abs(st["sumsq"] - float(gram)) / float(gram) <= 1e-5
    assert abs(st["sum"] - float(ssum)) <= 1e-5 * float(gram * st["n"]) ** 0.5


def test_layout_normalisation_and_errors():
    dsc, _ = _dsc()
    q, k, v = make_qkv(2, 8, 64, 40, 77, seed=1, device="cuda")
    W = synthetic_w(2, 64, 77).cuda()
    base = dsc.region_attention(q, k, v, W, 1.0)
    # contiguous [B,H,L,D] tensors are re-laid out, same numbers
    assert torch.equal(dsc.region_attention(q.contiguous(), k.contiguous(), v.contiguous(), W, 1.0), base)
    with pytest.raises(ValueError):
        dsc.region_attention(q, k, v, W[:, :32], 1.0)
    with pytest.raises(ValueError):
        dsc.region_attention(q, k, v, synthetic_w(3, 64, 77).cuda(), 1.0)
    with pytest.raises(RuntimeError):
        dsc.region_attention(q.cpu(), k.cpu(), v.cpu(), W.cpu(), 1.0)
    with pytest.raises(TypeError):
        dsc.region_attention(q.float(), k.float(), v.float(), W, 1.0)
    with pytest.raises(ValueError):  # additive masks must broadcast to [B, H, L, S] (tests/test_gpu_masks.py covers the rest)
        dsc.region_attention(q, k, v, W, 1.0, attn_mask=torch.zeros(63, 77, device="cuda"))
    with pytest.raises(TypeError):
        dsc.region_attention(q, k, v, W, 1.0, attn_mask=torch.zeros(64, 77, device="cuda", dtype=torch.int32))


class _Attn(nn.Module):
    def __init__(self, C, H, D, ctx=768):
        super().__init__()
        self.heads, self.scale = H, D**-0.5
        self.to_q, self.to_k, self.to_v = nn.Linear(C, H * D, bias=False), nn.Linear(ctx, H * D, bias=False), nn.Linear(ctx, H * D, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(H * D, C), nn.Dropout(0.0)])
        self.spatial_norm = self.group_norm = self.norm_cross = None
        self.residual_connection, self.rescale_output_factor = False, 1.0


@pytest.mark.parametrize("C,D,L", [(320, 40, 4096), (640, 80, 1024), (1280, 160, 256)])
def test_processor_matches_reference_processor_restatement(C, D, L):
    """Whole processor (projections in fp16 on PyTorch + our kernels) vs the fp32 restatement of
    AttnProcessor2_0.  The projections are not ours and add their own fp16 rounding: gate 4e-3."""
    dsc, _ = _dsc()
    torch.manual_seed(C)
    attn32 = _Attn(C, 8, D).cuda()
    attn16 = _Attn(C, 8, D).cuda().half()
    attn16.load_state_dict(attn32.state_dict())
    attn32.load_state_dict({k_: v_.float() for k_, v_ in attn16.state_dict().items()})  # identical (fp16-representable) weights
    hs = torch.randn(2, L, C, device="cuda").half()
    ctx = torch.randn(2, 77, 768, device="cuda").half()
    rp = {"region_state": {L: synthetic_w(2, L, 77)}, "sigma": torch.tensor(6.0, device="cuda"), "weight_func": weight_func}
    proc = dsc.RegionAttnProcessor()
    with torch.no_grad():
        got = proc(attn16, hs, encoder_hidden_states=ctx, region_prompt=rp)
        want = oa.processor_forward(attn32, hs.float(), ctx.float(), {**rp, "region_state": {L: rp["region_state"][L].cuda()}})
        # self-attention and region-less calls take the stock path
        self16 = _Attn(C, 8, D, ctx=C).cuda().half()
        a = proc(self16, hs, region_prompt=rp)
        b = oa.processor_forward(self16, hs, None, rp)
    assert rel_l2(got.float(), want) <= 4e-3
    assert torch.allclose(a.float(), b.float(), atol=1e-2, rtol=1e-2)
    with pytest.raises(NotImplementedError):
        proc(attn16, hs, encoder_hidden_states=ctx, region_prompt={**rp, "weight_func": lambda w, s, qk: w * s})
    with pytest.raises(KeyError):
        proc(attn16, hs[:, :64], encoder_hidden_states=ctx, region_prompt=rp)


def test_tc5_two_warpgroup_variant_matches_oracle(dsc_config):
    """D = 40 has two tcgen05 variants; the default tests exercise x4, this one forces x2 (software-pipelined heads)."""
    dsc, _ = _dsc()
    dsc_config("xattn_impl", "tc5")
    dsc_config("tc5_variant", "x2")
    for (B, H, L, S) in [(2, 8, 1024, 77), (16, 8, 4096, 77), (3, 12, 200, 40)]:
        q, k, v = make_qkv(B, H, L, 40, S, seed=L + S, device="cuda")
        W = synthetic_w(B, L, S).cuda()
        out = dsc.region_attention(q, k, v, W, 6.0)
        assert rel_l2(out.float(), _oracle(q, k, v, W, 6.0)) <= TOL


def test_config3_shapes_768_batch4_with_suppression():
    """BASELINE configs[2]: 768x768 (9216 / 2304 / 576 / 144 queries), batch 4 + CFG = 8, maps with negative entries
    (nonzero S' suppression) from the reference-generated golden region maps."""
    dsc, _ = _dsc()
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "region_768_4reg.npz"))
    for (L, D) in [(9216, 40), (2304, 80), (576, 160), (144, 160)]:
        W1 = torch.from_numpy(z[f"W_{L}"])  # [2, L, 77], uncond == cond rows (reference quirk)
        assert (W1 < 0).any() and (W1 > 0).any()
        W = W1.repeat(4, 1, 1).cuda()       # encode_region_map(...).repeat(num_images_per_prompt): [8, L, 77]
        q, k, v = make_qkv(8, 8, L, D, 77, seed=L, device="cuda")
        for sigma in (14.6146, 0.3350):
            out = dsc.region_attention(q, k, v, W, sigma)
            assert rel_l2(out.float(), _oracle(q, k, v, W, sigma)) <= TOL


@pytest.mark.parametrize("L,D,S,B", [(1024, 80, 154, 2), (4096, 40, 231, 2), (256, 160, 154, 4), (200, 40, 100, 3),
                                     (64, 160, 81, 2), (1024, 80, 308, 1)])
def test_long_prompts_run_as_key_chunks(L, D, S, B):
    """Long-prompt modes of the reference concatenate 77-token windows (S = 77 k, prompt_parser.py:161-194; the region
    map follows the id length, encode_region_map_function.py:30).  More than 80 keys run as chunks of 80 with the std
    of the WHOLE call and a log-sum-exp merge: same result as the oracle's single softmax over all keys."""
    dsc, att = _dsc()
    q, k, v = make_qkv(B, 8, L, D, S, seed=S + L, device="cuda")
    W = synthetic_w(B, L, S).cuda()
    W[:, : L // 3, S - 3] = -0.4  # weights in the last chunk too
    for sigma in (9.0, 0.05):
        out = dsc.region_attention(q, k, v, W, sigma)
        ref = _oracle(q, k, v, W, sigma)
        err = rel_l2(out.float(), ref)
        assert err <= TOL, f"L={L} D={D} S={S} sigma={sigma}: rel-L2 {err:.3e}"
    st = att.read_stats(att.score_stats(q, k))
    want, _, _ = _std64(q, k)
    assert abs(st["std"] - want) / want <= STD_TOL and st["n"] == B * 8 * L * S
    with pytest.raises(Exception):
        qq, kk, vv = make_qkv(1, 8, 64, 40, 481, seed=1, device="cuda")
        dsc.region_attention(qq, kk, vv, torch.zeros(1, 64, 481, device="cuda"), 1.0)


@pytest.mark.parametrize("B,L,S", [(2, 1024, 77), (16, 4096, 77), (3, 200, 40), (1, 64, 77), (4, 9216, 80)])
def test_gram_identity_stats_kernel(dsc_config, B, L, S):
    """DSC_XATTN_STATS_IMPL=gram: pass 1 without forming a single score (sum a^2 = scale^2 <Q^T Q, K^T K>, SURVEY 8(f)
    rank 2), D = 40.  Same published statistics as the score-based kernels, and the forward pass that consumes them."""
    dsc, att = _dsc()
    dsc_config("stats_impl", "gram")
    q, k, v = make_qkv(B, 8, L, 40, S, seed=B + L, device="cuda")
    st = att.read_stats(att.score_stats(q, k))
    want, wsum, wsq = _std64(q, k)
    assert st["n"] == B * 8 * L * S
    assert abs(st["std"] - want) / want <= STD_TOL, (st["std"], want)
    assert abs(st["sumsq"] - wsq) / wsq <= 1e-5 and abs(st["sum"] - wsum) <= 1e-5 * (wsq * st["n"]) ** 0.5
    W = synthetic_w(B, L, S).cuda()
    out = dsc.region_attention(q, k, v, W, 8.0)
    assert rel_l2(out.float(), _oracle(q, k, v, W, 8.0)) <= TOL


@pytest.mark.parametrize("n_adapters,with_masks", [(1, False), (2, True)])
def test_ip_adapter_processor_matches_oracle(n_adapters, with_masks):
    """Drop-in for the reference's IPAdapterAttnProcessor2_0 (attention_modify.py:506-700): region-masked text branch on
    the CUDA path + image-prompt branches, against the oracle restatement (itself pinned to the reference class)."""
    from diffusionspatialcontrol_b200 import RegionIPAdapterAttnProcessor
    from oracle import ip_adapter as oip

    torch.manual_seed(5)
    C, D, L, B = 640, 80, 1024, 2
    attn = _Attn(C, 8, D).cuda()
    hs, ctx = torch.randn(B, L, C, device="cuda"), torch.randn(B, 77, 768, device="cuda")
    ip = [torch.randn(B, 4 * (i + 1), 768, device="cuda") for i in range(n_adapters)]
    tokens, scales = [t.shape[1] for t in ip], [0.7, 0.3][:n_adapters]
    oracle = oip.OracleIPAdapterProcessor(C, 768, num_tokens=tokens, scale=scales).cuda()
    ours = RegionIPAdapterAttnProcessor(C, 768, num_tokens=tokens, scale=scales).cuda().half()
    ours.load_state_dict(oracle.state_dict())
    masks = None
    if with_masks:
        masks = torch.zeros(n_adapters, 1, 64, 64, device="cuda")
        masks[0, :, :, :32] = 1.0
        masks[1, :, 20:, :] = 1.0
    W = synthetic_w(B, L, 77)
    rp = {"region_state": {L: W}, "sigma": torch.tensor(6.5), "weight_func": weight_func}
    attn16 = _Attn(C, 8, D).cuda().half()
    attn16.load_state_dict(attn.state_dict())
    with torch.no_grad():
        # fp16-representable inputs on both sides
        hs16, ctx16, ip16 = hs.half(), ctx.half(), [t.half() for t in ip]
        attn32 = _Attn(C, 8, D).cuda()
        attn32.load_state_dict({k: v.float() for k, v in attn16.state_dict().items()})
        oracle32 = oip.OracleIPAdapterProcessor(C, 768, num_tokens=tokens, scale=scales).cuda()
        oracle32.load_state_dict({k: v.float() for k, v in ours.state_dict().items()})
        torch.backends.cuda.matmul.allow_tf32 = False
        want = oracle32(attn32, hs16.float(), encoder_hidden_states=(ctx16.float(), [t.float() for t in ip16]),
                        region_prompt={**rp, "region_state": {L: W.cuda()}}, ip_adapter_masks=masks)
        got = ours(attn16, hs16, encoder_hidden_states=(ctx16, ip16), region_prompt=rp,
                   ip_adapter_masks=masks.half() if masks is not None else None)
        plain = ours(attn16, hs16, encoder_hidden_states=(ctx16, [torch.zeros_like(t) for t in ip16]), region_prompt=rp)
    assert rel_l2(got.float(), want) <= 4e-3
    assert not torch.allclose(got, plain, atol=1e-3)  # the image prompt contributes


@pytest.mark.parametrize("D,L", [(40, 1024), (80, 512)])
def test_compact_region_map_edge_cases(D, L, dsc_config):
    """The compact form (weighted key columns only) with the maximum of 16 columns incl. the first and the last key,
    in arbitrary positions; 17 columns have no compact form and take the dense map."""
    dsc, att = _dsc()
    dsc_config("xattn_impl", "tc5")
    dsc_config("no_fused", "1")
    B, S = 2, 77
    q, k, v = make_qkv(B, 8, L, D, S, seed=D + L, device="cuda")
    g = torch.Generator().manual_seed(5)
    cols = sorted({0, 76} | set(torch.randperm(75, generator=g)[:14].add(1).tolist()))
    assert len(cols) == 16
    W = torch.zeros(B, L, S)
    for j, c in enumerate(cols):
        W[:, (j * 37) % L :: 3, c] = 0.1 * (j + 1) * (-1) ** j
    W = W.cuda()
    comp = att.compact_region_map(W)
    assert comp is not None and comp[1] == cols and comp[0].shape == (B, L, 20)
    for sigma in (11.0, 0.0):
        out = dsc.region_attention(q, k, v, W, sigma, compact=comp)
        ref = _oracle(q, k, v, W, sigma)
        assert rel_l2(out.float(), ref) <= TOL
        dense = dsc.region_attention(q, k, v, W, sigma)
        assert rel_l2(out.float(), dense.float()) <= 3e-4  # same math, keys visited in another order
    W17 = W.clone()
    W17[:, 5, [c for c in range(77) if c not in cols][0]] = 1.0
    assert att.compact_region_map(W17) is None
    with pytest.raises(ValueError):
        dsc.region_attention(q, k, v, W, 1.0, compact=(comp[0][:, :, :16].contiguous(), cols))
