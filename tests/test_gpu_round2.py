"""Parity holes closed in round 2: the baddbmm-variant processor against outputs of the
reference class, the fused sampler step against the diffusers (VP) form, non-uint8 region maps, captured == eager bits
for our kernels, and latent gates on the configurations bench.py actually runs (batch 8 + CUDA graph + channels_last;
BASELINE configs[2]: 768^2, 4 regions, S' > 0, batch 4)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import attention as oa
from oracle import region_map as orm
from oracle import sampler as osm

from .helpers import (NEG_IDS, PROMPT_IDS, VOCAB, AttnModule, StubTokenizer, baddbmm_fixture, ellipse_map, make_qkv, rect_map,
                      rel_l2, synthetic_w, two_rect_state, weight_func)

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


# ---------------------------------------------------------------------------------------------- a-4
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "proc_baddbmm_*.npz"))))
def test_baddbmm_processor_matches_reference_class_output(path):
    """``RegionAttnProcessorBaddbmm`` (fp16, CUDA kernels) against the fp32 output the UNMODIFIED reference
    ``AttnProcessor`` (attention_modify.py:107-207) produced for the same module and inputs (scripts/gen_golden.py), and
    against the oracle's restatement of that class (pinned bit-identical in tests/test_oracle_attention.py).  The fp16
    projections are PyTorch's and add their own rounding: gate 4e-3, as for the SDPA-style processor."""
    from diffusionspatialcontrol_b200 import RegionAttnProcessor
    from diffusionspatialcontrol_b200.attention_processor import RegionAttnProcessorBaddbmm

    attn32, hs, ctx, rp, want = baddbmm_fixture(path)
    attn16 = AttnModule(hs.shape[-1], attn32.heads, round(attn32.scale ** -2))
    attn16.load_state_dict(attn32.state_dict())
    attn16 = attn16.cuda().half()
    rp_dev = {**rp, "sigma": rp["sigma"].cuda()}
    with torch.no_grad():
        got = RegionAttnProcessorBaddbmm()(attn16, hs.cuda().half(), encoder_hidden_states=ctx.cuda().half(), region_prompt=rp_dev)
        twin = RegionAttnProcessor()(attn16, hs.cuda().half(), encoder_hidden_states=ctx.cuda().half(), region_prompt=rp_dev)
        orc = oa.processor_forward_baddbmm(attn32.cuda(), hs.cuda(), ctx.cuda(), {**rp, "region_state": {
            L: w.cuda() for L, w in rp["region_state"].items()}})
    assert rel_l2(got.float(), want) <= 4e-3
    assert rel_l2(got.float(), orc) <= 4e-3
    assert torch.equal(got, twin)  # attn.scale == head_dim ** -0.5 here: both processors run the same call


def test_baddbmm_processor_uses_attn_scale():
    """The baddbmm variant scales Q K^T by ``attn.scale`` (attention_modify.py:58-64), the SDPA-style one by
    1/sqrt(head_dim) whatever the module says (:77): a module with another scale tells them apart."""
    from diffusionspatialcontrol_b200 import RegionAttnProcessor
    from diffusionspatialcontrol_b200.attention_processor import RegionAttnProcessorBaddbmm

    torch.manual_seed(4)
    for C, H, D, L in ((320, 8, 40, 384), (640, 8, 80, 128)):
        attn32 = AttnModule(C, H, D)
        attn32.scale = 0.7 * D ** -0.5
        attn16 = AttnModule(C, H, D)
        attn16.load_state_dict(attn32.state_dict())
        attn16.scale = attn32.scale
        attn16 = attn16.cuda().half()
        attn32.load_state_dict({k: v.float() for k, v in attn16.state_dict().items()})
        attn32 = attn32.cuda()
        hs, ctx = torch.randn(2, L, C, device="cuda").half(), torch.randn(2, 77, 768, device="cuda").half()
        rp = {"region_state": {L: synthetic_w(2, L, 77).cuda()}, "sigma": torch.tensor(6.0, device="cuda"), "weight_func": weight_func}
        with torch.no_grad():
            got = RegionAttnProcessorBaddbmm()(attn16, hs, encoder_hidden_states=ctx, region_prompt=rp)
            sdpa = RegionAttnProcessor()(attn16, hs, encoder_hidden_states=ctx, region_prompt=rp)
            want = oa.processor_forward_baddbmm(attn32, hs.float(), ctx.float(), rp)
            want_sdpa = oa.processor_forward(attn32, hs.float(), ctx.float(), rp)
        assert rel_l2(got.float(), want) <= 4e-3 and rel_l2(sdpa.float(), want_sdpa) <= 4e-3
        assert rel_l2(got.float(), want_sdpa) > 2e-2  # the scale really differs


# ---------------------------------------------------------------------------------------------- a-7
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_fused_step_matches_diffusers_vp_form(dtype):
    """25 fused steps (dsc_dpmpp2m_step) against an INDEPENDENT fp64 implementation of the same sampler in the diffusers
    formulation (DPMSolverMultistepScheduler: dpmsolver++, order 2, midpoint, karras sigmas, VP variables;
    oracle/sampler.py::sample_dpmpp_2m_vp) -- the k-diffusion (VE) form is covered in test_gpu_region_sampler.py."""
    from diffusionspatialcontrol_b200.sampler import KarrasSchedule, dpmpp2m_step

    torch.manual_seed(0)
    n, shape, g = 3, (3, 4, 16, 16), 7.5
    sched = KarrasSchedule(25)
    sig = sched.sigma_list()
    noise = torch.randn(shape, dtype=torch.float64)
    A = torch.randn(16, 16, dtype=torch.float64) * 0.05

    def eps_fn(x_in, sigma):
        e = torch.tanh(x_in @ A)
        e[n:] = e[n:] * 1.1 + 0.05
        return e

    sigmas64 = sched.sigmas.double()
    model = lambda x_, s: osm.cfg_denoise(eps_fn, x_, s, g)
    want = osm.sample_dpmpp_2m_vp(model, noise * (sigmas64[0] ** 2 + 1) ** 0.5, sigmas64)
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


# ---------------------------------------------------------------------------------------------- a-5
def test_non_uint8_region_maps_are_binarised_in_their_own_dtype():
    """The reference tests ``map < 255`` in the map's own dtype (encode_region_map_function.py:49): 300 is outside, -1 and
    254.5 are inside.  A uint8 cast before the test would flip all three."""
    from types import SimpleNamespace

    from diffusionspatialcontrol_b200 import encode_region_map

    pipe = SimpleNamespace(tokenizer=StubTokenizer(), unet=SimpleNamespace(down_blocks=[0] * 4), vae_scale_factor=8,
                           do_classifier_free_guidance=True)
    ids, neg = np.array([PROMPT_IDS]), np.array([NEG_IDS])
    m_int = np.full((512, 512), 300, np.int32)
    m_int[100:300, 64:256] = -1
    m_flt = np.full((512, 512), 255.0, np.float32)
    m_flt[256:448, 200:456] = 254.5
    m_flt[0:64, 0:64] = 511.0  # wraps to 255 as uint8 (outside), but 511 -> 255 only by luck; 300 -> 44 would be inside
    state = {"A girl": {"map": m_int, "weight": 0.5, "mask_outsides": 0.1},
             "bridge": {"map": m_flt, "weight": 0.7, "mask_outsides": 0.0}}
    tok = lambda phrase: StubTokenizer()(phrase).input_ids  # noqa: E731
    got = encode_region_map(pipe, state, 512, 512, 1, text_ids=[neg, ids])
    want = orm.encode_region_map(state, tok, 512, 512, 1, text_ids=[neg, ids])
    for L in want:
        assert torch.equal(got[L].cpu(), want[L]), L
    assert float(got[4096].abs().sum()) > 0


# ---------------------------------------------------------------------------------------------- f-3
def test_attention_call_captured_in_a_cuda_graph_equals_eager_bits():
    """Our kernels are bit-reproducible under stream capture: the same attention call (both launches, programmatic
    dependent launch edge included) replayed from a CUDA graph returns exactly the eager bits, on both kernel families.
    (The whole-pipeline graph test gates on cosine instead: cuBLAS / cuDNN pick other algorithms under capture.)"""
    from diffusionspatialcontrol_b200 import attention as att

    for (B, L, D) in ((4, 4096, 40), (4, 1024, 80), (2, 256, 160)):
        q, k, v = make_qkv(B, 8, L, D, 77, seed=L, device="cuda")
        W = att.padded_region_map(synthetic_w(2, L, 77).cuda())
        compact = att.compact_region_map(W)
        sigma = torch.tensor(5.0, device="cuda")
        ws = torch.zeros(att.workspace_bytes(B, 8, L, D, 77), dtype=torch.uint8, device="cuda")
        prepared = att.prepared_supported(8, D, 77, len(compact[1]))
        kv = att.prepare_kv(k, v, compact[1]) if prepared else None
        out = torch.empty(B, L, 8 * D, device="cuda", dtype=torch.float16)

        def call():
            if prepared:
                return att.region_attention_prepared(q, kv, compact, sigma, workspace=ws, out=out)
            return att.region_attention(q, k, v, W, sigma, workspace=ws, compact=compact)

        eager = call().clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            call()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            res = call()
        for _ in range(3):
            res.zero_()
            g.replay()
            assert torch.equal(res, eager), (B, L, D)


# ---------------------------------------------------------------------------------------------- pipeline gates
def _embeds():
    g1, g2 = torch.Generator().manual_seed(1), torch.Generator().manual_seed(2)
    return torch.randn(1, 77, 768, generator=g1), torch.randn(1, 77, 768, generator=g2)


def _reference_latents(unet32, state, ids, n, height, width, noise, cond, uncond, dev):
    """fp32 UNet + restated reference processor + restated k-diffusion loop (all on the GPU for speed)."""
    unet32.set_attn_processor(oa.OracleAttnProcessor())
    rs = orm.encode_region_map(state, lambda p: VOCAB[p], width, height, n, text_ids=ids)
    rs = {L: t.to(dev) for L, t in rs.items()}
    ctx = torch.cat([uncond.expand(n, -1, -1), cond.expand(n, -1, -1)]).to(dev)
    train = osm.sd15_train_sigmas().to(dev)

    def eps_fn(x_in, sigma):
        t = osm.sigma_to_t(sigma.reshape(1), train.log())
        rp = {"region_state": rs, "sigma": sigma, "weight_func": oa.weight_func}
        return unet32(x_in, t, ctx, cross_attention_kwargs={"region_prompt": rp})

    with torch.no_grad():
        return osm.txt2img_latents(eps_fn, noise.to(dev), steps=25, guidance=7.5).cpu()


def _unets(dev):
    from diffusionspatialcontrol_b200.unet_sd15 import UNetSD15

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    unet32 = UNetSD15().eval()
    unet16 = UNetSD15().eval()
    unet16.load_state_dict(unet32.state_dict())
    return unet16.to(dev, torch.float16), unet32.to(dev)


def test_bench_configuration_latent_cosine_batch8_graph_channels_last():
    """The configuration bench.py times (BASELINE configs[1]: batch 8 -> attention batch 16, one captured CUDA graph per
    denoising step, channels_last host UNet, cudnn.benchmark, K/V hoisted out of the graph, prepared K/V images) against
    the fp32 reference loop: final-latent cosine >= 0.999 at a fixed seed."""
    from diffusionspatialcontrol_b200.distributed import unit_noise
    from diffusionspatialcontrol_b200.pipeline import RegionTxt2ImgPipeline, SyntheticTokenizer

    dev = torch.device("cuda")
    unet16, unet32 = _unets(dev)
    old = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    try:
        unet16 = unet16.to(memory_format=torch.channels_last)
        cond, uncond = _embeds()
        ids = [np.array([NEG_IDS]), np.array([PROMPT_IDS])]
        state = two_rect_state(512, 512)
        n = 8
        noise = unit_noise(0, n, (4, 64, 64))
        pipe = RegionTxt2ImgPipeline(unet16, SyntheticTokenizer(VOCAB), use_cuda_graph=True)
        with torch.no_grad():
            ours = pipe.txt2img(cond, uncond, ids, state, noise.to(dev), 512, 512, 25, 7.5).float().cpu()
            again = pipe.txt2img(cond, uncond, ids, state, noise.to(dev), 512, 512, 25, 7.5).float().cpu()
        assert torch.equal(ours, again)  # replays of one graph: same bits
    finally:
        torch.backends.cudnn.benchmark = old
    ref = _reference_latents(unet32, state, ids, n, 512, 512, noise, cond, uncond, dev)
    assert torch.isfinite(ours).all()
    cos = torch.nn.functional.cosine_similarity(ours.flatten(), ref.flatten(), dim=0)
    assert cos >= 0.999, f"final-latent cosine {cos:.6f}"
    per_image = torch.nn.functional.cosine_similarity(ours.flatten(1), ref.flatten(1), dim=1)
    assert per_image.min() >= 0.998, per_image


def test_config3_768_four_regions_suppression_batch4_end_to_end():
    """BASELINE configs[2]: 768 x 768 (96 x 96 latent, 9216 / 2304 / 576 / 144 queries), 4 regions with S' > 0, batch 4,
    25 steps, eager and as a CUDA graph, against the fp32 reference loop."""
    from diffusionspatialcontrol_b200.distributed import unit_noise
    from diffusionspatialcontrol_b200.pipeline import RegionTxt2ImgPipeline, SyntheticTokenizer

    dev = torch.device("cuda")
    unet16, unet32 = _unets(dev)
    cond, uncond = _embeds()
    ids = [np.array([NEG_IDS]), np.array([PROMPT_IDS])]
    state = {
        "A girl": {"map": rect_map(768, 768, 30, 400, 40, 350), "weight": 0.5, "mask_outsides": 0.2},
        "bridge": {"map": ellipse_map(768, 768, 500, 520, 160, 210), "weight": 0.7, "mask_outsides": 0.0},
        "sitting": {"map": rect_map(768, 768, 420, 700, 400, 740), "weight": 0.4, "mask_outsides": 0.1},
        "on the": {"map": ellipse_map(768, 768, 200, 600, 120, 90), "weight": 1.0, "mask_outsides": 0.3},
    }
    n = 4
    noise = unit_noise(5, n, (4, 96, 96))
    with torch.no_grad():
        eager = RegionTxt2ImgPipeline(unet16, SyntheticTokenizer(VOCAB)).txt2img(
            cond, uncond, ids, state, noise.to(dev), 768, 768, 25, 7.5).float().cpu()
        graph = RegionTxt2ImgPipeline(unet16, SyntheticTokenizer(VOCAB), use_cuda_graph=True).txt2img(
            cond, uncond, ids, state, noise.to(dev), 768, 768, 25, 7.5).float().cpu()
    ref = _reference_latents(unet32, state, ids, n, 768, 768, noise, cond, uncond, dev)
    cos = lambda a, b: float(torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0))  # noqa: E731
    assert torch.isfinite(eager).all() and torch.isfinite(graph).all()
    assert cos(eager, ref) >= 0.999, cos(eager, ref)
    assert cos(graph, ref) >= 0.999, cos(graph, ref)
    # negative weights really reach the kernels: the S' terms change the result
    no_sup = {k: {**v, "mask_outsides": 0.0} for k, v in state.items()}
    with torch.no_grad():
        other = RegionTxt2ImgPipeline(unet16, SyntheticTokenizer(VOCAB)).txt2img(
            cond, uncond, ids, no_sup, noise.to(dev), 768, 768, 25, 7.5).float().cpu()
    assert cos(other, eager) < 0.99999
