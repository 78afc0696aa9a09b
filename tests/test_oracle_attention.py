"""The oracle restatement of the attention path is pinned (a) against the UNMODIFIED reference module
executed in this container and (b) against the committed golden vectors (which were produced by that
module, scripts/gen_golden.py) so the pin travels to machines without /root/reference."""
import glob
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import attention as oa
from oracle import ref_loader

from .helpers import AttnModule, baddbmm_fixture, make_qkv, synthetic_w, weight_func

needs_ref = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "attn_*.npz"))))
def test_oracle_matches_golden(path):
    z = np.load(path)
    H = int(z["heads"])
    q, k, v = (torch.from_numpy(z[n]).float() for n in "qkv")
    B, L, HD = q.shape
    D = HD // H
    view = lambda t: t.view(B, -1, H, D).transpose(1, 2)
    out = oa.region_attention(view(q), view(k), view(v), torch.from_numpy(z["W"]).clone(), torch.tensor(float(z["sigma"])))
    out = out.transpose(1, 2).reshape(B, L, HD)
    assert torch.equal(out, torch.from_numpy(z["out"]))  # same ops in the same order: bit-identical
    assert float(oa.score_std(view(q), view(k))) == pytest.approx(float(z["std"]), rel=0, abs=0)


@needs_ref
@pytest.mark.parametrize("B,H,L,D,S,Bw,sigma", [(2, 8, 64, 160, 77, 2, 14.6), (4, 8, 256, 40, 77, 2, 0.5), (2, 4, 40, 80, 77, 1, 3.0)])
def test_oracle_matches_reference_function(B, H, L, D, S, Bw, sigma):
    ref = ref_loader.attention_modify()
    q, k, v = (t.float() for t in make_qkv(B, H, L, D, S, seed=7))
    W = synthetic_w(Bw, L, S)
    a = ref.scaled_dot_product_attention_regionstate(q, k, v, weight_func=weight_func, region_state=W.clone(), sigma=torch.tensor(sigma))
    b = oa.region_attention(q, k, v, W.clone(), torch.tensor(sigma))
    assert torch.equal(a, b)


class _Attn(nn.Module):
    def __init__(self, C, H, D, ctx=768):
        super().__init__()
        self.heads, self.scale = H, D**-0.5
        self.to_q, self.to_k, self.to_v = nn.Linear(C, H * D, bias=False), nn.Linear(ctx, H * D, bias=False), nn.Linear(ctx, H * D, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(H * D, C), nn.Dropout(0.0)])
        self.spatial_norm = self.group_norm = self.norm_cross = None
        self.residual_connection, self.rescale_output_factor = False, 1.0


@needs_ref
def test_oracle_processor_matches_reference_processor():
    ref = ref_loader.attention_modify()
    torch.manual_seed(3)
    attn = _Attn(320, 8, 40)
    hs, ctx = torch.randn(2, 256, 320), torch.randn(2, 77, 768)
    rp = {"region_state": {256: synthetic_w(2, 256, 77)}, "sigma": torch.tensor(5.0), "weight_func": weight_func}
    with torch.no_grad():
        a = ref.AttnProcessor2_0()(attn, hs, encoder_hidden_states=ctx, region_prompt=rp)
        b = oa.processor_forward(attn, hs, ctx, rp)
        a_self = ref.AttnProcessor2_0()(attn_self := _Attn(320, 8, 40, ctx=320), hs, region_prompt=rp)
        b_self = oa.processor_forward(attn_self, hs, None, rp)
    assert torch.equal(a, b) and torch.equal(a_self, b_self)


def test_weight_func_is_whole_tensor_unbiased_std():
    qk = torch.randn(3, 5, 7)
    w = torch.ones(1, 5, 7)
    n = qk.numel()
    want = ((qk - qk.mean()) ** 2).sum().div(n - 1).sqrt() * 2.0
    assert torch.allclose(oa.weight_func(w, 2.0, qk), want * w)


def test_batch_coupling_is_part_of_the_contract():
    """std spans the batch: the same sample gives a different output when its batch neighbours change."""
    q, k, v = (t.float() for t in make_qkv(2, 8, 64, 40, 77, seed=11))
    W = synthetic_w(1, 64, 77)
    full = oa.region_attention(q, k, v, W.clone(), 10.0)
    q2 = q.clone()
    q2[1] *= 3.0
    other = oa.region_attention(q2, k, v, W.clone(), 10.0)
    assert not torch.allclose(full[0], other[0])


class _AttnFull(_Attn):
    """adds what the baddbmm variant touches (attention_modify.py:160-191, :41-68)"""

    upcast_attention = False
    upcast_softmax = False

    def head_to_batch_dim(self, t, out_dim=3):
        b, n, c = t.shape
        t = t.reshape(b, n, self.heads, c // self.heads).permute(0, 2, 1, 3)
        return t.reshape(b * self.heads, n, c // self.heads) if out_dim == 3 else t

    def batch_to_head_dim(self, t):
        bh, n, d = t.shape
        return t.reshape(bh // self.heads, self.heads, n, d).permute(0, 2, 1, 3).reshape(bh // self.heads, n, d * self.heads)

    def prepare_attention_mask(self, m, *_a, **_k):
        return m


@needs_ref
def test_reference_baddbmm_processor_equals_sdpa_style_processor():
    """AttnProcessor (:107-207) and AttnProcessor2_0 (:414-503) agree on the region path, so one
    restatement (and one pair of kernels) covers both."""
    ref = ref_loader.attention_modify()
    torch.manual_seed(5)
    attn = _AttnFull(320, 8, 40)
    hs, ctx = torch.randn(2, 64, 320), torch.randn(2, 77, 768)
    rp = {"region_state": {64: synthetic_w(2, 64, 77)}, "sigma": torch.tensor(3.0), "weight_func": weight_func}
    with torch.no_grad():
        a = ref.AttnProcessor()(attn, hs, encoder_hidden_states=ctx, region_prompt=rp)
        b = ref.AttnProcessor2_0()(attn, hs, encoder_hidden_states=ctx, region_prompt=rp)
        c = oa.processor_forward(attn, hs, ctx, rp)
    assert torch.allclose(a, b, atol=2e-6, rtol=1e-5) and torch.equal(b, c)


@needs_ref
def test_oracle_baddbmm_processor_matches_reference_processor():
    """the oracle's restatement of ``AttnProcessor`` (attention_modify.py:107-207, :39-70) is bit-identical
    to the unmodified reference class."""
    ref = ref_loader.attention_modify()
    torch.manual_seed(8)
    for C, H, D, L in ((320, 8, 40, 96), (640, 8, 80, 40)):
        attn = AttnModule(C, H, D)
        hs, ctx = torch.randn(2, L, C), torch.randn(2, 77, 768)
        rp = {"region_state": {L: synthetic_w(2, L, 77)}, "sigma": torch.tensor(3.0), "weight_func": weight_func}
        with torch.no_grad():
            a = ref.AttnProcessor()(attn, hs, encoder_hidden_states=ctx, region_prompt=rp)
            b = oa.processor_forward_baddbmm(attn, hs, ctx, rp)
            a_self = ref.AttnProcessor()(attn_self := AttnModule(C, H, D, ctx=C), hs, region_prompt=rp)
            b_self = oa.processor_forward_baddbmm(attn_self, hs, None, rp)
        assert torch.equal(a, b) and torch.equal(a_self, b_self)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "proc_baddbmm_*.npz"))))
def test_oracle_baddbmm_processor_matches_golden(path):
    """... and to the outputs the reference class produced (scripts/gen_golden.py), which travel to the GPU box."""
    attn, hs, ctx, rp, want = baddbmm_fixture(path)
    with torch.no_grad():
        got = oa.processor_forward_baddbmm(attn, hs, ctx, rp)
    assert torch.allclose(got, want, atol=1e-6, rtol=1e-5)


# ------------------------------------------------------------------ additive attention masks (a = Q K^T + M)
@needs_ref
def test_reference_function_mask_semantics_are_what_the_product_mirrors():
    """attention_modify.py:84-89 on the unmodified module: a BOOL mask is never applied (:86-87 only rewrites the mask
    tensor), a float mask is added in place into a [L, S] bias (:89) -- fine for masks that broadcast INTO [L, S], a
    RuntimeError for the 4-D tensor the processor builds (:452).  The oracle restates exactly that."""
    ref = ref_loader.attention_modify()
    f = ref.scaled_dot_product_attention_regionstate
    B, H, L, D, S = 2, 8, 64, 40, 77
    q, k, v = (t.float() for t in make_qkv(B, H, L, D, S, seed=21))
    W, sig = synthetic_w(B, L, S), torch.tensor(3.0)
    base = f(q, k, v, weight_func=weight_func, region_state=W.clone(), sigma=sig)
    mb = torch.ones(B, H, 1, S, dtype=torch.bool)
    mb[..., 40:] = False
    assert torch.equal(f(q, k, v, attn_mask=mb.clone(), weight_func=weight_func, region_state=W.clone(), sigma=sig), base)
    assert torch.equal(oa.region_attention(q, k, v, W.clone(), sig, attn_mask=mb.clone()), base)
    for shape in ((B, H, 1, S), (B, H, L, S), (1, 1, 1, S)):
        with pytest.raises(RuntimeError):
            f(q, k, v, attn_mask=torch.zeros(shape), weight_func=weight_func, region_state=W.clone(), sigma=sig)
        with pytest.raises(RuntimeError):
            oa.region_attention(q, k, v, W.clone(), sig, attn_mask=torch.zeros(shape))
    g = torch.Generator().manual_seed(2)
    for shape in ((L, S), (1, S), (S,)):
        m = torch.randn(shape, generator=g)
        m[..., 60:] -= 5.0
        a = f(q, k, v, attn_mask=m.clone(), weight_func=weight_func, region_state=W.clone(), sigma=sig)
        b = oa.region_attention(q, k, v, W.clone(), sig, attn_mask=m.clone())
        assert torch.equal(a, b) and not torch.allclose(a, base, atol=1e-3)


@needs_ref
def test_reference_processors_with_a_mask():
    """The SDPA-style processor (:414-503) ignores a bool mask and raises for a float one on the region path; the baddbmm
    processor (:107-207) adds its mask (baddbmm input, beta = 1) and the oracle's restatement is bit-identical to it."""
    ref = ref_loader.attention_modify()
    torch.manual_seed(12)
    C, H, D, L, B = 320, 8, 40, 96, 2
    attn = AttnModule(C, H, D)
    hs, ctx = torch.randn(B, L, C), torch.randn(B, 77, 768)
    rp = {"region_state": {L: synthetic_w(B, L, 77)}, "sigma": torch.tensor(3.0), "weight_func": weight_func}
    m = torch.randn(B * H, 1, 77) * 0.5
    m[:, :, 50:] -= 6.0
    with torch.no_grad():
        base = ref.AttnProcessor2_0()(attn, hs, encoder_hidden_states=ctx, region_prompt=rp)
        mb = torch.ones(B * H, 1, 77, dtype=torch.bool)
        mb[..., 30:] = False
        assert torch.equal(ref.AttnProcessor2_0()(attn, hs, encoder_hidden_states=ctx, attention_mask=mb.clone(), region_prompt=rp), base)
        assert torch.equal(oa.processor_forward(attn, hs, ctx, rp, attention_mask=mb.clone()), base)
        with pytest.raises(RuntimeError):
            ref.AttnProcessor2_0()(attn, hs, encoder_hidden_states=ctx, attention_mask=m.clone(), region_prompt=rp)
        with pytest.raises(RuntimeError):
            oa.processor_forward(attn, hs, ctx, rp, attention_mask=m.clone())
        for mask in (m, m.expand(B * H, L, 77).contiguous() + torch.randn(B * H, L, 77) * 0.1):
            a = ref.AttnProcessor()(attn, hs, encoder_hidden_states=ctx, attention_mask=mask.clone(), region_prompt=rp)
            b = oa.processor_forward_baddbmm(attn, hs, ctx, rp, attention_mask=mask.clone())
            assert torch.equal(a, b) and not torch.allclose(a, base, atol=1e-3)
        # a -inf entry: the std of the masked scores is NaN and the beta term poisons every row, in the reference too
        m_inf = m.clone()
        m_inf[:, :, 70:] = float("-inf")
        a = ref.AttnProcessor()(attn, hs, encoder_hidden_states=ctx, attention_mask=m_inf.clone(), region_prompt=rp)
        assert torch.isnan(a).all() and torch.isnan(oa.processor_forward_baddbmm(attn, hs, ctx, rp, attention_mask=m_inf)).all()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "procm_baddbmm_*.npz"))))
def test_oracle_baddbmm_processor_with_mask_matches_golden(path):
    from .helpers import baddbmm_mask_fixture

    attn, hs, ctx, rp, want, mask = baddbmm_mask_fixture(path)
    with torch.no_grad():
        got = oa.processor_forward_baddbmm(attn, hs, ctx, rp, attention_mask=mask)
        unmasked = oa.processor_forward_baddbmm(attn, hs, ctx, rp)
    assert torch.allclose(got, want, atol=1e-6, rtol=1e-5)
    assert not torch.allclose(unmasked, want, atol=1e-3)  # the mask matters in this fixture
