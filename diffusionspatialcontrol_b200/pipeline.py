"""The 25-step txt2img latent loop around the hot path, restated for measurement
(reference source/modules/model_k_diffusion.py:1027-1175, k-diffusion formulation; the callers are
out of scope as a port -- SURVEY.md 2).  Latents only: no text encoder, no VAE.

What is ours here: region maps built on the device (region_map.py), the attention processor
(attention_processor.py), the fused sampler step (sampler.py).  The UNet is ordinary PyTorch.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional

import torch

from .attention import compact_region_map, padded_region_map, prepare_kv, prepared_supported, workspace_bytes
from .attention_processor import RegionAttnProcessor
from .region_map import encode_region_map
from .sampler import KarrasSchedule, dpmpp2m_step

# the reference's weight_func (source/app.py:1004); the processor verifies the callable it is handed
reference_weight_func = lambda w, sigma, qk: w * sigma * qk.std()  # noqa: E731


class SyntheticTokenizer:
    """Stands in for the CLIP tokenizer (no vocab files offline): phrase -> fixed ids."""

    model_max_length = 77

    def __init__(self, vocab: Dict[str, List[int]]):
        self.vocab = {k: list(v) for k, v in vocab.items()}

    def __call__(self, text, max_length=None, truncation=True, add_special_tokens=False, **_):
        return SimpleNamespace(input_ids=list(self.vocab[text]))


class RegionTxt2ImgPipeline:
    """Minimal stand-in for the reference pipelines' txt2img: same inputs that reach the hot path
    (prompt embeddings, token ids, region_map_state, weight_func), same per-step region_prompt dict
    (model_k_diffusion.py:1098-1103), DPM++ 2M Karras, CFG."""

    vae_scale_factor = 8
    do_classifier_free_guidance = True

    def __init__(self, unet, tokenizer, processor: Optional[RegionAttnProcessor] = None, use_cuda_graph: bool = False):
        self.unet = unet
        self.tokenizer = tokenizer
        # K / V are projections of the text embeddings: computed once per generation (eager: the processor's cache;
        # CUDA graph: static buffers refreshed outside the graph, see _refresh_static_kv), never once per step
        self.processor = processor if processor is not None else RegionAttnProcessor(cache_kv=True)
        self.unet.set_attn_processor(self.processor)  # same hook as reference app.py:479-481
        self.use_cuda_graph = use_cuda_graph
        self._graphs: Dict[tuple, dict] = {}

    @property
    def device(self):
        return next(self.unet.parameters()).device

    @property
    def dtype(self):
        return next(self.unet.parameters()).dtype

    # ---- one UNet evaluation, optionally as a captured CUDA graph -------------------------------------------
    def _graph_state(self, n: int, height: int, width: int, region_state, weight_func) -> dict:
        """Static buffers + captured graph of ONE denoising-step UNet call (all 32 attention layers, our two passes
        per cross-attention layer included) for a given batch shape.  Launch-bound glue (hundreds of small PyTorch
        kernels per step) is replayed with one graph launch; our kernels are capture-safe (no host sync, workspace
        pre-allocated, tensor maps passed by value)."""
        # The region maps live in static buffers in the layout the kernels read (padded rows + compact form); what a
        # captured graph bakes in -- which key columns carry weights, or that a map is all zero (plain SDPA) -- is part
        # of the graph's key, so another prompt layout captures another graph instead of replaying a wrong one.
        compacts = {L: compact_region_map(w) for L, w in region_state.items()}
        sig = tuple((L, tuple(w.shape), None if compacts[L] is None else tuple(compacts[L][1]))
                    for L, w in sorted(region_state.items()))
        key = (n, height, width, self.dtype, sig)
        st = self._graphs.get(key)
        if st is not None:
            st["compacts"] = compacts
            return st
        dev, dt = self.device, self.dtype
        st = {
            "x": torch.zeros((2 * n, self.unet.in_channels, height // 8, width // 8), device=dev, dtype=dt),
            "t": torch.zeros((), device=dev, dtype=torch.float32),
            "sigma": torch.zeros((), device=dev, dtype=torch.float32),
            "ctx": torch.zeros((2 * n, 77, self.unet_ctx_dim()), device=dev, dtype=dt),
            "rs": {L: padded_region_map(torch.zeros(tuple(w.shape), device=dev, dtype=torch.float32))
                   for L, w in region_state.items()},
            "rsc": {L: (None if c is None else torch.zeros_like(c[0])) for L, c in compacts.items()},
            "compacts": compacts,
        }
        for L in region_state:
            c = compacts[L]
            self.processor.register_static_map(
                st["rs"][L], compact=None if c is None else (st["rsc"][L], list(c[1])),
                zero=c is not None and len(c[1]) == 0)
        # the 16 x 2 K / V projections (and the K / V^T images of the tcgen05 kernels) live in static buffers
        # that are refreshed once per generation OUTSIDE the graph: the captured step holds no to_k / to_v GEMM
        st["kv"] = []
        for m in self.unet.modules():
            if getattr(m, "is_cross_attention", False):
                k = torch.zeros((2 * n, 77, m.to_k.out_features), device=dev, dtype=dt)
                v = torch.zeros_like(k)
                head_dim = m.to_k.out_features // m.heads
                images = {}
                for L, c in compacts.items():
                    if c is not None and prepared_supported(m.heads, head_dim, 77, len(c[1])) and tuple(c[1]) not in images:
                        view = lambda t: t.view(2 * n, 77, m.heads, head_dim).transpose(1, 2)
                        images[tuple(c[1])] = prepare_kv(view(k), view(v), c[1])
                st["kv"].append((m, k, v, images))
                self.processor.register_static_kv(m, st["ctx"], k, v, images)
        rp = {"region_state": st["rs"], "sigma": st["sigma"], "weight_func": weight_func}
        kw = {"region_prompt": rp}
        # the statistics workspace of this graph: allocated here, outside the capture, and owned by the graph state (the
        # replays of one graph are serialised on their stream; two graphs never share a ticket word)
        st["ws"] = torch.zeros(workspace_bytes(2 * n, 8, (height // 8) * (width // 8), 40, 77), dtype=torch.uint8, device=dev)
        prev_ws, self.processor.workspace = self.processor.workspace, st["ws"]
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):  # warm-up on a side stream (allocations, autotuning)
                for _ in range(2):
                    self.unet(st["x"], st["t"], st["ctx"], cross_attention_kwargs=kw)
            torch.cuda.current_stream(dev).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                # (a channels_last UNet returns a channels_last eps: the sampler kernel wants plain NCHW order)
                st["eps"] = self.unet(st["x"], st["t"], st["ctx"], cross_attention_kwargs=kw).contiguous()
        finally:
            self.processor.workspace = prev_ws
        st["graph"] = g
        self._graphs[key] = st
        return st

    @staticmethod
    def _refresh_static_kv(st: dict) -> None:
        """Recompute the static K / V projections (and K / V^T images) from the static text-embedding buffer: once per
        generation, outside the captured graph."""
        for m, k, v, images in st["kv"]:
            k.copy_(m.to_k(st["ctx"]))
            v.copy_(m.to_v(st["ctx"]))
            head_dim = k.shape[-1] // m.heads
            view = lambda t: t.view(t.shape[0], t.shape[1], m.heads, head_dim).transpose(1, 2)
            for cols, img in images.items():
                prepare_kv(view(k), view(v), cols, out=img.image)

    def unet_ctx_dim(self) -> int:
        for m in self.unet.modules():
            if getattr(m, "is_cross_attention", False):
                return m.to_k.in_features
        return 768

    @torch.no_grad()
    def txt2img(self, prompt_embeds: torch.Tensor, negative_prompt_embeds: torch.Tensor, text_ids,
                region_map_state, noise: torch.Tensor, height: int = 512, width: int = 512,
                num_inference_steps: int = 25, guidance_scale: float = 7.5,
                weight_func=reference_weight_func, region_state=None) -> torch.Tensor:
        """noise: unit-normal [n, 4, h/8, w/8] (device, any float dtype).  Returns final latents fp32."""
        dev, dt = self.device, self.dtype
        n = noise.shape[0]
        sched = KarrasSchedule(num_inference_steps)
        sig = sched.sigma_list()
        sig_dev = sched.sigmas.to(dev)  # fp32 on the device: region_prompt["sigma"] is sig_dev[i] (no host sync)
        t_dev = sched.timesteps.to(dev)
        ctx = torch.cat([negative_prompt_embeds.expand(n, -1, -1), prompt_embeds.expand(n, -1, -1)]).to(dev, dt)
        if region_state is None:
            region_state = encode_region_map(self, region_map_state, width=width, height=height,
                                             num_images_per_prompt=n, text_ids=text_ids, device=dev)
        x = noise.to(dev, torch.float32) * (sig[0] ** 2 + 1) ** 0.5  # model_k_diffusion.py:1043
        x = x.contiguous()
        den_prev = torch.zeros_like(x)
        if self.use_cuda_graph:
            st = self._graph_state(n, height, width, region_state, weight_func)
            st["ctx"].copy_(ctx)
            self._refresh_static_kv(st)
            for L, w in region_state.items():
                st["rs"][L].copy_(w)
                if st["rsc"][L] is not None:
                    st["rsc"][L].copy_(st["compacts"][L][0])
            unet_in = st["x"]
            unet_in.copy_(torch.cat([x, x]).mul_(sched.c_in(0)))
            for i in range(num_inference_steps):
                st["t"].copy_(t_dev[i])
                st["sigma"].copy_(sig_dev[i])
                st["graph"].replay()
                last = i == num_inference_steps - 1
                # the fused step writes the next UNet input straight into the graph's static input buffer
                dpmpp2m_step(x, st["eps"], den_prev, None if last else unet_in,
                             sig[i - 1] if i > 0 else 0.0, sig[i], sig[i + 1], guidance_scale, first=(i == 0))
            return x
        unet_in = torch.cat([x, x]).mul_(sched.c_in(0)).to(dt).contiguous()
        unet_in_next = torch.empty_like(unet_in)
        for i in range(num_inference_steps):
            region_prompt = {"region_state": region_state, "sigma": sig_dev[i], "weight_func": weight_func}
            eps = self.unet(unet_in, t_dev[i], ctx, cross_attention_kwargs={"region_prompt": region_prompt})
            last = i == num_inference_steps - 1
            dpmpp2m_step(x, eps.contiguous(), den_prev, None if last else unet_in_next,
                         sig[i - 1] if i > 0 else 0.0, sig[i], sig[i + 1], guidance_scale, first=(i == 0))
            unet_in, unet_in_next = unet_in_next, unet_in
        return x
