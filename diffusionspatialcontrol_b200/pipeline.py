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

    def __init__(self, unet, tokenizer, processor: Optional[RegionAttnProcessor] = None):
        self.unet = unet
        self.tokenizer = tokenizer
        self.processor = processor if processor is not None else RegionAttnProcessor(cache_kv=True)
        self.unet.set_attn_processor(self.processor)  # same hook as reference app.py:479-481

    @property
    def device(self):
        return next(self.unet.parameters()).device

    @property
    def dtype(self):
        return next(self.unet.parameters()).dtype

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
