"""DPM++ 2M Karras sampling around the hot path (companion K4).

Host-side mirror of what the reference runs per generation:
  sigma schedule   reference source/modules/model_k_diffusion.py:848-859 (k_diffusion.get_sigmas_karras)
  sigma -> t       reference source/modules/external_k_diffusion.py:65-77
  per-step update  k_diffusion.sampling.sample_dpmpp_2m + the denoiser scalings / CFG around it
                   (reference source/modules/external_k_diffusion.py:95-114, model_k_diffusion.py:1162-1166)
The per-step elementwise work is ONE fused CUDA kernel (dsc_dpmpp2m_step); schedules are tiny host math.
"""
from __future__ import annotations

import ctypes
import math
from typing import List, Optional

import torch

from . import _lib
from ._lib import check, lib

_DTYPES = {torch.float16: _lib.DTYPE_F16, torch.bfloat16: _lib.DTYPE_BF16}


def sd15_train_sigmas() -> torch.Tensor:
    """sqrt((1-abar)/abar) of SD-1.5's scaled-linear schedule (what DiscreteEpsDDPMDenoiser stores)."""
    betas = torch.linspace(0.00085**0.5, 0.012**0.5, 1000, dtype=torch.float32) ** 2
    abar = torch.cumprod(1.0 - betas, dim=0)
    return ((1 - abar) / abar) ** 0.5


def get_sigmas_karras(n: int, sigma_min: float, sigma_max: float, rho: float = 7.0) -> torch.Tensor:
    ramp = torch.linspace(0, 1, n)
    lo, hi = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    sig = (hi + ramp * (lo - hi)) ** rho
    return torch.cat([sig, sig.new_zeros([1])])


def sigma_to_t(sigma: torch.Tensor, log_sigmas: torch.Tensor) -> torch.Tensor:
    log_sigma = sigma.log()
    dists = log_sigma - log_sigmas[:, None]
    low_idx = dists.ge(0).cumsum(dim=0).argmax(dim=0).clamp(max=log_sigmas.shape[0] - 2)
    high_idx = low_idx + 1
    low, high = log_sigmas[low_idx], log_sigmas[high_idx]
    w = ((low - log_sigma) / (low - high)).clamp(0, 1)
    return ((1 - w) * low_idx + w * high_idx).view(sigma.shape)


@_lib.nvtx("dpmpp2m_step")
def dpmpp2m_step(x: torch.Tensor, eps_uc: torch.Tensor, den_prev: torch.Tensor,
                 unet_in_next: Optional[torch.Tensor], sigma_prev: float, sigma: float, sigma_next: float,
                 cfg: float, first: bool) -> None:
    """In-place fused step.  x, den_prev: fp32 [n,...]; eps_uc: fp16/bf16 [2n,...] (uncond rows first);
    unet_in_next: fp16/bf16 [2n,...] or None."""
    if not x.is_cuda:
        raise RuntimeError("diffusionspatialcontrol_b200 has no CPU path")
    if x.dtype != torch.float32 or den_prev.dtype != torch.float32:
        raise TypeError("x and den_prev are kept in float32")
    if eps_uc.dtype not in _DTYPES:
        raise TypeError("eps_uc must be float16 or bfloat16")
    n = x.numel()
    if eps_uc.numel() != 2 * n or den_prev.numel() != n:
        raise ValueError("shape mismatch: eps_uc must hold 2x the elements of x")
    if not (x.is_contiguous() and eps_uc.is_contiguous() and den_prev.is_contiguous()):
        raise ValueError("tensors must be contiguous")
    nxt = None
    if unet_in_next is not None:
        if unet_in_next.dtype != eps_uc.dtype or unet_in_next.numel() != 2 * n or not unet_in_next.is_contiguous():
            raise ValueError("unet_in_next must match eps_uc in dtype/size and be contiguous")
        nxt = unet_in_next.data_ptr()
    with torch.cuda.device(x.device):
        check(lib.dsc_dpmpp2m_step(x.data_ptr(), eps_uc.data_ptr(), den_prev.data_ptr(), nxt, n,
                                   float(sigma_prev), float(sigma), float(sigma_next), float(cfg), int(bool(first)),
                                   _DTYPES[eps_uc.dtype],
                                   ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))


class KarrasSchedule:
    """25-step (default) Karras schedule of the SD-1.5 denoiser, with the fractional timesteps."""

    def __init__(self, steps: int = 25):
        train = sd15_train_sigmas()
        self.sigmas = get_sigmas_karras(steps, train[0].item(), train[-1].item())  # fp32, steps+1 (last = 0)
        self.timesteps = sigma_to_t(self.sigmas[:-1], train.log())
        self.steps = steps

    def sigma_list(self) -> List[float]:
        return [float(s) for s in self.sigmas]

    def c_in(self, i: int) -> float:
        s = float(self.sigmas[i])
        return 1.0 / math.sqrt(s * s + 1.0)
