"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the DPM++ 2M Karras sampler loop that drives the
hot path.  Never imported by the product package.

PARITY UNPINNED against third-party code: the arithmetic lives in ``k_diffusion==0.1.1.post1``
(``sampling.sample_dpmpp_2m``, ``sampling.get_sigmas_karras``, ``sampling.append_zero``) and
``diffusers==0.27.2`` (``DPMSolverMultistepScheduler``), pinned in
/root/reference/source/requirements.txt:3,5 -- neither is vendored under /root/reference nor
installed here, and the reference holds no tests or golden vectors for them.  The functions below
restate the published algorithms (Lu et al., DPM-Solver++ multistep 2M; Karras et al. 2022 rho
schedule) and are anchored on the reference's own call sites:

  sigma schedule      /root/reference/source/modules/model_k_diffusion.py:848-859  (karras, sigma_min/max
                      = first/last of the model's 1000 training sigmas)
  latent init         /root/reference/source/modules/model_k_diffusion.py:1043     (randn * sqrt(sigma0^2+1))
  denoiser scalings   /root/reference/source/modules/external_k_diffusion.py:95-98, :109-114
  sigma_to_t          /root/reference/source/modules/external_k_diffusion.py:65-77
  CFG on denoised     /root/reference/source/modules/model_k_diffusion.py:1162-1166
  sampler selection   /root/reference/source/app.py:198   ('DPM++ 2M Karras' -> sample_dpmpp_2m, karras)
  diffusers twin      /root/reference/source/app.py:250, source/modules/model_diffusers.py:346-347,376-386

What IS checked (tests/test_oracle_sampler.py): the VE (k-diffusion) and VP (diffusers) formulations
below agree to fp64 round-off over 25 steps on a toy denoiser, and the schedule reproduces the
known-answer values of SURVEY.md Appendix B.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import torch


def sd15_train_sigmas(dtype=torch.float32) -> torch.Tensor:
    """SD-1.5 scaled-linear betas (0.00085..0.012, 1000 steps) -> sigma_t = sqrt((1-abar)/abar).

    This is what ``DiscreteEpsDDPMDenoiser.__init__`` (external_k_diffusion.py:90-91) stores."""
    betas = torch.linspace(0.00085**0.5, 0.012**0.5, 1000, dtype=torch.float32) ** 2
    alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
    return (((1 - alphas_cumprod) / alphas_cumprod) ** 0.5).to(dtype)


def get_sigmas_karras(n: int, sigma_min: float, sigma_max: float, rho: float = 7.0) -> torch.Tensor:
    """k_diffusion.sampling.get_sigmas_karras + append_zero (fp32, like the library)."""
    ramp = torch.linspace(0, 1, n)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return torch.cat([sigmas, sigmas.new_zeros([1])])


def sigma_to_t(sigma: torch.Tensor, log_sigmas: torch.Tensor) -> torch.Tensor:
    """external_k_diffusion.py:65-77 with quantize=False (fractional timestep)."""
    log_sigma = sigma.log()
    dists = log_sigma - log_sigmas[:, None]
    low_idx = dists.ge(0).cumsum(dim=0).argmax(dim=0).clamp(max=log_sigmas.shape[0] - 2)
    high_idx = low_idx + 1
    low, high = log_sigmas[low_idx], log_sigmas[high_idx]
    w = ((low - log_sigma) / (low - high)).clamp(0, 1)
    t = (1 - w) * low_idx + w * high_idx
    return t.view(sigma.shape)


def cfg_denoise(eps_fn: Callable, x: torch.Tensor, sigma: torch.Tensor, guidance: float) -> torch.Tensor:
    """model_fn of model_k_diffusion.py:1091-1171 around DiscreteEpsDDPMDenoiser.forward.

    ``eps_fn(x_in[2n], sigma) -> eps[2n]`` with rows (uncond..., cond...)."""
    x2 = torch.cat([x] * 2)
    c_in = 1 / (sigma**2 + 1) ** 0.5
    eps = eps_fn(x2 * c_in, sigma)
    den = x2 + eps * (-sigma)
    den_u, den_c = den.chunk(2)
    return den_u + guidance * (den_c - den_u)


def sample_dpmpp_2m(model: Callable, x: torch.Tensor, sigmas: torch.Tensor) -> torch.Tensor:
    """k-diffusion's DPM-Solver++(2M) in VE space.  ``model(x, sigma) -> denoised``."""
    t_fn = lambda s: s.log().neg()
    old_denoised: Optional[torch.Tensor] = None
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i])
        t, t_next = t_fn(sigmas[i]), t_fn(sigmas[i + 1])
        h = t_next - t
        if old_denoised is None or sigmas[i + 1] == 0:
            x = (sigmas[i + 1] / sigmas[i]) * x - (-h).expm1() * denoised
        else:
            h_last = t - t_fn(sigmas[i - 1])
            r = h_last / h
            denoised_d = (1 + 1 / (2 * r)) * denoised - (1 / (2 * r)) * old_denoised
            x = (sigmas[i + 1] / sigmas[i]) * x - (-h).expm1() * denoised_d
        old_denoised = denoised
    return x


def sample_dpmpp_2m_vp(model: Callable, x_ve0: torch.Tensor, sigmas: torch.Tensor) -> torch.Tensor:
    """diffusers' DPMSolverMultistepScheduler (dpmsolver++, order 2, midpoint, karras sigmas,
    final_sigmas_type='zero', lower_order_final) written in its own VP variables.

    Takes/returns VE-space latents so it can be compared 1:1 with ``sample_dpmpp_2m``; the same
    ``model(x_ve, sigma) -> denoised`` callable is used (x0-prediction is space-independent)."""
    alpha = lambda s: 1 / (s**2 + 1) ** 0.5
    sample = x_ve0 * alpha(sigmas[0])  # diffusers: latents * init_noise_sigma(=1) in VP space
    m_prev = None
    n = len(sigmas) - 1
    for i in range(n):
        s0, s1 = sigmas[i], sigmas[i + 1]
        a0, a1 = alpha(s0), alpha(s1)
        sig0, sig1 = s0 * a0, s1 * a1
        m0 = model(sample / a0, s0)  # x0 prediction
        lam0 = torch.log(a0) - torch.log(sig0)
        if s1 == 0:
            # final step: sigma_t = 0, alpha_t = 1, exp(-h) -> 0
            sample = m0.clone()
            m_prev = m0
            continue
        lam1 = torch.log(a1) - torch.log(sig1)
        h = lam1 - lam0
        first_order = m_prev is None or i == n - 1
        if first_order:
            sample = (sig1 / sig0) * sample - a1 * (torch.exp(-h) - 1.0) * m0
        else:
            sp, ap = sigmas[i - 1], alpha(sigmas[i - 1])
            lamp = torch.log(ap) - torch.log(sp * ap)
            h0 = lam0 - lamp
            r0 = h0 / h
            D0, D1 = m0, (1.0 / r0) * (m0 - m_prev)
            sample = (sig1 / sig0) * sample - a1 * (torch.exp(-h) - 1.0) * D0 - 0.5 * a1 * (torch.exp(-h) - 1.0) * D1
        m_prev = m0
    return sample  # sigma_last = 0 -> alpha = 1 -> VP == VE


def txt2img_latents(
    eps_fn: Callable,
    noise: torch.Tensor,  # [n, 4, h, w] unit normal
    steps: int = 25,
    guidance: float = 7.5,
    sigma_cb: Optional[Callable] = None,
) -> torch.Tensor:
    """The 25-step loop of model_k_diffusion.py:1027-1175 (latents only; no VAE)."""
    train = sd15_train_sigmas()
    sigmas = get_sigmas_karras(steps, train[0].item(), train[-1].item()).to(noise)
    x = noise * (sigmas[0] ** 2 + 1) ** 0.5

    def model(x_, sigma):
        if sigma_cb is not None:
            sigma_cb(sigma)
        return cfg_denoise(eps_fn, x_, sigma, guidance)

    return sample_dpmpp_2m(model, x, sigmas)
