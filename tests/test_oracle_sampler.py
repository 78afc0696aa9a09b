"""Sampler oracle: PARITY UNPINNED against k_diffusion / diffusers (not available, see
oracle/sampler.py).  Checked here: schedule known-answer values and agreement of
the VE (k-diffusion) and VP (diffusers) formulations of DPM-Solver++(2M)."""
import torch

from oracle import sampler as osm


def test_train_sigmas_and_karras_known_answers():
    train = osm.sd15_train_sigmas()
    assert abs(train[0].item() - 0.0291675) < 1e-6 and abs(train[-1].item() - 14.6146469) < 1e-4
    s = osm.get_sigmas_karras(25, train[0].item(), train[-1].item())
    want = [14.6146, 12.2830, 10.2778, 8.5600, 7.0944, 5.8494, 4.7965, 3.9105, 3.1686, 2.5508, 2.0392, 1.6183, 1.2741,
            0.9947, 0.7695, 0.5895, 0.4469, 0.3350, 0.2480, 0.1811, 0.1303, 0.0923, 0.0642, 0.0437, 0.0292, 0.0]
    assert torch.allclose(s, torch.tensor(want), atol=6e-5)
    t = osm.sigma_to_t(s[:-1], train.log())
    want_t = [999.00, 969.76, 938.45, 904.78, 868.42, 828.96, 785.95, 738.87, 687.15, 630.27, 567.87, 500.00, 427.49,
              352.29, 277.60, 207.46, 145.94, 96.03, 58.89, 33.63, 17.86, 8.74, 3.80, 1.28, 0.00]
    assert torch.allclose(t, torch.tensor(want_t), atol=6e-3)


def test_ve_and_vp_formulations_agree():
    torch.manual_seed(0)
    train = osm.sd15_train_sigmas(torch.float64)
    sig = osm.get_sigmas_karras(25, train[0].item(), train[-1].item()).double()
    A = torch.randn(16, 16, dtype=torch.float64) * 0.1

    def model(x, sigma):  # smooth toy denoiser
        return torch.tanh(x @ A) / (1 + sigma) + 0.3 * x / (1 + sigma**2)

    x0 = torch.randn(4, 16, dtype=torch.float64) * (sig[0] ** 2 + 1) ** 0.5
    a = osm.sample_dpmpp_2m(model, x0.clone(), sig)
    b = osm.sample_dpmpp_2m_vp(model, x0.clone(), sig)
    assert torch.allclose(a, b, rtol=1e-9, atol=1e-9)


def test_first_and_last_steps_are_first_order():
    sig = torch.tensor([2.0, 1.0, 0.0], dtype=torch.float64)
    calls = []

    def model(x, sigma):
        calls.append(float(sigma))
        return 0.5 * x

    x = osm.sample_dpmpp_2m(model, torch.ones(3, dtype=torch.float64), sig)
    # step 1: x = (1/2) x + (1 - 1/2) * 0.5 x = 0.75 ; step 2 (sigma_next = 0): x = denoised = 0.375
    assert torch.allclose(x, torch.full((3,), 0.375, dtype=torch.float64)) and calls == [2.0, 1.0]
