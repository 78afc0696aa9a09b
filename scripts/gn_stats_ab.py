"""Which plain-PyTorch formulation of the channels_last GroupNorm statistics is fastest?  (UNet host, not the hot path.)
python scripts/gn_stats_ab.py"""
import json

import torch

dev = torch.device("cuda")
G = 32


def timed(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) * 1e3 / n


for (N, C, H, W) in ((16, 320, 64, 64), (16, 640, 32, 32), (16, 1280, 16, 16), (16, 1280, 8, 8), (16, 640, 64, 64), (16, 960, 64, 64),
                     (16, 1920, 32, 32), (16, 2560, 16, 16)):
    x = torch.randn(N, C, H, W, device=dev, dtype=torch.float16).add_(0.3).contiguous(memory_format=torch.channels_last)
    xl = x.permute(0, 2, 3, 1)
    Cg = C // G
    xv = xl.reshape(N, H * W, G, Cg)
    xc = xl.reshape(N, H * W, C)
    ones = torch.ones(N, 1, H * W, device=dev, dtype=torch.float16)
    ref_var, ref_mean = torch.var_mean(xv.double(), dim=(1, 3), correction=0)

    def a_():
        return torch.var_mean(xv, dim=(1, 3), correction=0)

    def b_():
        v, m = torch.var_mean(xc, dim=1, correction=0)
        m, v = m.float().view(N, G, Cg), v.float().view(N, G, Cg)
        mg = m.mean(-1)
        return (v + m * m).mean(-1) - mg * mg, mg

    def c_():
        s = xc.sum(dim=1, dtype=torch.float32).view(N, G, Cg).sum(-1) / (H * W * Cg)
        q = torch.linalg.vector_norm(xc, dim=1, dtype=torch.float32).square().view(N, G, Cg).sum(-1) / (H * W * Cg)
        return q - s * s, s

    def d_():
        s = torch.bmm(ones, xc).float().view(N, G, Cg).sum(-1) / (H * W * Cg)
        q = torch.bmm(ones, xc * xc).float().view(N, G, Cg).sum(-1) / (H * W * Cg)
        return q - s * s, s

    def e_():  # one pass over a [N, HW/r, r*C] view first? (longer contiguous rows for the column reduction)
        v, m = torch.var_mean(xc.float(), dim=1, correction=0)
        m, v = m.view(N, G, Cg), v.view(N, G, Cg)
        mg = m.mean(-1)
        return (v + m * m).mean(-1) - mg * mg, mg

    row = {"shape": [N, C, H, W], "MB": round(x.numel() * 2 / 2**20, 1)}
    for name, fn in (("var_mean_dims13", a_), ("var_mean_dim1_then_groups", b_), ("sum+norm_dim1_fp32", c_), ("bmm_ones", d_),
                     ("float_var_mean_dim1", e_)):
        v, m = fn()
        row[name] = {"us": round(timed(fn), 1), "rel_err_var": float(((v.double() - ref_var).abs() / ref_var).max()),
                     "abs_err_mean": float((m.double() - ref_mean).abs().max())}
    print(json.dumps(row))
