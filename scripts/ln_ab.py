"""LayerNorm over the channel dim of the transformer blocks' [B, L, C] activations: is there a faster plain-PyTorch form than
F.layer_norm (48 launches, 50 us each, 1.7 TB/s in the UNet step)?  python scripts/ln_ab.py"""
import json
import torch
import torch.nn.functional as F

dev = torch.device("cuda")


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); b.synchronize()
    return a.elapsed_time(b) * 1e3 / n


for (B, L, C) in ((16, 4096, 320), (16, 1024, 640), (16, 256, 1280), (16, 64, 1280)):
    x = torch.randn(B, L, C, device=dev, dtype=torch.float16) * 1.5 + 0.3
    w = (torch.randn(C, device=dev) * 0.3 + 1).half()
    b = (torch.randn(C, device=dev) * 0.2).half()
    ref = F.layer_norm(x.float(), (C,), w.float(), b.float(), 1e-5)
    R = B * L

    def ln():
        return F.layer_norm(x, (C,), w, b, 1e-5)

    def gn1():
        return F.group_norm(x.reshape(R, C, 1), 1, w, b, 1e-5).reshape(B, L, C)

    def bn():
        y = F.batch_norm(x.reshape(1, R, C), None, None, None, None, True, 0.0, 1e-5)
        return torch.addcmul(b, y.reshape(B, L, C), w)

    def manual():
        var, mean = torch.var_mean(x, dim=-1, correction=0, keepdim=True)
        rstd = torch.rsqrt(var.float() + 1e-5).to(x.dtype)
        return torch.addcmul(b, (x - mean) * rstd, w)

    row = {"shape": [B, L, C], "MB": round(x.numel() * 2 / 2**20, 1)}
    for name, fn in (("F.layer_norm", ln), ("group_norm_1_group_per_row", gn1), ("batch_norm_rows_as_channels+affine", bn), ("var_mean+3_passes", manual)):
        try:
            y = fn()
            row[name] = {"us": round(timed(fn), 1), "max_err": round(float((y.float() - ref).abs().max()), 4)}
        except Exception as e:
            row[name] = {"error": repr(e)[:80]}
    print(json.dumps(row))
