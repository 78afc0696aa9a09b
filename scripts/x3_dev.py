"""Dev check + timing of the 3-warpgroup tcgen05 kernels (prepared K/V image) against the x4 kernels and an fp32
evaluation of the same formula.  Usage: python scripts/x3_dev.py [B L [iters]] ; DSC_LIB selects a debug build."""
from __future__ import annotations

import ctypes
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionspatialcontrol_b200 import _lib, attention as att  # noqa: E402


def fp32_formula(q, k, v, W, sigma):
    """softmax(QK^T/sqrt(D) + sigma*std*W) V in fp32 (dev-tool restatement; parity proper lives in tests/ vs oracle/)."""
    q, k, v = q.float(), k.float(), v.float()
    a = q @ k.transpose(-2, -1) / math.sqrt(q.size(-1))
    B, H, L, S = a.shape
    std = a.std()
    a = a + (W * sigma * std).repeat_interleave(B // W.shape[0], dim=0)[:, None]
    return torch.softmax(a, dim=-1) @ v, float(std)


def watchdog():
    fn = getattr(_lib.lib, "dsc_debug_x3_watchdog", None) if hasattr(_lib.lib, "_name") else None
    try:
        fn = _lib.lib.dsc_debug_x3_watchdog
    except AttributeError:
        return None
    out = (ctypes.c_uint32 * 5)()
    fn(out)
    return list(out)


def run(B, L, iters, dtype=torch.float16, H=8, D=40, S=77):
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(1234 + L)
    q = torch.randn(B, L, H * D, device=dev, dtype=dtype, generator=g)
    k = torch.randn(B, S, H * D, device=dev, dtype=dtype, generator=g)
    v = torch.randn(B, S, H * D, device=dev, dtype=dtype, generator=g)
    W = torch.zeros(B, L, S, device=dev)
    W[:, : L // 2, 1:3] = 0.5
    W[:, L // 3:, 6] = 0.7
    W[:, L // 4: L // 2, 6] = -0.2
    W = att.padded_region_map(W)
    compact = att.compact_region_map(W)
    view = lambda t: t.view(B, -1, H, D).transpose(1, 2)
    q4, k4, v4 = view(q), view(k), view(v)
    sigma = 7.0
    kv = att.prepare_kv(k4, v4, compact[1])
    torch.cuda.synchronize()
    out3 = att.region_attention_prepared(q4, kv, compact, sigma)
    torch.cuda.synchronize()
    wd = watchdog()
    st3 = att.read_stats(att.get_workspace(dev))
    out4 = att.region_attention(q4, k4, v4, W, sigma, compact=compact)
    torch.cuda.synchronize()
    st4 = att.read_stats(att.get_workspace(dev))
    ref, std_ref = fp32_formula(q4, k4, v4, W, sigma)
    rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())
    rec = {"B": B, "L": L, "dtype": str(dtype).split(".")[-1], "watchdog": wd, "std_x3": st3["std"], "std_x4": st4["std"], "std_ref": std_ref,
           "rel_x3_vs_fp32": rel(out3, ref), "rel_x4_vs_fp32": rel(out4, ref), "rel_x3_vs_x4": rel(out3, out4),
           "nan_x3": bool(torch.isnan(out3).any())}
    if iters > 0:
        flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")

        def timeit(fn):
            ts = []
            for _ in range(iters):
                flush.zero_()
                flush[: flush.numel() // 2].view(torch.int64).sum()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            ts.sort()
            mid = ts[len(ts) // 5: len(ts) - len(ts) // 5]  # trimmed mean: the event clock ticks in ~1 us steps
            return round(sum(mid) / len(mid), 2)

        o = torch.empty(B, L, H * D, device=dev, dtype=dtype)
        for name, ps in (("x3_stats", 1), ("x3_forward", 2), ("x3_both", 3)):
            rec["us_" + name] = timeit(lambda: att.region_attention_prepared(q4, kv, compact, sigma, passes=ps, out=o))
        rec["us_x4_stats"] = timeit(lambda: att.score_stats(q4, k4))
        rec["us_x4_call"] = timeit(lambda: att.region_attention(q4, k4, v4, W, sigma, compact=compact))
        rec["us_prepare_kv"] = timeit(lambda: att.prepare_kv(k4, v4, compact[1], out=kv.image))
        nbytes = 2 * B * H * L * D * 3 + 2 * B * H * S * D * 3 + 4 * B * L * S
        rec["frac_x3"] = nbytes / (rec["us_x3_both"] * 1e-6) / 1e9 / 6547.8
        rec["frac_x4"] = nbytes / (rec["us_x4_call"] * 1e-6) / 1e9 / 6547.8
    print(json.dumps(rec), flush=True)
    return rec


if __name__ == "__main__":
    if len(sys.argv) >= 3:
        run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    else:
        for B, L in ((1, 128), (2, 256), (1, 4096), (2, 1000), (16, 4096)):
            run(B, L, 0)
        run(16, 4096, 20)
        run(16, 4096, 20, dtype=torch.bfloat16)
        run(32, 4096, 20)
        run(8, 9216, 20)
