"""A/B timing of the prepared-K/V call for library variants: python scripts/x3_ab.py [iters] [shapes "BxLxD,..."] ; DSC_LIB selects
the build.  Prints one JSON line: per shape the trimmed-mean us of pass 1 / pass 2 / the call (L2 flushed) + rel-L2 vs fp32."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from diffusionspatialcontrol_b200 import attention as att  # noqa: E402
from x3_dev import fp32_formula  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
shapes = sys.argv[2] if len(sys.argv) > 2 else "16x4096x40,16x1024x80,16x256x160,16x64x160"
S = 77
dev = torch.device("cuda")
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
res = {"lib": os.path.basename(os.environ.get("DSC_LIB", "libdsc_b200.so"))}
for sh in shapes.split(","):
    B, L, D = (int(x) for x in sh.split("x"))
    H = 8
    g = torch.Generator(device="cuda").manual_seed(1234 + L)
    q = torch.randn(B, L, H * D, device=dev, dtype=torch.float16, generator=g)
    k = torch.randn(B, S, H * D, device=dev, dtype=torch.float16, generator=g)
    v = torch.randn(B, S, H * D, device=dev, dtype=torch.float16, generator=g)
    W = torch.zeros(B, L, S, device=dev)
    W[:, : L // 2, 1:3] = 0.5
    W[:, L // 3:, 6] = 0.7
    W = att.padded_region_map(W)
    compact = att.compact_region_map(W)
    view = lambda t: t.view(B, -1, H, D).transpose(1, 2)
    q4, k4, v4 = view(q), view(k), view(v)
    kv = att.prepare_kv(k4, v4, compact[1])
    o = torch.empty(B, L, H * D, device=dev, dtype=torch.float16)
    out = att.region_attention_prepared(q4, kv, compact, 7.0)
    ref, _ = fp32_formula(q4, k4, v4, W, 7.0)
    rel = float((out.float() - ref).norm() / ref.norm())

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
        mid = ts[len(ts) // 5: len(ts) - len(ts) // 5]
        return round(sum(mid) / len(mid), 2)

    r = [timeit(lambda: att.region_attention_prepared(q4, kv, compact, 7.0, passes=ps, out=o)) for ps in (1, 2, 3)]
    res[sh] = {"p1": r[0], "p2": r[1], "call": r[2], "rel": round(rel, 6)}
print(json.dumps(res), flush=True)
