"""The smallest tail shapes through both entry points, for ONE compute-sanitizer tool per gpurun call:
prepared-K/V call (single launch and pass 1 + pass 2) and raw-K/V call, ragged L (37, 100, 144), D in {40, 80, 160}, batch 2.
Usage: compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionspatialcontrol_b200 import attention as att  # noqa: E402

dev = torch.device("cuda")
H, S = 8, 77
for (B, L, D) in ((2, 37, 80), (2, 144, 160), (2, 100, 40), (1, 64, 160)):
    g = torch.Generator(device="cuda").manual_seed(L)
    q = torch.randn(B, L, H * D, device=dev, dtype=torch.float16, generator=g)
    k = torch.randn(B, S, H * D, device=dev, dtype=torch.float16, generator=g)
    v = torch.randn(B, S, H * D, device=dev, dtype=torch.float16, generator=g)
    W = torch.zeros(B, L, S, device=dev)
    W[:, : L // 2, 1:3] = 0.5
    W = att.padded_region_map(W)
    compact = att.compact_region_map(W)
    view = lambda t: t.view(B, -1, H, D).transpose(1, 2)
    kv = att.prepare_kv(view(k), view(v), compact[1])
    a = att.region_attention_prepared(view(q), kv, compact, 7.0)
    att.region_attention_prepared(view(q), kv, compact, 7.0, passes=1)
    b = att.region_attention_prepared(view(q), kv, compact, 7.0, passes=2)
    c = att.region_attention(view(q), view(k), view(v), W, 7.0, compact=compact)
    torch.cuda.synchronize()
    print(B, L, D, "single == two launches:", bool(torch.equal(a, b)), "rel vs raw path:",
          float((a.float() - c.float()).norm() / c.float().norm()), flush=True)
print("ok")
