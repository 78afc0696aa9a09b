"""Does Q stay in L2 between pass 1 and pass 2?  Times (CUDA events, L2 flushed first): stats | stats,stats | forward |
stats,forward for several batch sizes (Q = 2.6 MB per batch row at L=4096)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionspatialcontrol_b200 import attention as att  # noqa: E402

H, D, S, L = 8, 40, 77, 4096
dev = torch.device("cuda")
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def med(fn, iters=15):
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
    return round(ts[len(ts) // 2], 2)


for B in (2, 4, 8, 16, 32):
    q = torch.randn(B, L, H * D, device=dev).half()
    k = torch.randn(B, S, H * D, device=dev).half()
    v = torch.randn(B, S, H * D, device=dev).half()
    W = torch.zeros(B, L, S, device=dev)
    W[:, : L // 2, 1:3] = 0.5
    W = att.padded_region_map(W)
    compact = att.compact_region_map(W)
    view = lambda t: t.view(B, -1, H, D).transpose(1, 2)
    q4 = view(q)
    kv = att.prepare_kv(view(k), view(v), compact[1])
    out = torch.empty(B, L, H * D, device=dev, dtype=torch.float16)
    st = lambda: att.region_attention_prepared(q4, kv, compact, 7.0, passes=1, out=out)
    fw = lambda: att.region_attention_prepared(q4, kv, compact, 7.0, passes=2, out=out)
    st(); fw(); torch.cuda.synchronize()
    rec = {"B": B, "q_mb": round(q.numel() * 2 / 1e6, 1), "stats": med(st), "stats_stats": med(lambda: (st(), st())),
           "fwd": med(fw), "stats_fwd": med(lambda: (st(), fw())), "fwd_fwd": med(lambda: (fw(), fw())),
           "stats_x3": med(lambda: (st(), st(), st()))}
    print(json.dumps(rec), flush=True)
