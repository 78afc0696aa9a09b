"""In-kernel span of the prepared-K/V call (-DDSC_CTATIME build selected with DSC_LIB): globaltimer at the start / end of every
CTA of pass 1 and pass 2 (or of the two phases of the single launch), L2 flushed before every call.
Usage: DSC_LIB=.../libdsc_ctatime.so python scripts/x3_span.py [n] [shapes BxLxD,...]"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionspatialcontrol_b200 import _lib, attention as att  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
shapes = sys.argv[2] if len(sys.argv) > 2 else "16x4096x40"
H, S = 8, 77
dev = torch.device("cuda")
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for sh in shapes.split(","):
    B, L, D = (int(x) for x in sh.split("x"))
    q = torch.randn(B, L, H * D, device=dev).half()
    k = torch.randn(B, S, H * D, device=dev).half()
    v = torch.randn(B, S, H * D, device=dev).half()
    W = torch.zeros(B, L, S, device=dev)
    W[:, : L // 2, 1:3] = 0.5
    W[:, L // 3:, 6] = 0.7
    W = att.padded_region_map(W)
    compact = att.compact_region_map(W)
    view = lambda t: t.view(B, -1, H, D).transpose(1, 2)
    kv = att.prepare_kv(view(k), view(v), compact[1])
    out = torch.empty(B, L, H * D, device=dev, dtype=torch.float16)
    grid = min(148, B * (H * D // 160) * ((L + 127) // 128))
    rows = []
    q_src = q.clone()
    for it in range(n + 3):
        flush.zero_()
        flush[: flush.numel() // 2].view(torch.int64).sum()
        if os.environ.get("X3_SPAN_FRESH_Q"):  # as in the UNet: Q has just been written by the projection (L2-resident)
            q.copy_(q_src)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        att.region_attention_prepared(view(q), kv, compact, 7.0, out=out)
        b.record()
        torch.cuda.synchronize()
        cta = np.zeros((2, 160, 2), dtype=np.uint64)
        _lib.lib.dsc_debug_x3_cta(cta.ctypes.data_as(ctypes.c_void_p))
        c = cta[:, :grid].astype(np.int64)
        t0 = c[0, :, 0].min()
        if it >= 3:
            rows.append((a.elapsed_time(b) * 1e3, (c[1, :, 1].max() - t0) / 1e3, (c[0, :, 1].max() - t0) / 1e3, (np.median(c[0, :, 1]) - t0) / 1e3,
                         (c[1, :, 0].min() - t0) / 1e3, (np.median(c[1, :, 1]) - t0) / 1e3, np.median(c[1, :, 1] - c[1, :, 0]) / 1e3))
    r = np.array(rows)
    m = r.mean(axis=0)
    print(json.dumps({"lib": os.path.basename(os.environ.get("DSC_LIB", "")), "shape": sh, "event_us": round(m[0], 2), "span_us": round(m[1], 2),
                      "span_min": round(r[:, 1].min(), 2), "span_max": round(r[:, 1].max(), 2), "p1_last_end": round(m[2], 2),
                      "p1_median_end": round(m[3], 2), "p2_first_start": round(m[4], 2), "p2_median_end": round(m[5], 2),
                      "p2_median_cta_us": round(m[6], 2)}), flush=True)
