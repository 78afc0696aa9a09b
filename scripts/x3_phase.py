"""Cycles per phase of the pass-2 consumer loop (-DDSC_PHASE build selected with DSC_LIB): register accumulators, no stores in
the loop.  Usage: DSC_LIB=.../libdsc_phase.so python scripts/x3_phase.py [B L [D]]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionspatialcontrol_b200 import _lib, attention as att  # noqa: E402

B, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 4096)
D = int(sys.argv[3]) if len(sys.argv) > 3 else 40
H, S = 8, 77
dev = torch.device("cuda")
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
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ts = []
for _ in range(6):
    flush.zero_()
    flush[: flush.numel() // 2].view(torch.int64).sum()
    ev[0].record()
    att.region_attention_prepared(view(q), kv, compact, 7.0)
    ev[1].record()
    torch.cuda.synchronize()
    ts.append(ev[0].elapsed_time(ev[1]) * 1e3)
print("call us:", " ".join(f"{t:.1f}" for t in ts))
out = np.zeros((4, 12, 8), dtype=np.uint32)
_lib.lib.dsc_debug_x3_phase(out.ctypes.data_as(ctypes.c_void_p))
names = ["top/kvfree", "wait S", "S ld+sfree", "beta/W/max", "wait PV", "turn", "M+STTM", "st wait+prdy"]
print("phase            " + " ".join(f"{n:>12s}" for n in names) + "        total")
for b in range(4):
    for w in range(0, 12):
        r = out[b, w]
        if r.sum() == 0:
            continue
        print(f"block {b} warp {w:2d}: " + " ".join(f"{int(x):12d}" for x in r) + f" {int(r.sum()):12d}")
items = None
