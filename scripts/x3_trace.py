"""In-kernel timeline of the 3-warpgroup tcgen05 kernels (-DDSC_TRACE build selected with DSC_LIB): block 0's consumer
warpgroups, TMA producer and tensor-core issuers, plus the start / end of every CTA (globaltimer).
Usage: DSC_LIB=.../libdsc_trace.so python scripts/x3_trace.py [B L [D]] > profiles/...txt"""
import ctypes
import math
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
for _ in range(3):
    flush.zero_()
    flush[: flush.numel() // 2].view(torch.int64).sum()
    att.region_attention_prepared(view(q), kv, compact, 7.0)
torch.cuda.synchronize()
out = np.zeros((2, 8, 1024, 2), dtype=np.int64)
cnt = np.zeros((2, 8), dtype=np.int32)
cta = np.zeros((2, 160, 2), dtype=np.uint64)
_lib.lib.dsc_debug_x3_trace(out.ctypes.data_as(ctypes.c_void_p), cnt.ctypes.data_as(ctypes.c_void_p), cta.ctypes.data_as(ctypes.c_void_p))
names = ["consumer wg0", "consumer wg1", "consumer wg2", "producer", "issuer 0", "issuer 1", "issuer 2", "drain warp 0"]
for ps, pname in ((0, "pass 1 (stats)"), (1, "pass 2 (forward)")):
    c = cta[ps, :148].astype(np.int64)
    t0 = c[:, 0].min()
    print(f"===== {pname}: CTA starts {int((c[:,0]-t0).min())}..{int((c[:,0]-t0).max())} ns, ends {int((c[:,1]-t0).min())}..{int((c[:,1]-t0).max())} ns, "
          f"durations {int((c[:,1]-c[:,0]).min())}..{int((c[:,1]-c[:,0]).max())} ns")
    ends = np.sort(c[:, 1] - t0)
    print("      CTA end times (ns, sorted, every 10th): " + " ".join(str(int(x)) for x in ends[::10]) + f" ... {int(ends[-1])}")
    print("      CTA durations by block id (ns): " + " ".join(str(int(x)) for x in (c[:, 1] - c[:, 0])))
    if ps == 1:
        s1 = cta[0, :148].astype(np.int64)
        print(f"      pass 2 first CTA start - pass 1 first CTA start = {int(t0 - s1[:,0].min())} ns; pass 1 last end - pass 1 first start = {int(s1[:,1].max() - s1[:,0].min())} ns; "
              f"pass 2 last end - pass 1 first start = {int(c[:,1].max() - s1[:,0].min())} ns")
    base = min(out[ps, w, 0, 1] for w in range(8) if cnt[ps, w] > 0)
    for w in range(8):
        n = cnt[ps, w]
        print(f"--- {names[w]} ({n} events): tag@cycle(+delta)")
        prev = base
        line = []
        for e in range(n):
            tag, clk = out[ps, w, e]
            line.append(f"{tag}@{clk - base}(+{clk - prev})")
            prev = clk
        print(" ".join(line))
