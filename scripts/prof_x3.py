"""One flushed call of the prepared-K/V path (the single cooperative launch) for ncu.  Usage: python scripts/prof_x3.py [B L reps [D]]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionspatialcontrol_b200 import attention as att  # noqa: E402

B, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 4096)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
H, S = 8, 77
D = int(sys.argv[4]) if len(sys.argv) > 4 else 40
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
out = torch.empty(B, L, H * D, device=dev, dtype=torch.float16)
for _ in range(reps):
    flush.zero_()
    flush[: flush.numel() // 2].view(torch.int64).sum()
    att.region_attention_prepared(view(q), kv, compact, 7.0, out=out)
torch.cuda.synchronize()
print("ok")
