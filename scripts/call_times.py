"""Distribution of the flushed call times the way bench.py measures them (two alternating input sets, device sigma, mean):
python scripts/call_times.py B L D [n]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

B, L, D = (int(x) for x in sys.argv[1:4])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 50
dev = torch.device("cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
from diffusionspatialcontrol_b200 import attention as att  # noqa: E402

vw = lambda t: t.view(B, -1, 8, D).transpose(1, 2)
sets = []
for i in range(2):
    q = torch.randn(B, L, 8 * D, device=dev, dtype=torch.float16)
    k = torch.randn(B, 77, 8 * D, device=dev, dtype=torch.float16)
    v = torch.randn(B, 77, 8 * D, device=dev, dtype=torch.float16)
    W = torch.zeros(B, L, 77, device=dev)
    W[:, : L // 2, 1:3] = 0.5
    W = att.padded_region_map(W)
    compact = att.compact_region_map(W)
    kv = att.prepare_kv(vw(k), vw(v), compact[1])
    sets.append((q, compact, kv, torch.empty_like(q)))
sigma = torch.tensor(7.0, device=dev)
ws = att.get_workspace(dev, att.workspace_bytes(B, 8, L, D, 77))
for i in range(6):
    q, compact, kv, out = sets[i % 2]
    att.region_attention_prepared(vw(q), kv, compact, sigma, workspace=ws, out=out)
ts = []
for it in range(n):
    flush.zero_()
    flush[: flush.numel() // 2].view(torch.int64).sum()
    q, compact, kv, out = sets[it % 2]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    att.region_attention_prepared(vw(q), kv, compact, sigma, workspace=ws, out=out)
    b.record()
    b.synchronize()
    ts.append(round(a.elapsed_time(b) * 1e3, 1))
print(B, L, D, "mean", round(sum(ts) / len(ts), 2), "median", sorted(ts)[len(ts) // 2], "in order:", ts)
