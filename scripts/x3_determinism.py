"""Bit-reproducibility probe of the prepared-K/V kernels: pass 1 sums and pass 2 outputs over repeated calls.
Usage: python scripts/x3_determinism.py [B L D reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionspatialcontrol_b200 import attention as att  # noqa: E402

B, L, D = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (16, 4096, 40)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
H, S = 8, 77
dev = torch.device("cuda")
torch.manual_seed(0)
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ws = att.get_workspace(dev)
sums = set()
for r in range(reps):
    if r % 2:
        flush.zero_()
    att.region_attention_prepared(view(q), kv, compact, 7.0, passes=1)
    st = att.read_stats(ws)
    sums.add((st["sum"], st["sumsq"], st["std"]))
print("pass 1 distinct (sum, sumsq, std):", len(sums), sorted(sums)[:3])
base = None
for r in range(reps):
    if r % 2:
        flush.zero_()
    o = att.region_attention_prepared(view(q), kv, compact, 7.0, passes=2).clone()
    if base is None:
        base = o
    else:
        d = (o != base)
        if d.any():
            idx = d.nonzero()
            print(f"pass 2 rep {r}: {int(d.sum())} differing elements; first: {idx[:5].tolist()}; "
                  f"rows {sorted(set(idx[:, 2].tolist()))[:10]} heads {sorted(set(idx[:, 1].tolist()))} batches {sorted(set(idx[:, 0].tolist()))[:10]} "
                  f"max abs diff {float((o.float() - base.float()).abs().max()):.3e}")
print("pass 2 done")
both = None
for r in range(reps):
    o = att.region_attention_prepared(view(q), kv, compact, 7.0).clone()
    if both is None:
        both = o
    elif not torch.equal(o, both):
        d = (o != both)
        print(f"both rep {r}: {int(d.sum())} differing elements, max abs diff {float((o.float() - both.float()).abs().max()):.3e}")
print("done")
