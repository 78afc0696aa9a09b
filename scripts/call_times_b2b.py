"""Back-to-back attention calls over rotating input sets whose total footprint exceeds the L2 several times (no flush kernel
between the calls: by the time a set is used again its lines have been evicted), ONE CUDA-event pair around the whole run:
python scripts/call_times_b2b.py B L D [footprint_MB]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionspatialcontrol_b200 import attention as att  # noqa: E402

B, L, D = (int(x) for x in sys.argv[1:4])
foot = int(sys.argv[4]) if len(sys.argv) > 4 else 512
H, S = 8, 77
dev = torch.device("cuda")
per_set = 2 * B * L * H * D * 2 + B * L * 20 * 4
n_sets = max(4, min(48, -(-foot * (1 << 20) // per_set)))
vw = lambda t: t.view(B, -1, H, D).transpose(1, 2)
sets = []
for i in range(n_sets):
    q = torch.randn(B, L, H * D, device=dev, dtype=torch.float16)
    k = torch.randn(B, S, H * D, device=dev, dtype=torch.float16)
    v = torch.randn(B, S, H * D, device=dev, dtype=torch.float16)
    W = torch.zeros(B, L, S, device=dev)
    W[:, : L // 2, 1:3] = 0.5
    W = att.padded_region_map(W)
    compact = att.compact_region_map(W)
    kv = att.prepare_kv(vw(k), vw(v), compact[1])
    sets.append((q, compact, kv, torch.empty_like(q)))
sigma = torch.tensor(7.0, device=dev)
ws = att.get_workspace(dev, att.workspace_bytes(B, H, L, D, S))
call = lambda t: att.region_attention_prepared(vw(t[0]), t[2], t[1], sigma, workspace=ws, out=t[3])
for t in sets:
    call(t)
torch.cuda.synchronize()
res = []
for rep in range(5):
    n = 3 * n_sets
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        call(sets[i % n_sets])
    b.record()
    b.synchronize()
    res.append(a.elapsed_time(b) * 1e3 / n)
# the same run as ONE CUDA graph (how the pipeline issues the call: kernel nodes, no per-call host work)
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    ws_g = torch.zeros(att.workspace_bytes(B, H, L, D, S), dtype=torch.uint8, device=dev)
    call_g = lambda t: att.region_attention_prepared(vw(t[0]), t[2], t[1], sigma, workspace=ws_g, out=t[3])
    call_g(sets[0])
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
n = 3 * n_sets
with torch.cuda.graph(g):
    for i in range(n):
        call_g(sets[i % n_sets])
g.replay()
torch.cuda.synchronize()
res_g = []
for rep in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    b.synchronize()
    res_g.append(a.elapsed_time(b) * 1e3 / n)
bytes_alg = 2 * B * H * L * D * 3 + 2 * B * H * S * D * 3 + 4 * B * L * S
us = sorted(res)[len(res) // 2]
ug = sorted(res_g)[len(res_g) // 2]
print(json.dumps({"graph_us_per_call": [round(x, 2) for x in res_g], "graph_median_us": round(ug, 2),
                  "graph_frac_of_6547.8": round(bytes_alg / (ug * 1e-6) / 1e9 / 6547.8, 3)}))
print(json.dumps({"B": B, "L": L, "D": D, "n_sets": n_sets, "footprint_MB": round(n_sets * per_set / 2**20), "us_per_call": [round(x, 2) for x in res],
                  "median_us": round(us, 2), "frac_of_6547.8": round(bytes_alg / (us * 1e-6) / 1e9 / 6547.8, 3)}))
