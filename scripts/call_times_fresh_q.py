"""The call as it sits in the UNet: Q has just been written by the to_q projection (L2-resident, dirty), everything else is
cold.  One CUDA graph of N x [write Q (a copy kernel standing in for the projection), attention call] over rotating input sets
(footprint >= 4 x L2), minus the same graph with the copies only:  python scripts/call_times_fresh_q.py B L D"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionspatialcontrol_b200 import attention as att  # noqa: E402

B, L, D = (int(x) for x in sys.argv[1:4])
H, S = 8, 77
dev = torch.device("cuda")
per_set = 3 * B * L * H * D * 2 + B * L * 20 * 4
n_sets = max(4, min(32, -(-(640 << 20) // per_set)))
vw = lambda t: t.view(B, -1, H, D).transpose(1, 2)
sets = []
for i in range(n_sets):
    src = torch.randn(B, L, H * D, device=dev, dtype=torch.float16)
    q = torch.empty_like(src)
    k = torch.randn(B, S, H * D, device=dev, dtype=torch.float16)
    v = torch.randn(B, S, H * D, device=dev, dtype=torch.float16)
    W = torch.zeros(B, L, S, device=dev)
    W[:, : L // 2, 1:3] = 0.5
    compact = att.compact_region_map(att.padded_region_map(W))
    sets.append((src, q, compact, att.prepare_kv(vw(k), vw(v), compact[1]), torch.empty_like(q)))
sigma = torch.tensor(7.0, device=dev)
ws = torch.zeros(att.workspace_bytes(B, H, L, D, S), dtype=torch.uint8, device=dev)


def graph_of(with_call):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for t in sets:
            t[1].copy_(t[0])
            att.region_attention_prepared(vw(t[1]), t[3], t[2], sigma, workspace=ws, out=t[4])
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    n = 3 * n_sets
    with torch.cuda.graph(g):
        for i in range(n):
            t = sets[i % n_sets]
            t[1].copy_(t[0])
            if with_call:
                att.region_attention_prepared(vw(t[1]), t[3], t[2], sigma, workspace=ws, out=t[4])
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / n)
    return sorted(ts)[2]


both, copy_only = graph_of(True), graph_of(False)
print(json.dumps({"B": B, "L": L, "D": D, "input_sets": n_sets, "us_copy_plus_call": round(both, 2), "us_copy_only": round(copy_only, 2),
                  "us_call_with_L2_resident_Q": round(both - copy_only, 2)}))
