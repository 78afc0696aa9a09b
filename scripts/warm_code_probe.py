"""How much of a flushed-L2 attention call is cold CODE / launch state rather than cold DATA?  After the L2 flush a tiny
call of the same kernels (other buffers, 128 query rows) runs before the timed call: instructions, tensor-map
descriptors' cache lines and the shared-memory carve-out are then warm, the timed call's own data is still cold.
Informational: the official figures in bench.py / microbench.py use the plain flush."""
import ctypes, math, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import diffusionspatialcontrol_b200 as dsc
from diffusionspatialcontrol_b200 import attention as att
B, H, L, D, S = 16, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 40, 77
mk = lambda b, l: (torch.randn(b, l, H * D, device="cuda", dtype=torch.float16), torch.randn(b, S, H * D, device="cuda", dtype=torch.float16), torch.randn(b, S, H * D, device="cuda", dtype=torch.float16))
view = lambda t: t.view(t.shape[0], -1, H, D).transpose(1, 2)
def wmap(b, l):
    W = torch.zeros(b, l, S, device="cuda"); W[:, : l // 2, 1:3] = 0.5
    W = att.padded_region_map(W); return W, att.compact_region_map(W)
sets = [(mk(B, L), wmap(B, L)) for _ in range(2)]
tiny = (mk(2 if D != 40 else 16, 1024), wmap(2 if D != 40 else 16, 1024))  # same kernel family as the timed call
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
def call(s):
    (q, k, v), (W, c) = s
    return dsc.region_attention(view(q), view(k), view(v), W, 7.0, compact=c)
for s in sets: call(s)
call(tiny)
res = {"plain": [], "warm_code": []}
for it in range(40):
    for mode in ("plain", "warm_code"):
        flush.zero_(); flush[: flush.numel() // 2].view(torch.int64).sum()
        if mode == "warm_code": call(tiny)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(sets[it % 2]); b.record(); b.synchronize()
        res[mode].append(a.elapsed_time(b) * 1e3)
for m, v in res.items():
    v = sorted(v); print(f"L={L} D={D} {m:10s} median {v[len(v)//2]:.1f} us  mean {sum(v)/len(v):.1f} us")
