"""Run the two attention passes a few times on one shape (for ncu). Usage: prof_one.py B L D [iters]"""
import sys, os, math, ctypes, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffusionspatialcontrol_b200 as dsc
B, L, D = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
H, S = 8, 77
dtype = torch.float16
q = torch.randn(B, L, H * D, device="cuda", dtype=dtype)
k = torch.randn(B, S, H * D, device="cuda", dtype=dtype)
v = torch.randn(B, S, H * D, device="cuda", dtype=dtype)
W = torch.zeros(B, L, S, device="cuda"); W[:, : L // 2, 1:3] = 0.5
if os.environ.get("DSC_W_LAYOUT", "compact") in ("padded", "compact"):
    from diffusionspatialcontrol_b200.attention import padded_region_map
    W = padded_region_map(W)
from diffusionspatialcontrol_b200.attention import compact_region_map
COMPACT = compact_region_map(W) if os.environ.get("DSC_W_LAYOUT", "compact") == "compact" else None
view = lambda t: t.view(B, -1, H, D).transpose(1, 2)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for _ in range(iters):
    flush.zero_()
    o = dsc.region_attention(view(q), view(k), view(v), W, 7.0, compact=COMPACT)
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
