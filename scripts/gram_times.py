"""Per-CTA globaltimer breakdown of the Gram stats kernel (-DDSC_TRACE build, DSC_LIB=...): gram_times.py B L [flush]"""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["DSC_XATTN_STATS_IMPL"] = "gram"
from diffusionspatialcontrol_b200 import _lib, attention as att
B, L = int(sys.argv[1]), int(sys.argv[2]); flush_on = len(sys.argv) > 3
H, D, S = 8, 40, 77
q = torch.randn(B, L, H * D, device="cuda", dtype=torch.float16); k = torch.randn(B, S, H * D, device="cuda", dtype=torch.float16)
view = lambda t: t.view(B, -1, H, D).transpose(1, 2)
raw = ctypes.CDLL(str(_lib.LIB_PATH)); buf = (ctypes.c_ulonglong * (160 * 8))()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for it in range(4):
    if flush_on:
        flush.zero_(); flush[: flush.numel() // 2].view(torch.int64).sum()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); att.score_stats(view(q), view(k)); b.record(); b.synchronize()
    raw.dsc_debug_gram_times(buf)
    n = min(148, B * ((L + 63) // 64))
    T = [[buf[c * 8 + j] for j in range(7)] for c in range(n)]
    t0 = min(r[0] for r in T); tend = max(r[6] for r in T)
    names = ["entry", "init_done", "first_tile", "tiles_done", "flushed", "cta_synced", "published"]
    print(f"iter {it}: event {a.elapsed_time(b)*1e3:.1f} us; kernel span (first entry -> last publish) {(tend - t0)/1e3:.1f} us; entry skew {(max(r[0] for r in T)-t0)/1e3:.1f} us")
    for j in range(1, 7):
        d = [(r[j] - r[j - 1]) / 1e3 for r in T]
        print(f"   {names[j-1]:>11s} -> {names[j]:<11s}: min {min(d):6.2f}  avg {sum(d)/len(d):6.2f}  max {max(d):6.2f} us")
