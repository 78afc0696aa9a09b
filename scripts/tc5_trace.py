"""Timeline of block 0 from a -DDSC_TRACE build: python scripts/tc5_trace.py B L D"""
import ctypes, os, sys, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffusionspatialcontrol_b200 as dsc
from diffusionspatialcontrol_b200 import _lib
B, L, D = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]); H, S = 8, 77
q = torch.randn(B, L, H * D, device="cuda", dtype=torch.float16); k = torch.randn(B, S, H * D, device="cuda", dtype=torch.float16); v = torch.randn_like(k)
W = torch.zeros(B, L, S, device="cuda"); W[:, : L // 2, 1:3] = 0.5
if os.environ.get("DSC_W_LAYOUT", "compact") in ("padded", "compact"):
    from diffusionspatialcontrol_b200.attention import padded_region_map
    W = padded_region_map(W)
from diffusionspatialcontrol_b200.attention import compact_region_map
COMPACT = compact_region_map(W) if os.environ.get("DSC_W_LAYOUT", "compact") == "compact" else None
view = lambda t: t.view(B, -1, H, D).transpose(1, 2)
raw = ctypes.CDLL(str(_lib.LIB_PATH))
out = (ctypes.c_longlong * (4 * 512 * 2))(); cnt = (ctypes.c_int * 4)()
from diffusionspatialcontrol_b200 import attention as att
only_stats = len(sys.argv) > 4 and sys.argv[4] == "stats"
for it in range(3):
    if only_stats:
        att.score_stats(view(q), view(k))
    else:
        dsc.region_attention(view(q), view(k), view(v), W, 7.0, compact=COMPACT)
    torch.cuda.synchronize()
    raw.dsc_debug_trace(out, cnt)   # also resets; stats+fwd traces are concatenated per call
names = {50: "k.prefetch_issued", 51: "k.smem_zeroed", 52: "k.bars_tmem_ready", 31: "kv.enter", 32: "kv.loads_issued", 33: "kv.k_stored", 34: "kv.v_stored", 35: "kv.fenced", 1: "k.start", 2: "k.init_done", 3: "c.run_begin", 4: "c.kv_staged", 5: "c.first_tile_landed", 6: "c.first_q_staged", 7: "c.run_loop_done", 8: "c.run_drained", 9: "k.pre_final_sync", 30: "k.final_sync_done", 17: "c.o_ready", 18: "c.o_stored", 40: "p.loads_issued", 41: "p.stage_free", 42: "p.stored", 10: "c.pre_s_wait", 11: "c.s_ready", 12: "c.S_loaded", 13: "c.lookahead_done", 14: "c.softmax_done", 15: "c.o_drained", 16: "c.p_arrived",
         20: "m.pre_q_wait", 21: "m.q_ready", 22: "m.qk_issued", 23: "m.pre_p_wait", 24: "m.p_ready", 25: "m.pv_issued"}
for wi, wn in enumerate(["consumer wg0 (warp0)", "consumer wg1 (warp4)", "x4: producer (warp16) / x2: mma wg0", "x4: consumer wg2 (warp8) / x2: mma wg1"]):
    n = cnt[wi]; ev = [(out[(wi * 512 + j) * 2], out[(wi * 512 + j) * 2 + 1]) for j in range(n)]
    print("==", wn, n, "events")
    if not ev: continue
    t0 = ev[0][1]; prev = t0
    for tag, t in ev[:400]:
        print(f"   {names.get(tag, tag):18s} t={t - t0:8d}  +{t - prev:6d}")
        prev = t
