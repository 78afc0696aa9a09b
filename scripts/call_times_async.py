"""Flushed call time with and without a host synchronize between the timed iterations (same flush + call sequence and
event placement as bench.py; without the per-iteration synchronize the GPU never idles between iterations, so the SM
clock stays where a long step holds it):  python scripts/call_times_async.py B L D [n]"""
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

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


def sm_clock():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        return pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    except Exception:
        return None


def run(sync_each, n, warm=3):
    evs, clocks, stop = [], [], False

    def sampler():
        while not stop:
            clocks.append(sm_clock())
            time.sleep(0.002)

    th = threading.Thread(target=sampler)
    th.start()
    for it in range(-warm, n):
        flush.zero_()
        flush[: flush.numel() // 2].view(torch.int64).sum()
        q, compact, kv, out = sets[it % 2]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        att.region_attention_prepared(vw(q), kv, compact, sigma, workspace=ws, out=out)
        b.record()
        if sync_each:
            b.synchronize()
        if it >= 0:
            evs.append((a, b))
    torch.cuda.synchronize()
    stop = True
    th.join()
    ts = [round(a.elapsed_time(b) * 1e3, 1) for a, b in evs]
    ck = sorted(c for c in clocks if c)
    return ts, (ck[len(ck) // 2] if ck else None)


for i in range(6):
    q, compact, kv, out = sets[i % 2]
    att.region_attention_prepared(vw(q), kv, compact, sigma, workspace=ws, out=out)
import json
for rep in range(2):
    for sync_each in (True, False):
        ts, ck = run(sync_each, n)
        print(json.dumps({"B": B, "L": L, "D": D, "sync_each_iteration": sync_each, "mean_us": round(sum(ts) / len(ts), 2),
                          "median_us": sorted(ts)[len(ts) // 2], "min_us": min(ts), "sm_mhz_median": ck}))
