"""UNet step (batch 16, one CUDA graph) with the GroupNorm statistics per channel first above a size threshold:
python scripts/gn_mode_ab.py  (thresholds in MiB; a huge one = the one-pass Welford reduction everywhere)"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionspatialcontrol_b200 import unet_sd15
from diffusionspatialcontrol_b200 import RegionAttnProcessor
from diffusionspatialcontrol_b200.pipeline import reference_weight_func
dev = torch.device("cuda")
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
net = unet_sd15.UNetSD15().to(dev, torch.float16).eval().to(memory_format=torch.channels_last)
net.set_attn_processor(RegionAttnProcessor(cache_kv=False))
B = 16
x = torch.randn(B, 4, 64, 64, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
ctx = torch.randn(B, 77, 768, device=dev, dtype=torch.float16)
t = torch.tensor(500.0, device=dev)
rs = {L: torch.zeros(B, L, 77, device=dev) for L in (4096, 1024, 256, 64)}
for L in rs: rs[L][:, : L // 2, 1:3] = 0.5
kw = {"region_prompt": {"region_state": rs, "sigma": torch.tensor(7.0, device=dev), "weight_func": reference_weight_func}}
outs = {}
for mib in [1 << 20] + [int(a) for a in (sys.argv[1:] or ["32", "16", "8", "0", "32"])]:
    unet_sd15.GroupNorm.PER_CHANNEL_BYTES = mib << 20
    with torch.no_grad():
        for _ in range(3): y = net(x, t, ctx, cross_attention_kwargs=kw)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2): y = net(x, t, ctx, cross_attention_kwargs=kw)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            y = net(x, t, ctx, cross_attention_kwargs=kw)
        for _ in range(3): g.replay()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): g.replay()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    outs[mib] = y.float().clone()
    ref = outs[1 << 20]
    cos = float(torch.nn.functional.cosine_similarity(ref.flatten(), outs[mib].flatten(), dim=0))
    print(f"per-channel statistics from {mib} MiB: {dt*1e3:.2f} ms per UNet step (batch 16, graph); cosine vs Welford everywhere {cos:.7f}", flush=True)
    del g
