"""Kernel launches of ONE eager UNet step (batch 16) by name: count, total and mean device time -- where the launch-bound
small kernels of the host come from.  python scripts/unet_launch_count.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionspatialcontrol_b200 import unet_sd15, RegionAttnProcessor
from diffusionspatialcontrol_b200.pipeline import reference_weight_func
from torch.profiler import ProfilerActivity, profile
dev = torch.device("cuda")
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
net = unet_sd15.UNetSD15().to(dev, torch.float16).eval().to(memory_format=torch.channels_last)
net.set_attn_processor(RegionAttnProcessor(cache_kv=True))
B = 16
x = torch.randn(B, 4, 64, 64, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
ctx = torch.randn(B, 77, 768, device=dev, dtype=torch.float16)
t = torch.tensor(500.0, device=dev)
rs = {L: torch.zeros(B, L, 77, device=dev) for L in (4096, 1024, 256, 64)}
for L in rs: rs[L][:, : L // 2, 1:3] = 0.5
kw = {"region_prompt": {"region_state": rs, "sigma": torch.tensor(7.0, device=dev), "weight_func": reference_weight_func}}
with torch.no_grad():
    for _ in range(3): net(x, t, ctx, cross_attention_kwargs=kw)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        net(x, t, ctx, cross_attention_kwargs=kw)
        torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0 or e.count > 0]
tot = sum(e.device_time_total for e in ev); n = sum(e.count for e in ev)
print(f"{n} launches, {tot/1e3:.2f} ms device time")
small = [e for e in ev if e.device_time_total / max(e.count, 1) < 6.0]
print(f"{sum(e.count for e in small)} launches under 6 us mean, {sum(e.device_time_total for e in small)/1e3:.2f} ms")
for e in sorted(ev, key=lambda e: -e.count)[:40]:
    print(f"{e.count:5d} {e.device_time_total/1e3:8.3f} ms {e.device_time_total/max(e.count,1):8.1f} us  {e.key[:120]}")
