import os, sys, torch
sys.path.insert(0, '/root/repo')
from diffusionspatialcontrol_b200.unet_sd15 import UNetSD15
from diffusionspatialcontrol_b200 import RegionAttnProcessor
from diffusionspatialcontrol_b200.pipeline import reference_weight_func
from torch.profiler import ProfilerActivity, profile
dev = torch.device("cuda")
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
net = UNetSD15().to(dev, torch.float16).eval().to(memory_format=torch.channels_last)
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
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        net(x, t, ctx, cross_attention_kwargs=kw)
        torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=40, max_shapes_column_width=90))
