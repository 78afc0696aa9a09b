"""Quick A/B of PyTorch-side settings for the UNet host (not our kernels): channels_last, cudnn.benchmark, CUDA graph."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionspatialcontrol_b200.unet_sd15 import UNetSD15
from diffusionspatialcontrol_b200 import RegionAttnProcessor
from diffusionspatialcontrol_b200.pipeline import reference_weight_func
dev = torch.device("cuda")
def run(channels_last, bench, graph):
    torch.backends.cudnn.benchmark = bench
    torch.manual_seed(0)
    net = UNetSD15().to(dev, torch.float16).eval()
    if channels_last: net = net.to(memory_format=torch.channels_last)
    net.set_attn_processor(RegionAttnProcessor(cache_kv=False))
    B = 16
    x = torch.randn(B, 4, 64, 64, device=dev, dtype=torch.float16)
    if channels_last: x = x.contiguous(memory_format=torch.channels_last)
    ctx = torch.randn(B, 77, 768, device=dev, dtype=torch.float16)
    t = torch.tensor(500.0, device=dev)
    rs = {L: torch.zeros(B, L, 77, device=dev) for L in (4096, 1024, 256, 64)}
    for L in rs: rs[L][:, : L // 2, 1:3] = 0.5
    sig = torch.tensor(7.0, device=dev)
    kw = {"region_prompt": {"region_state": rs, "sigma": sig, "weight_func": reference_weight_func}}
    with torch.no_grad():
        for _ in range(3): y = net(x, t, ctx, cross_attention_kwargs=kw)
        torch.cuda.synchronize()
        if graph:
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2): y = net(x, t, ctx, cross_attention_kwargs=kw)
            torch.cuda.current_stream().wait_stream(s)
            with torch.cuda.graph(g):
                y = net(x, t, ctx, cross_attention_kwargs=kw)
            fn = g.replay
        else:
            fn = lambda: net(x, t, ctx, cross_attention_kwargs=kw)
        for _ in range(3): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"channels_last={channels_last} cudnn.benchmark={bench} graph={graph}: {dt*1e3:.2f} ms per UNet step (batch 16), finite={bool(torch.isfinite(y.float()).all())}", flush=True)
    if os.environ.get("DSC_UNET_PROFILE") and not graph:
        from torch.profiler import ProfilerActivity, profile
        with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
            net(x, t, ctx, cross_attention_kwargs=kw)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=90), flush=True)
    return y
outs = {}
for cfg in [(False, True, False), (True, True, False), (True, True, True), (False, True, True)]:
    try: outs[cfg] = run(*cfg).float()
    except Exception as e: print(cfg, "FAILED", repr(e)[:300], flush=True)
if (False, True, False) in outs and (True, True, False) in outs:
    a, b = outs[(False, True, False)], outs[(True, True, False)]
    print("channels_last (NHWC GroupNorm) vs NCHW (F.group_norm) output: cosine",
          float(torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0)), "max abs", float((a - b).abs().max()))
