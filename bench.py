#!/usr/bin/env python
"""bench.py -- images/s of the SD-1.5 512x512 25-step region-controlled generation (BASELINE.json metric),
with the masked cross-attention roofline and the CPU baseline next to it.

  python bench.py --gpus N --steps K --warmup W            (ours; N>1 under torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU path, rank 0 only)

A "step" is ONE pass of the hot path over one batch: a complete 25-step DPM++ 2M Karras generation of
8 images (attention batch 16 with CFG) = 400 region-masked cross-attention calls (800 launches of our two
passes) + 25 fused sampler steps.  Workload = BASELINE configs[1]: SD-1.5-architecture UNet, random init
(seed 0), fp16, 512x512, 2 regions ('A girl', 'bridge'), CFG 7.5, synthetic prompt embeddings.  Latents
only (no VAE).  `value` has inputs resident in HBM; `e2e` goes through the public pipeline call with
pinned HOST buffers (noise, embeddings, region masks -> H2D, region maps rebuilt on the device, final
latents -> D2H) inside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HEIGHT = WIDTH = 512
IMAGES_PER_UNIT = 8
DENOISE_STEPS = 25
GUIDANCE = 7.5
PROMPT_IDS = [49406, 320, 1611, 4919, 525, 518, 2465] + [49407] * 70
VOCAB = {"A girl": [320, 1611], "bridge": [2465]}
METRIC = "images/s (SD1.5 512x512, 25 steps DPM++ 2M Karras, CFG 7.5, 2 regions)"


def region_state_host():
    import numpy as np

    m1 = np.full((HEIGHT, WIDTH), 255, np.uint8)
    m1[178:326, 50:305] = 0
    m2 = np.full((HEIGHT, WIDTH), 255, np.uint8)
    m2[317:471, 52:300] = 0
    return {"A girl": {"map": m1, "weight": 0.5, "mask_outsides": 0.0},
            "bridge": {"map": m2, "weight": 0.7, "mask_outsides": 0.0}}


def prompt_embeds():
    g1, g2 = torch.Generator().manual_seed(1), torch.Generator().manual_seed(2)
    return torch.randn(1, 77, 768, generator=g1), torch.randn(1, 77, 768, generator=g2)  # cond, uncond


def text_ids():
    import numpy as np

    return [np.array([[49406] + [49407] * 76]), np.array([PROMPT_IDS])]


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, val in zip(names, f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------- ours
def build_pipeline(device):
    from diffusionspatialcontrol_b200.pipeline import RegionTxt2ImgPipeline, SyntheticTokenizer
    from diffusionspatialcontrol_b200.unet_sd15 import UNetSD15

    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True  # PyTorch-side conv autotuning for the UNet host (not our kernels)
    unet = UNetSD15().to(device=device, dtype=torch.float16).eval()
    if not os.environ.get("DSC_BENCH_NO_CHANNELS_LAST"):
        unet = unet.to(memory_format=torch.channels_last)
    return RegionTxt2ImgPipeline(unet, SyntheticTokenizer(VOCAB), use_cuda_graph=not os.environ.get("DSC_BENCH_NO_GRAPH"))


def lib_launches(B, H, L, D, S):
    from diffusionspatialcontrol_b200._lib import lib

    return int(lib.dsc_xattn_call_launches(B, H, L, D, S))


def cross_attention_shapes_list():
    from diffusionspatialcontrol_b200.unet_sd15 import cross_attention_shapes

    return list(cross_attention_shapes(HEIGHT, WIDTH))


def attention_roofline(device):
    """Live CUDA-event timing of the two attention passes (L2 flushed before every launch) on the dominant
    layer shape of the workload, plus the byte-weighted figure over all 16 layers of one UNet step."""
    from diffusionspatialcontrol_b200 import attention as att
    from diffusionspatialcontrol_b200._lib import check, lib
    from diffusionspatialcontrol_b200.unet_sd15 import cross_attention_shapes

    I4, I3 = ctypes.c_int64 * 4, ctypes.c_int64 * 3
    B, H, S = 2 * IMAGES_PER_UNIT, 8, 77
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=device)
    st = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    ws = att.get_workspace(device)
    peak, how = peak_hbm()
    per_shape = {}
    for (L, D) in sorted(set(cross_attention_shapes(HEIGHT, WIDTH)), reverse=True):
        vw = lambda t: t.view(B, -1, H, D).transpose(1, 2)
        sc = 1 / math.sqrt(D)
        sets = []  # two input/output sets used alternately: a timed launch never sees buffers its predecessor touched
        for i in range(2):
            q = torch.randn(B, L, H * D, device=device, dtype=torch.float16)
            k = torch.randn(B, S, H * D, device=device, dtype=torch.float16)
            v = torch.randn(B, S, H * D, device=device, dtype=torch.float16)
            W = torch.zeros(B, L, S, device=device)
            W[:, : L // 2, 1:3] = 0.5
            W = att.padded_region_map(W)  # the device layout encode_region_map produces (rows 80 floats apart)
            Wc, cols = att.compact_region_map(W)  # + the compact form the processor derives once per map (2 weighted columns)
            sets.append((q, k, v, W, torch.empty_like(q), Wc))
        qs, ks, vs = I4(*vw(sets[0][0]).stride()), I4(*vw(sets[0][1]).stride()), I4(*vw(sets[0][2]).stride())
        os_ = I3(*sets[0][4].stride())

        cols_arr = (ctypes.c_int32 * len(cols))(*cols)

        def k1(t):
            q, k, v, W, out, Wc = t
            check(lib.dsc_xattn_stats(q.data_ptr(), k.data_ptr(), qs, ks, None, B, H, L, D, S, sc, 0, ws.data_ptr(), st))

        def k2(t):
            q, k, v, W, out, Wc = t
            check(lib.dsc_xattn_forward(q.data_ptr(), k.data_ptr(), v.data_ptr(), qs, ks, vs, W.data_ptr(), B, W.stride(1), None, 7.0,
                                        ws.data_ptr(), out.data_ptr(), os_, B, H, L, D, S, sc, 0, st))

        def call(t):  # one attention call through the C ABI: pass 1 + pass 2 (pass 2 a programmatic dependent launch of
            q, k, v, W, out, Wc = t  # pass 1), or ONE fused cooperative launch where the problem fits on chip (small layers)
            check(lib.dsc_xattn_call_cw(q.data_ptr(), k.data_ptr(), v.data_ptr(), qs, ks, vs, W.data_ptr(), B, W.stride(1),
                                        Wc.data_ptr(), len(cols), cols_arr, None, 7.0, ws.data_ptr(), out.data_ptr(), os_,
                                        B, H, L, D, S, sc, 0, st))

        for i in range(10):
            call(sets[i % 2])
        t1, t2, tc = [], [], []
        n_launch = 0
        for it in range(50):
            for fn, acc in ((k1, t1), (k2, t2), (call, tc)):
                if fn is not call and it >= 20:
                    continue  # the single passes are informational: 20 samples each
                flush.zero_()                                  # evict our inputs (512 MiB write) ...
                flush[: flush.numel() // 2].view(torch.int64).sum()  # ... and leave clean lines behind
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(sets[n_launch % 2]); b.record(); b.synchronize()
                n_launch += 1
                acc.append(a.elapsed_time(b))
        nbytes = 2 * B * H * L * D * 3 + 2 * B * H * S * D * 3 + 4 * B * L * S
        per_shape[(L, D)] = {"ms_stats": sum(t1) / len(t1), "ms_forward": sum(t2) / len(t2), "ms_call": sum(tc) / len(tc),
                             "ms_call_median": sorted(tc)[len(tc) // 2], "n_calls": len(tc), "bytes": nbytes}
        del sets
    (L0, D0) = max(per_shape, key=lambda s: per_shape[s]["bytes"])
    d = per_shape[(L0, D0)]
    ach = d["bytes"] / (d["ms_call"] * 1e-3) / 1e9
    tot_b = sum(per_shape[s]["bytes"] for s in cross_attention_shapes(HEIGHT, WIDTH))
    tot_t = sum(per_shape[s]["ms_call"] for s in cross_attention_shapes(HEIGHT, WIDTH))
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum of both passes from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.isfile(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_call")
    return {
        "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
        "peak_source": how,
        "kernel": "one attention call = dsc_xattn_call_cw, the call the processor makes (at this shape: two tcgen05 kernels, pass 2 a "
                  "programmatic dependent launch of pass 1 and fed with the compact region map), one CUDA-event pair around it",
        "shape": {"B": B, "H": H, "L": L0, "D": D0, "S": S, "dtype": "f16"},
        "algorithmic_bytes_per_call": d["bytes"], "avg_ms_call": d["ms_call"], "median_ms_call": d["ms_call_median"],
        "timed_calls": d["n_calls"],
        "avg_ms_stats_alone": d["ms_stats"], "avg_ms_forward_alone": d["ms_forward"],
        "all_16_layers": {"bytes_per_unet_step": tot_b, "ms_per_unet_step": tot_t,
                          "achieved": tot_b / (tot_t * 1e-3) / 1e9, "frac": tot_b / (tot_t * 1e-3) / 1e9 / peak},
        "per_shape": {f"{L}x{D}": {"ms_call": v["ms_call"], "bytes": v["bytes"], "launches": lib_launches(B, H, L, D, S),
                                   "frac": v["bytes"] / (v["ms_call"] * 1e-3) / 1e9 / peak}
                      for (L, D), v in sorted(per_shape.items(), reverse=True)},
        "l2": "flushed before every timed call (512 MiB write, then a 256 MiB read so that L2 holds clean lines); two input "
              "sets alternate, so no timed call reads buffers the previous one touched; 10 warm-up calls",
    }


def cpu_baseline_sample(n_samples=1, warmup=0):
    """The reference's CPU path (oracle port: fp32 PyTorch UNet host + restated reference processor +
    restated DPM++ 2M loop), timed on a BOUNDED sample: one denoising step of a batch-1 generation
    (UNet batch 2 with CFG, 16 region-masked cross-attention calls).  images/s = 1 / (25 * t_step)."""
    from diffusionspatialcontrol_b200.unet_sd15 import UNetSD15
    from oracle import attention as oa
    from oracle import region_map as orm
    from oracle import sampler as osm

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    unet = UNetSD15().eval()
    unet.set_attn_processor(oa.OracleAttnProcessor())
    cond, uncond = prompt_embeds()
    ctx = torch.cat([uncond, cond])
    rs = orm.encode_region_map(region_state_host(), lambda p: VOCAB[p], WIDTH, HEIGHT, 1, text_ids=text_ids())
    train = osm.sd15_train_sigmas()
    sig = osm.get_sigmas_karras(DENOISE_STEPS, train[0].item(), train[-1].item())
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4, HEIGHT // 8, WIDTH // 8, generator=g) * (sig[0] ** 2 + 1) ** 0.5
    times = []

    def eps_fn(x_in, sigma):
        t = osm.sigma_to_t(sigma.reshape(1), train.log())
        rp = {"region_state": rs, "sigma": sigma, "weight_func": oa.weight_func}
        return unet(x_in, t, ctx, cross_attention_kwargs={"region_prompt": rp})

    with torch.no_grad():
        for i in range(warmup + n_samples):
            t0 = time.perf_counter()
            osm.cfg_denoise(eps_fn, x, sig[min(i, DENOISE_STEPS - 1)], GUIDANCE)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    t_step = sum(times) / len(times)
    return {
        "value": 1.0 / (DENOISE_STEPS * t_step), "unit": "images/s", "cores": cores, "kind": "port",
        "sample": f"{len(times)} denoising step(s) of a batch-1 512x512 generation (UNet batch 2, 16 region-masked "
                  f"cross-attention calls each), fp32, {t_step:.2f} s/step, extrapolated to 25 steps",
        "seconds_per_denoise_step": t_step,
    }, times


def run_reference(args, rank, world):
    if rank != 0:
        return
    base, times = cpu_baseline_sample(n_samples=max(1, args.steps), warmup=max(0, min(args.warmup, 1)))
    t_step = base["seconds_per_denoise_step"]
    line = {
        "metric": METRIC, "value": base["value"], "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_step * DENOISE_STEPS, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": "SD1.5-arch UNet random-init, 512x512, 2 regions, DPM++ 2M Karras 25 steps, CFG 7.5, "
                               "batch 1, reference CPU path (oracle port), bounded sample per step"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist

    from diffusionspatialcontrol_b200.distributed import unit_noise
    from diffusionspatialcontrol_b200.region_map import encode_region_map

    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    pipe = build_pipeline(device)
    cond, uncond = prompt_embeds()
    ids = text_ids()
    state = region_state_host()
    n_img = IMAGES_PER_UNIT
    lat_shape = (4, HEIGHT // 8, WIDTH // 8)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(x):
        if world > 1:
            out = torch.empty((world, *x.shape), dtype=torch.float16, device=device)
            dist.all_gather_into_tensor(out, x.to(torch.float16).contiguous())  # the path's one collective
            return out
        return x

    # ---- resident-input arm ------------------------------------------------------------------
    cond_d, uncond_d = cond.to(device), uncond.to(device)
    rs_d = encode_region_map(pipe, state, WIDTH, HEIGHT, n_img, text_ids=ids, device=device)
    total_steps = args.warmup + args.steps
    noises = [unit_noise(rank + world * i, n_img, lat_shape).to(device) for i in range(total_steps)]

    def step_resident(i):
        x = pipe.txt2img(cond_d, uncond_d, ids, None, noises[i], HEIGHT, WIDTH, DENOISE_STEPS, GUIDANCE, region_state=rs_d)
        return gather(x)

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.warmup, total_steps):
        step_resident(i)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)

    # ---- end-to-end arm: host buffers in, host buffers out --------------------------------------
    host_noise = [unit_noise(rank + world * i, n_img, lat_shape).pin_memory() for i in range(total_steps)]
    host_cond, host_uncond = cond.pin_memory(), uncond.pin_memory()
    host_out = torch.empty((n_img, *lat_shape), dtype=torch.float32).pin_memory()
    h2d = host_noise[0].numel() * 4 + host_cond.numel() * 4 * 2 + sum(v["map"].nbytes for v in state.values())
    d2h = host_out.numel() * 4

    def step_e2e(i):
        nz = host_noise[i].to(device, non_blocking=True)
        c, u = host_cond.to(device, non_blocking=True), host_uncond.to(device, non_blocking=True)
        x = pipe.txt2img(c, u, ids, state, nz, HEIGHT, WIDTH, DENOISE_STEPS, GUIDANCE)  # uploads masks, builds W on the device
        gather(x)
        host_out.copy_(x, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for i in range(min(args.warmup, 1)):
        step_e2e(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.warmup, total_steps):
        step_e2e(i)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        return
    images = world * n_img * args.steps
    line = {
        "metric": METRIC, "value": images / (ms * 1e-3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {
            "workload": "BASELINE configs[1]: SD1.5-arch UNet random-init (seed 0), 512x512, 2 regions ('A girl','bridge'), "
                        "DPM++ 2M Karras 25 steps, CFG 7.5, batch 8 per GPU (attention batch 16), fp16; latents only (no VAE)",
            "step": "one 25-step generation of 8 images per GPU (400 attention calls, 25 sampler steps)",
            "images_per_step_per_gpu": n_img, "parallelism": f"dp{world} (seed-batch sharding, one all_gather of latents per step)",
            "l2": "working set per UNet step (~1.7 GB weights + activations) far exceeds the 126 MB L2; the roofline "
                  "microbench flushes L2 before every timed launch",
        },
        "e2e": {"value": images / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": args.steps * DENOISE_STEPS * (1 + sum(
            lib_launches(2 * IMAGES_PER_UNIT, 8, L, D, 77) for (L, D) in cross_attention_shapes_list())),
        "clocks": clocks,
        "impl": "dsc_b200",
    }
    line["roofline"] = attention_roofline(device)
    if world == 1 and not args.no_cpu_baseline:
        base, _ = cpu_baseline_sample(n_samples=2, warmup=1)
        line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        with torch.no_grad():
            run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
