#!/usr/bin/env python
"""bench.py -- images/s of the SD-1.5 512x512 25-step region-controlled generation (BASELINE.json metric),
with the masked cross-attention roofline and the CPU baseline next to it.

  python bench.py --gpus N --steps K --warmup W            (ours; N>1 under torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU path, rank 0 only)

A "step" is ONE pass of the hot path over one batch: a complete 25-step DPM++ 2M Karras generation of
8 images (attention batch 16 with CFG) = 400 region-masked cross-attention calls (one launch each: both passes
in one cooperative kernel) + 25 fused sampler steps + 16 K/V^T image builds.  Workload = BASELINE configs[1]: SD-1.5-architecture UNet, random init
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


def prepared_launches(B, H, L, D, S):
    """Launches one prepared-K/V call issues (1: both passes in one cooperative launch; 2 when that form is switched off)."""
    from diffusionspatialcontrol_b200 import _lib
    return int(_lib.lib.dsc_xattn_call_prepared_launches(B, H, L, D, S))


def lib_launches(B, H, L, D, S):
    from diffusionspatialcontrol_b200._lib import lib

    return int(lib.dsc_xattn_call_launches(B, H, L, D, S))


def cross_attention_shapes_list():
    from diffusionspatialcontrol_b200.unet_sd15 import cross_attention_shapes

    return list(cross_attention_shapes(HEIGHT, WIDTH))


def _eager_reference_call(q4, k4, v4, W_cpu, sigma):
    """The reference's eager op sequence for one region call on the same GPU, fp16, as its processor runs it
    (attention_modify.py:479-481 + :74-103): the region map is uploaded host -> device on EVERY call (:481)."""
    from oracle import attention as oa

    return oa.region_attention(q4, k4, v4, W_cpu.to(q4.device), sigma)


def _time_call_shape(device, B, H, L, D, S, flush, peak, n_timed=30, eager_reps=3):
    """CUDA-event time of ONE attention call as the processor issues it (L2 flushed before every timed call, two
    alternating input sets), plus the eager reference sequence on the same inputs."""
    from diffusionspatialcontrol_b200 import attention as att

    vw = lambda t: t.view(B, -1, H, D).transpose(1, 2)
    sets = []
    for i in range(2):
        q = torch.randn(B, L, H * D, device=device, dtype=torch.float16)
        k = torch.randn(B, S, H * D, device=device, dtype=torch.float16)
        v = torch.randn(B, S, H * D, device=device, dtype=torch.float16)
        W = torch.zeros(B, L, S, device=device)
        W[:, : L // 2, 1:3] = 0.5
        W = att.padded_region_map(W)  # the device layout encode_region_map produces (rows 80 floats apart)
        compact = att.compact_region_map(W)  # + the compact form the processor derives once per map
        prepared = att.prepared_supported(H, D, S, len(compact[1]))
        kv = att.prepare_kv(vw(k), vw(v), compact[1]) if prepared else None  # once per generation in the pipeline
        sets.append((q, k, v, W, compact, kv, torch.empty_like(q)))
    sigma = torch.tensor(7.0, device=device)
    ws = att.get_workspace(device, att.workspace_bytes(B, H, L, D, S))

    def call(t, passes=3):
        q, k, v, W, compact, kv, out = t
        if kv is not None:  # SD-1.5 layers with 40-wide heads: prepared K / V^T image (dsc_xattn_call_prepared)
            return att.region_attention_prepared(vw(q), kv, compact, sigma, workspace=ws, passes=passes, out=out)
        return att.region_attention(vw(q), vw(k), vw(v), W, sigma, workspace=ws, compact=compact)  # dsc_xattn_call_cw

    def timed(fn, n, warm=3):
        # the first `warm` iterations run the SAME flush + call sequence untimed: the first pass through a new sequence
        # (lazy module loads, allocator growth) leaves the host behind the GPU, and the gap between the start event and a
        # late launch would be counted as kernel time (measured: 158 us instead of 44 us for iteration 0)
        ts = []
        for it in range(-warm, n):
            flush.zero_()                                        # evict our inputs (512 MiB write) ...
            flush[: flush.numel() // 2].view(torch.int64).sum()  # ... and leave clean lines behind
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(sets[it % 2]); b.record(); b.synchronize()
            if it >= 0:
                ts.append(a.elapsed_time(b))
        return ts

    for i in range(6):
        call(sets[i % 2])
    tc = timed(call, n_timed)
    rec = {"B": B, "L": L, "D": D, "ms_call": sum(tc) / len(tc), "ms_call_median": sorted(tc)[len(tc) // 2], "n_calls": len(tc),
           "bytes": 2 * B * H * L * D * 3 + 2 * B * H * S * D * 3 + 4 * B * L * S,
           "path": "prepared" if sets[0][5] is not None else "raw", "launches": prepared_launches(B, H, L, D, S) if sets[0][5] is not None else lib_launches(B, H, L, D, S)}
    rec["frac"] = rec["bytes"] / (rec["ms_call"] * 1e-3) / 1e9 / peak
    if sets[0][5] is not None:
        rec["ms_stats_alone"] = sum(t1 := timed(lambda t: call(t, 1), 10)) / len(t1)
        rec["ms_forward_alone"] = sum(t2 := timed(lambda t: call(t, 2), 10)) / len(t2)
    if eager_reps:
        q, k, v, W, *_ = sets[0]
        W_cpu = W.cpu().contiguous()
        _eager_reference_call(vw(q), vw(k), vw(v), W_cpu, sigma)
        te = timed(lambda t: _eager_reference_call(vw(t[0]), vw(t[1]), vw(t[2]), W_cpu, sigma), eager_reps)
        rec["ms_eager_reference"] = sum(te) / len(te)
        rec["speedup_vs_eager"] = rec["ms_eager_reference"] / rec["ms_call"]
    return rec


def _time_call_shape_in_graph(device, B, H, L, D, S, peak, footprint=1024 << 20):
    """The call as the pipeline issues it: kernel nodes of ONE CUDA graph, back to back.  3 x n_sets calls over n_sets rotating
    input sets whose footprint exceeds the L2 eight times over (inputs larger than L2 instead of a flush kernel between the
    calls: when a set comes round again its lines have been evicted), one CUDA-event pair around the replay; per-call time =
    elapsed / calls, the gaps between consecutive launches included."""
    from diffusionspatialcontrol_b200 import attention as att

    vw = lambda t: t.view(B, -1, H, D).transpose(1, 2)
    per_set = 2 * B * L * H * D * 2 + B * L * 20 * 4
    n_sets = max(4, min(96, -(-footprint // per_set)))
    sets = []
    for i in range(n_sets):
        q = torch.randn(B, L, H * D, device=device, dtype=torch.float16)
        k = torch.randn(B, S, H * D, device=device, dtype=torch.float16)
        v = torch.randn(B, S, H * D, device=device, dtype=torch.float16)
        W = torch.zeros(B, L, S, device=device)
        W[:, : L // 2, 1:3] = 0.5
        compact = att.compact_region_map(att.padded_region_map(W))
        if not att.prepared_supported(H, D, S, len(compact[1])):
            return None
        sets.append((q, compact, att.prepare_kv(vw(k), vw(v), compact[1]), torch.empty_like(q)))
    sigma = torch.tensor(7.0, device=device)
    ws = torch.zeros(att.workspace_bytes(B, H, L, D, S), dtype=torch.uint8, device=device)
    call = lambda t: att.region_attention_prepared(vw(t[0]), t[2], t[1], sigma, workspace=ws, out=t[3])
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for t in sets:
            call(t)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    n, g = 3 * n_sets, torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            call(sets[i % n_sets])
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) / n)
    ms = sum(ts) / len(ts)
    nbytes = 2 * B * H * L * D * 3 + 2 * B * H * S * D * 3 + 4 * B * L * S
    del g
    return {"ms_call": ms, "frac": nbytes / (ms * 1e-3) / 1e9 / peak, "bytes": nbytes, "calls_per_replay": n, "replays": len(ts),
            "input_sets": n_sets, "footprint_bytes": n_sets * per_set}


def attention_roofline(device, sweep=True):
    """Live CUDA-event timing of the attention call (L2 flushed before every launch) on the dominant layer shape of the
    workload, the byte-weighted figure over all 16 layers of one UNet step, and (BASELINE configs[4] / configs[2]) the
    sweep over every SD-1.5 cross-attention shape x attention batch 2..32 and the 768 x 768 shapes, each next to the
    reference's eager op sequence on the same GPU."""
    from diffusionspatialcontrol_b200.unet_sd15 import cross_attention_shapes

    B, H, S = 2 * IMAGES_PER_UNIT, 8, 77
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=device)
    peak, how = peak_hbm()
    layer_shapes = list(cross_attention_shapes(HEIGHT, WIDTH))
    per_shape = {}
    for (L, D) in sorted(set(layer_shapes), reverse=True):
        per_shape[(L, D)] = _time_call_shape(device, B, H, L, D, S, flush, peak, n_timed=50, eager_reps=3)
    (L0, D0) = max(per_shape, key=lambda s: per_shape[s]["bytes"])
    d = per_shape[(L0, D0)]
    tot_b = sum(per_shape[s]["bytes"] for s in layer_shapes)
    tot_t = sum(per_shape[s]["ms_call"] for s in layer_shapes)
    tot_e = sum(per_shape[s]["ms_eager_reference"] for s in layer_shapes)
    in_graph = {}
    for (L, D) in sorted(set(layer_shapes), reverse=True):
        r = _time_call_shape_in_graph(device, B, H, L, D, S, peak)
        if r is not None:
            in_graph[(L, D)] = r
    traffic, traffic_src = None, None  # dram__bytes_read.sum + dram__bytes_write.sum of both passes, ncu --set full
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.isfile(tpath):
        tj = json.load(open(tpath))
        traffic, traffic_src = tj.get("dram_bytes_per_call"), "static: " + tj.get("source", "profiles/r2_traffic.json")
    out = {
        "bound": "hbm", "achieved": d["bytes"] / (d["ms_call"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": d["frac"],
        "traffic": traffic, "traffic_source": traffic_src, "peak_source": how,
        "kernel": "one attention call as the processor issues it: dsc_xattn_call_prepared = xattn_x3_fused_kernel, ONE cooperative "
                  "launch whose persistent CTAs run pass 1 (std of the scores) and pass 2 (softmax + P V) as two phases of "
                  "warp-specialised tcgen05 code over the K/V^T image prepared once per generation",
        "shape": {"B": B, "H": H, "L": L0, "D": D0, "S": S, "dtype": "f16"},
        "algorithmic_bytes_per_call": d["bytes"], "avg_ms_call": d["ms_call"], "median_ms_call": d["ms_call_median"],
        "timed_calls": d["n_calls"], "avg_ms_stats_alone": d.get("ms_stats_alone"), "avg_ms_forward_alone": d.get("ms_forward_alone"),
        "all_16_layers": {"bytes_per_unet_step": tot_b, "ms_per_unet_step": tot_t, "achieved": tot_b / (tot_t * 1e-3) / 1e9,
                          "frac": tot_b / (tot_t * 1e-3) / 1e9 / peak, "ms_eager_reference": tot_e,
                          "speedup_vs_eager": tot_e / tot_t},
        "per_shape": {f"{L}x{D}": {k: v[k] for k in ("ms_call", "bytes", "launches", "frac", "path", "ms_eager_reference",
                                                    "speedup_vs_eager")}
                      for (L, D), v in sorted(per_shape.items(), reverse=True)},
        "in_graph": None if len(in_graph) != len(set(layer_shapes)) else {
            "what": "the same call as kernel nodes of ONE CUDA graph (how the pipeline issues it), back to back over rotating input "
                    "sets with a footprint of >= 8 x L2 where the shape allows (inputs larger than L2, no flush kernel), one event pair around each replay; "
                    "per-call time = elapsed / calls, launch gaps included",
            "frac": in_graph[(L0, D0)]["frac"], "achieved": in_graph[(L0, D0)]["bytes"] / (in_graph[(L0, D0)]["ms_call"] * 1e-3) / 1e9,
            "avg_ms_call": in_graph[(L0, D0)]["ms_call"],
            "all_16_layers": {"ms_per_unet_step": (tg := sum(in_graph[s]["ms_call"] for s in layer_shapes)),
                              "frac": tot_b / (tg * 1e-3) / 1e9 / peak},
            "per_shape": {f"{L}x{D}": {k: v[k] for k in ("ms_call", "frac", "calls_per_replay", "replays", "input_sets", "footprint_bytes")}
                          for (L, D), v in sorted(in_graph.items(), reverse=True)}},
        "l2": "flushed before every timed call (512 MiB write, then a 256 MiB read so that L2 holds clean lines); two input "
              "sets alternate, so no timed call reads buffers the previous one touched; warm-up calls first",
    }
    if out["in_graph"] is not None:
        # Headline = the launch duration of the kernel where it runs: a kernel node of a CUDA graph, inputs larger than L2.
        # The round-1 figure -- ONE L2-flushed call between its own event pair -- stays next to it: that method measures an
        # EMPTY kernel at 5-6 us (scripts/launch_overhead.cu, profiles/r2_launch_overhead_empty_kernel.jsonl), i.e. it adds the event /
        # launch latency of a lone launch to every call; ncu's duration of the same kernel is 38.1 us per launch
        # (profiles/r2_ncu_launch_list_bench_summary.csv), the in-kernel span 33.5 us (profiles/r2_x3_span.jsonl).
        g = out["in_graph"]
        out["single_call_flushed"] = {
            "what": "the round-1 method: ONE call between its own CUDA-event pair, L2 flushed before it (512 MiB write + 256 MiB "
                    "read), two alternating input sets; includes the 5-6 us an empty kernel measures this way",
            "frac": out["frac"], "achieved": out["achieved"], "avg_ms_call": out["avg_ms_call"],
            "median_ms_call": out["median_ms_call"], "timed_calls": out["timed_calls"],
            "all_16_layers_frac": out["all_16_layers"]["frac"], "all_16_layers_ms": out["all_16_layers"]["ms_per_unet_step"]}
        out["frac"], out["achieved"], out["avg_ms_call"] = g["frac"], g["achieved"], g["avg_ms_call"]
        out["method"] = ("in_graph: average launch duration of the call as a kernel node of ONE CUDA graph (how the pipeline issues "
                         "it), 3 x n calls back to back over n rotating input sets with a footprint of >= 8 x L2 (inputs larger than "
                         "L2 instead of a flush kernel), 3 warm-up replays, one CUDA-event pair per timed replay, gaps between "
                         "consecutive launches included; single_call_flushed = the round-1 method, for comparison")
        out["all_16_layers"]["frac_in_graph"] = g["all_16_layers"]["frac"]
        out["all_16_layers"]["ms_per_unet_step_in_graph"] = g["all_16_layers"]["ms_per_unet_step"]
    def graph_cols(Bs, L, D):
        g = in_graph.get((L, D)) if Bs == B else _time_call_shape_in_graph(device, Bs, H, L, D, S, peak, footprint=768 << 20)
        return {} if g is None else {"ms_call_in_graph": g["ms_call"], "frac_in_graph": g["frac"]}

    if sweep:
        rows = []
        for (L, D) in sorted(set(layer_shapes), reverse=True):  # BASELINE configs[4]: every shape x attention batch 2..32
            for Bs in (2, 8, 16, 32):
                r = per_shape[(L, D)] if Bs == B else _time_call_shape(device, Bs, H, L, D, S, flush, peak, n_timed=20, eager_reps=2)
                rows.append({"config": 4, **{k: r[k] for k in ("B", "L", "D", "ms_call", "frac", "path", "launches",
                                                               "ms_eager_reference", "speedup_vs_eager")},
                             **graph_cols(Bs, L, D)})
        for (L, D) in ((9216, 40), (2304, 80), (576, 160), (144, 160)):  # BASELINE configs[2]: 768 x 768, batch 4 (+ CFG twin)
            r = _time_call_shape(device, 8, H, L, D, S, flush, peak, n_timed=20, eager_reps=2)
            rows.append({"config": 2, **{k: r[k] for k in ("B", "L", "D", "ms_call", "frac", "path", "launches",
                                                           "ms_eager_reference", "speedup_vs_eager")},
                         **graph_cols(8, L, D)})
        out["sweep"] = rows
        out["sweep_columns"] = ("ms_call / frac: one L2-flushed call between its own event pair (single_call_flushed method); "
                                "ms_call_in_graph / frac_in_graph: the call as a graph kernel node over inputs larger than L2 (the "
                                "method of roofline.frac)")
    return out


def _load_unet_module():
    """unet_sd15.py is plain PyTorch: load it by path so that the reference arm never imports the package (and so never
    maps libdsc_b200.so)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("_dsc_unet_sd15", os.path.join(ROOT, "diffusionspatialcontrol_b200", "unet_sd15.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu_baseline_sample(n_samples=1, warmup=0, n_img=IMAGES_PER_UNIT):
    """The reference's CPU path (oracle port: fp32 PyTorch UNet host + restated reference processor + restated DPM++ 2M
    loop), timed on a BOUNDED sample of the SAME workload as our arm: single denoising steps of the batch-`n_img`
    generation (UNet batch 2*n_img with CFG, 16 region-masked cross-attention calls per step) on every host core.
    images/s = n_img / (25 * t_step).  None of our kernels, engine or library is on this path."""
    from oracle import attention as oa
    from oracle import region_map as orm
    from oracle import sampler as osm

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    unet = _load_unet_module().UNetSD15().eval()
    unet.set_attn_processor(oa.OracleAttnProcessor())
    cond, uncond = prompt_embeds()
    ctx = torch.cat([uncond.expand(n_img, -1, -1), cond.expand(n_img, -1, -1)])
    rs = orm.encode_region_map(region_state_host(), lambda p: VOCAB[p], WIDTH, HEIGHT, n_img, text_ids=text_ids())
    train = osm.sd15_train_sigmas()
    sig = osm.get_sigmas_karras(DENOISE_STEPS, train[0].item(), train[-1].item())
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n_img, 4, HEIGHT // 8, WIDTH // 8, generator=g) * (sig[0] ** 2 + 1) ** 0.5
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
        "value": n_img / (DENOISE_STEPS * t_step), "unit": "images/s", "cores": cores, "kind": "port",
        "sample": f"{len(times)} denoising step(s) of the batch-{n_img} 512x512 generation (UNet batch {2 * n_img}, 16 region-masked "
                  f"cross-attention calls each), fp32, {t_step:.2f} s/step; images/s = {n_img} / (25 x s/step)",
        "seconds_per_denoise_step": t_step,
    }, times


REFERENCE_ARM_BUDGET_S = 180.0  # timed region of --impl reference: "the whole run ends within a few minutes"


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the reference is pure Python
    with no installable package) on our arm's workload.  A timed "step" here is ONE denoising step of the
    generation (1/25 of our arm's step) so that `--steps K` stays bounded: `ms_per_step` is the time of that step,
    `value` = batch / (25 x that time).  Batch 8 like our arm whenever K such steps fit the time budget (one probe step at
    batch 8 decides); otherwise the largest of 4 / 2 / 1 that fits, stated in `config` (the per-image CPU cost is nearly
    independent of the batch)."""
    if rank != 0:
        return
    steps = max(1, args.steps)
    probe, _ = cpu_baseline_sample(n_samples=1, warmup=0, n_img=IMAGES_PER_UNIT)
    t8 = probe["seconds_per_denoise_step"]
    n_img = IMAGES_PER_UNIT
    while n_img > 1 and steps * t8 * n_img / IMAGES_PER_UNIT > REFERENCE_ARM_BUDGET_S:
        n_img //= 2
    base, times = cpu_baseline_sample(n_samples=steps, warmup=max(0, min(args.warmup, 1)), n_img=n_img)
    t_step = base["seconds_per_denoise_step"]
    line = {
        "metric": METRIC, "value": base["value"], "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": f"BASELINE configs[1] workload on the host CPU: SD1.5-arch UNet random-init (seed 0), 512x512, 2 regions "
                               f"('A girl','bridge'), DPM++ 2M Karras 25 steps, CFG 7.5, batch {n_img} (attention batch {2 * n_img}), "
                               f"fp32, reference CPU path (oracle port)",
                   "step": f"ONE denoising step of the batch-{n_img} generation (UNet batch {2 * n_img}, 16 region-masked cross-attention "
                           f"calls): a bounded sample, 1/25 of a generation; value = {n_img} images / (25 x ms_per_step)",
                   "batch": n_img, "same_batch_as_gpu_arm": n_img == IMAGES_PER_UNIT,
                   "probe": f"one batch-{IMAGES_PER_UNIT} step took {t8:.1f} s = {IMAGES_PER_UNIT / (DENOISE_STEPS * t8):.4f} images/s; "
                            f"{steps} timed steps must fit {REFERENCE_ARM_BUDGET_S:.0f} s",
                   "timed_region": f"{len(times)} such steps, wall clock, all host cores"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def gpu_reference_sample(pipe, device, cond, uncond, ids, state, noise):
    """SURVEY 8d "GPU reference on the same box": the SAME pipeline, UNet, batch 8, fp16 and sampler with the reference's
    processor (oracle restatement of AttnProcessor2_0: eager PyTorch ops, region map re-uploaded host -> device on every
    call as attention_modify.py:481 does) installed instead of ours; eager launches as in the reference (no CUDA graph).
    One whole 25-step generation timed after one warm-up generation."""
    from diffusionspatialcontrol_b200.pipeline import RegionTxt2ImgPipeline, SyntheticTokenizer
    from oracle import attention as oa
    from oracle import region_map as orm

    rs_cpu = orm.encode_region_map(state, lambda p: VOCAB[p], WIDTH, HEIGHT, IMAGES_PER_UNIT, text_ids=ids)  # CPU tensors
    ours = pipe.processor
    ref_pipe = RegionTxt2ImgPipeline(pipe.unet, SyntheticTokenizer(VOCAB), processor=oa.OracleAttnProcessor(), use_cuda_graph=False)
    try:
        ms = []
        for it in range(2):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            x = ref_pipe.txt2img(cond, uncond, ids, None, noise, HEIGHT, WIDTH, DENOISE_STEPS, GUIDANCE,
                                 weight_func=oa.weight_func, region_state=rs_cpu)
            b.record()
            b.synchronize()
            ms.append(a.elapsed_time(b))
    finally:
        pipe.unet.set_attn_processor(ours)
    return {"value": IMAGES_PER_UNIT / (ms[-1] * 1e-3), "unit": "images/s", "ms_per_generation": ms[-1],
            "what": "same UNet / batch 8 / fp16 / sampler on this GPU with the reference processor (eager PyTorch ops, region "
                    "map re-uploaded per call) instead of ours, no CUDA graph; one 25-step generation after one warm-up"}, x


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

    last = {}

    def step_resident(i):
        x = pipe.txt2img(cond_d, uncond_d, ids, None, noises[i], HEIGHT, WIDTH, DENOISE_STEPS, GUIDANCE, region_state=rs_d)
        last["gathered"] = gather(x)
        return last["gathered"]

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
    # ---- matched outputs (north_star: "scaling ... at matched outputs"; unit definition SURVEY 8e) -----------------
    # after the timed region rank 0 recomputes, on ITS GPU, the unit another rank produced in the last timed step and
    # compares it with the slice that rank contributed to the gather: a unit's result must not depend on where it ran
    dp_match = None
    if world > 1:
        peer = world - 1
        if rank == 0:
            unit = peer + world * (total_steps - 1)
            redo = pipe.txt2img(cond_d, uncond_d, ids, None, unit_noise(unit, n_img, lat_shape).to(device), HEIGHT, WIDTH,
                                DENOISE_STEPS, GUIDANCE, region_state=rs_d).to(torch.float16)
            theirs = last["gathered"][peer]
            mine_again = pipe.txt2img(cond_d, uncond_d, ids, None, noises[total_steps - 1], HEIGHT, WIDTH, DENOISE_STEPS, GUIDANCE,
                                      region_state=rs_d).to(torch.float16)
            dp_match = {
                "bit_identical": bool(torch.equal(redo, theirs)),
                "cosine": float(torch.nn.functional.cosine_similarity(redo.float().flatten(), theirs.float().flatten(), dim=0)),
                "max_abs_diff": float((redo.float() - theirs.float()).abs().max()),
                "own_unit_repeat_bit_identical": bool(torch.equal(mine_again, last["gathered"][0])),
                "what": f"unit {unit} (seeds {unit * n_img}..{unit * n_img + n_img - 1}) computed by rank {peer} of {world} in the last "
                        f"timed step vs recomputed by rank 0 on its own GPU; fp16 latents as gathered",
            }
        dist.barrier()
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
        # our kernels inside the timed region: per denoising step the fused sampler step + the attention launches of the 16
        # cross-attention layers; per generation one K / V^T image build for each 40-wide-head layer
        "gpu_launches": args.steps * (DENOISE_STEPS * (1 + sum(
            prepared_launches(2 * IMAGES_PER_UNIT, 8, L, D, 77) for (L, D) in cross_attention_shapes_list()))
            + len(cross_attention_shapes_list())),
        "clocks": clocks,
        "impl": "dsc_b200",
    }
    if dp_match is not None:
        # "match" = the cross-rank recomputation agrees as well as a repeat on the same rank does: bit-identical when the
        # PyTorch host picked the same cuDNN / cuBLAS algorithms on both ranks (cudnn.benchmark autotunes per process)
        line["dp_outputs_match"] = bool(dp_match["bit_identical"] or dp_match["cosine"] >= 0.9999)
        line["dp_outputs"] = dp_match
    line["roofline"] = attention_roofline(device, sweep=(world == 1 and not args.no_sweep))
    if world == 1 and not args.no_gpu_reference:
        ref, x_ref = gpu_reference_sample(pipe, device, cond_d, uncond_d, ids, state, noises[total_steps - 1])
        x_ours = pipe.txt2img(cond_d, uncond_d, ids, None, noises[total_steps - 1], HEIGHT, WIDTH, DENOISE_STEPS, GUIDANCE,
                              region_state=rs_d)
        ref["final_latent_cosine_ours_vs_reference_processor"] = float(
            torch.nn.functional.cosine_similarity(x_ours.float().flatten(), x_ref.float().flatten(), dim=0))
        ref["speedup_e2e_over_gpu_reference"] = line["e2e"]["value"] / ref["value"]
        line["gpu_reference"] = ref
    if world == 1 and not args.no_cpu_baseline:
        base, _ = cpu_baseline_sample(n_samples=1, warmup=1)  # two batch-8 denoising steps on the host cores (~10-30 s)
        line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
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
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # (NCCL prints its one-line version banner on stdout; the JSON line is the last line)
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
