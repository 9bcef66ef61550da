#!/usr/bin/env python
"""BASELINE config 4: generator-only Conv2DMod stack throughput at 256px, batch 64, bf16 vs fp32 (tensor-pipe roofline
sweep).  Full `Generator.forward` through the plan (21 Conv2DMod calls: 14 3x3 demod + 7 1x1 ToRGB), CUDA events on the
launching stream, per-layer times from sx_profile (events around every launch).  One JSON line per precision.

    python profiles/bench_generator_only.py --batch 64 --iters 20 --out gpurun_out/config4.jsonl
    # per-layer tensor-pipe utilisation (a run under ncu is never a bench value):
    ncu --set full --clock-control none -k regex:"conv_tc|conv_simt" -s <warm-up conv launches> -c 14 -o gpurun_out/cfg4 \
        python profiles/bench_generator_only.py --precisions bf16 --iters 1 --warmup 2
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import stylex_b200 as sx
from stylex_b200 import _native, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--precisions", default="bf16,fp32")
ap.add_argument("--out", default=None)
a = ap.parse_args()
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
_native.device_check()
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    peak_src = "measured (MEASURED_PEAKS.json)"
except Exception:
    peaks = {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}
    peak_src = "fallback (B200_PROFILING.md)"
G = sx.Generator(a.size, 514).to(dev)
G.load_state_dict(synthetic.make_generator_state(a.size, seed=42), strict=False)
plan = G.plan()
lat = synthetic.make_latents(a.batch, 42).to(dev)
noise = synthetic.make_noise(a.size, 42).to(dev)
styles = plan.styles(sx.styles_def_to_tensor([(lat, G.num_layers)]).contiguous())
flops_img = sum(2.0 * 9 * (ci * co + co * co) * (4 << l) ** 2 for l, (ci, co) in enumerate(plan.pairs))
for prec in a.precisions.split(","):
    for _ in range(a.warmup):
        img = plan.forward(styles, noise, precision=prec)
    torch.cuda.synchronize()
    _native.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        img = plan.forward(styles, noise, precision=prec)
    e1.record()
    torch.cuda.synchronize()
    prof = _native.profile_collect()
    _native.profile_enable(False)
    ms = e0.elapsed_time(e1) / a.iters
    conv = {k: v for k, v in prof.items() if k < 32}
    conv_ms = sum(v["ms"] for v in conv.values()) / a.iters
    conv_tf = sum(v["flops"] for v in conv.values()) / a.iters / (conv_ms * 1e-3) / 1e12
    names = {32: "modulate", 33: "upsample2x_modulate", 34: "torgb", 35: "demod", 37: "rgb_prev_up_blur"}
    peak = peaks["bf16_tflops"]            # burst figure: a ~10 ms forward timed alone
    line = {
        "metric": "generator_forward_images_per_sec_%dpx" % a.size, "value": a.batch / (ms * 1e-3), "unit": "images/s",
        "config": {"workload": "BASELINE config 4: generator-only forward, %dpx, batch %d, 14 Conv2DMod 3x3 + 7 ToRGB" % (a.size, a.batch),
                   "precision": prec, "iters": a.iters},
        "ms_per_forward": ms, "checksum": float(img.float().abs().mean()),
        "conv_tflops": conv_tf, "conv_ms": conv_ms, "conv_share": conv_ms / ms,
        "whole_forward_tflops": flops_img * a.batch / (ms * 1e-3) / 1e12,
        "roofline": {"bound": "tensor" if prec == "bf16" else "CUDA-core FFMA (tcgen05 has no fp32 MMA)", "achieved": conv_tf,
                     "peak": peak, "unit": "TFLOP/s", "frac": conv_tf / peak if prec == "bf16" else None,
                     "peak_source": peak_src + ", burst bf16"},
        "per_layer_tflops": {"conv%d" % k: round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) for k, v in sorted(conv.items()) if v["ms"] > 0},
        "per_layer_ms": {names.get(k, "conv%d" % k): round(v["ms"] / a.iters, 4) for k, v in sorted(prof.items())},
    }
    s = json.dumps(line)
    print(s, flush=True)
    if a.out:
        with open(a.out, "a") as f:
            f.write(s + "\n")
