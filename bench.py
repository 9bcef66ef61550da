#!/usr/bin/env python
"""bench.py -- AttFind coord-evals/s @256px (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference algorithm on the host cores

Workload (BASELINE.json configs[2], the configuration the metric is quoted on; it fits one GPU):
StylEx 256px FFHQ-shaped generator (capacity 16, S = 4512 style coordinates) + ResNet-18@224 2-way classifier,
random-init + calibrated, synthetic latents.  One STEP = the AttFind sweep of `--latents-per-step` latents per rank:
every style coordinate x {towards min, towards max} = 9024 coord-evals per latent, each a perturbed generator
(suffix) forward + classifier forward + logit delta.  Weak scaling: every rank sweeps its own latents, no data-path
collective; the per-coordinate minima/maxima come from a fixed pool of latents prepared before the timed region
(phase A/B of the notebook, 1 generator forward per latent -- 0.01 % of the job).

value  = coord-evals of all ranks / device time (CUDA events, max over ranks), inputs resident in HBM.
e2e    = the same through the public API from HOST buffers: pinned latents + noise -> H2D, styles, base image/logits,
         sweep, effects -> D2H, all inside the timed region.
roofline = the Conv2DMod tcgen05 kernel: executed algorithmic FLOPs / CUDA-event time of those launches inside the
         timed region, against the measured sustained bf16 peak (MEASURED_PEAKS.json).
cpu_baseline = the oracle (a CPU port of the reference loop: batch 1, one full generator forward per coord-eval)
         on this box's host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "attfind_coord_evals_per_sec_256px"
UNIT = "coord-evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--image-size", type=int, default=256)
    ap.add_argument("--classifier", default="resnet", choices=["resnet", "mobilenet"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--classifier-dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--classifier-mode", default="fused", choices=["fused", "eager"],
                    help="fused: BatchNorm folded + PyTorch's fused cuDNN conv+bias+ReLU ops; eager: the module as is")
    ap.add_argument("--stem", default="s2d", choices=["s2d", "plain"],
                    help="fused classifier stem: the 7x7/stride-2 conv re-expressed as a 4x4/stride-1 conv on the space-to-depth input, or as is")
    ap.add_argument("--maxpool", default="native", choices=["native", "torch"],
                    help="fused classifier stem max-pool: sx_maxpool3x3s2_nhwc (bit-identical) or F.max_pool2d")
    ap.add_argument("--preprocess", default="native", choices=["native", "torch"],
                    help="classifier input pipeline: one native kernel or torchvision resize + Normalize")
    ap.add_argument("--latents-per-step", type=int, default=1)
    ap.add_argument("--pool", type=int, default=16, help="latents per rank prepared up front (minima/maxima pool)")
    ap.add_argument("--max-batch", type=int, default=256)
    ap.add_argument("--cpu-sample-coords", type=int, default=160, help="style coordinates in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="sx_clocks_", suffix=".csv")
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def build_workload(size, kind, seed=42):
    """seeded synthetic config: generator state dict, classifier model (CPU), noise."""
    from stylex_b200 import synthetic
    sd = synthetic.make_generator_state(size, seed=seed)
    model = synthetic.make_classifier_model(kind, seed)
    noise = synthetic.make_noise(size, seed)
    return sd, model, noise


def conv_flops_per_image(pairs):
    tot = 0.0
    for l, (ci, co) in enumerate(pairs):
        hw = (4 << l) ** 2
        tot += 2.0 * 9 * (ci * co + co * co) * hw
    return tot


# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(args, sd, model_cpu, noise, latents, minmax, n_coords, repeats=1):
    """The reference algorithm (oracle port: batch 1, full G forward + classifier per coord-eval) on the host cores."""
    import torch
    import stylex_b200 as sx
    from oracle import stylex_oracle as O

    clf = sx.make_classifier(args.classifier, model_cpu, args.image_size)
    S = O.num_style_coords(sd)
    step = max(1, S // n_coords)
    sind = list(range(0, S, step))[:n_coords]
    t0 = time.perf_counter()
    evals = 0
    for _ in range(repeats):
        O.attfind_sweep(sd, clf.classify_images, latents[:1], noise, sindices=sind,
                        minmax_from=torch.stack([minmax[0], minmax[1]]))
        evals += 2 * len(sind) + 1          # + the base image of the latent
    dt = time.perf_counter() - t0
    return evals / dt, dt, len(sind)


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the Python reference cannot travel to the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from stylex_b200 import synthetic
    from oracle import stylex_oracle as O

    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd, model, noise = build_workload(args.image_size, args.classifier)
    L = len(O.generator_layout(sd))
    lat = synthetic.make_latents(8, 42)
    import stylex_b200 as sx
    clf = sx.make_classifier(args.classifier, model, args.image_size)
    calib = O.generator_forward(sd, O.styles_def_to_tensor([(lat, L)]), noise)
    synthetic.calibrate_classifier(model, clf.preprocess, calib, chunk=8)
    _, sc = O.generator_forward(sd, O.styles_def_to_tensor([(lat, L)]), noise, get_style_coords=True)
    minmax = (sc.min(0).values, sc.max(0).values)
    n_coords = max(2, args.cpu_sample_coords // 2)
    for _ in range(args.warmup):
        cpu_reference_rate(args, sd, model, noise, lat, minmax, 2)
    t0 = time.perf_counter()
    evals = 0
    for _ in range(args.steps):
        r, dt, nc = cpu_reference_rate(args, sd, model, noise, lat, minmax, n_coords)
        evals += 2 * nc + 1
    dt = time.perf_counter() - t0
    value = evals / dt
    sample = f"per step: 1 latent x {n_coords} style coords (every {O.num_style_coords(sd) // n_coords}th) x 2 directions + base image, batch 1"
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"StylEx {args.image_size}px generator (S={O.num_style_coords(sd)}) + {args.classifier} classifier, AttFind sweep",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import stylex_b200 as sx
    from stylex_b200 import _native, synthetic
    from stylex_b200 import dist as sxd

    torch.set_grad_enabled(False)
    rank, world, local = sxd.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: stylex_b200 has no CPU path (use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    _native.device_check()
    torch.backends.cudnn.benchmark = True
    size, kind = args.image_size, args.classifier
    sd, model_cpu, noise_cpu = build_workload(size, kind)
    G = sx.Generator(size, 514).to(dev)
    G.load_state_dict(sd, strict=False)
    G.precision = args.precision
    plan = G.plan()
    S, L = plan.S, G.num_layers
    noise = noise_cpu.to(dev)

    # classifier: calibrate on generated images (fp32 everywhere), then switch to the throughput configuration
    clf = sx.make_classifier(kind, copy.deepcopy(model_cpu).to(dev), size)
    calib_lat = synthetic.make_latents(32, 7).to(dev)
    G.precision = "fp32"
    calib = torch.cat([G(sx.styles_def_to_tensor([(calib_lat[i:i + 8], L)]).contiguous(), noise) for i in range(0, 32, 8)])
    G.precision = args.precision
    synthetic.calibrate_classifier(clf.model, clf.preprocess, calib, chunk=8)
    model_cal_cpu = copy.deepcopy(clf.model).cpu().float()
    if args.classifier_dtype == "bf16":
        clf.set_compute(torch.bfloat16, channels_last=True)
    else:
        clf.set_compute(torch.float32, channels_last=True)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False

    clf_mode = "eager"
    if args.classifier_mode == "fused" and kind == "resnet":
        try:  # validate the fused forward against the eager module on real generated images before trusting it
            probe = calib[:8]
            ref_logits = clf.classify_images(probe)
            clf.fuse_for_inference()
            if args.stem == "s2d":
                clf.fused.enable_s2d_stem()
            if args.maxpool == "native":
                clf.fused.enable_native_pool()
            got = clf.classify_images(probe)
            torch.cuda.synchronize()
            tol = 0.05 * float(ref_logits.abs().max()) + 0.05
            if not torch.isfinite(got).all() or float((got - ref_logits).abs().max()) > tol:
                raise RuntimeError(f"fused classifier deviates: {float((got - ref_logits).abs().max()):.3e} > {tol:.3e}")
            clf_mode = ("fused (BN folded, aten::cudnn_convolution_[add_]relu" + (", 7x7/2 stem as 4x4/1 on space-to-depth input" if args.stem == "s2d" else "")
                        + (", native 3x3/2 max-pool)" if args.maxpool == "native" else ")"))
        except Exception as e:  # noqa: BLE001 -- any failure means: keep the eager module, say so in the JSON
            clf.fused = None
            clf_mode = f"eager (fused path unavailable: {type(e).__name__}: {str(e)[:120]})"

    pre_mode = "torch (resize, sub, div, cast, permute)"
    if args.preprocess == "native" and kind == "resnet":
        try:
            probe = calib[:8]
            ref_logits = clf.classify_images(probe)
            clf.use_native_preprocess(True)
            got = clf.classify_images(probe)
            torch.cuda.synchronize()
            tol = 0.03 * float(ref_logits.abs().max()) + 0.03
            if not torch.isfinite(got).all() or float((got - ref_logits).abs().max()) > tol:
                raise RuntimeError(f"native preprocess deviates: {float((got - ref_logits).abs().max()):.3e} > {tol:.3e}")
            pre_mode = "native (sx_resize_aa_normalize: antialiased resize + normalise + cast + NHWC in one kernel)"
        except Exception as e:  # noqa: BLE001
            clf.native_preprocess = False
            pre_mode = f"torch (native preprocess unavailable: {type(e).__name__}: {str(e)[:120]})"

    class TimedClassifier:
        """records CUDA events around every classifier call so the step breakdown can name the PyTorch share."""

        def __init__(self, inner):
            self.inner, self.events, self.on = inner, [], False

        def classify_images(self, images):
            if not self.on:
                return self.inner.classify_images(images)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = self.inner.classify_images(images)
            b.record()
            self.events.append((a, b))
            return out

        def total_ms(self):
            return sum(a.elapsed_time(b) for a, b in self.events)

    clf = TimedClassifier(clf)

    # latent pool of the whole job; every rank derives the same minima/maxima from ALL of it (no collective)
    pool_n = args.pool * world
    lat_all = synthetic.make_latents(pool_n, 42)
    lat_dev = lat_all.to(dev)
    styles_pool = plan.styles(sx.styles_def_to_tensor([(lat_dev, L)]).contiguous())
    minima, maxima = sx.get_min_max_style_vectors(styles_pool[:, :S].contiguous())
    my_lo, my_hi = sx.attfind.shard_range(pool_n, rank, world)
    lps = args.latents_per_step
    need = (args.warmup + 2 * args.steps) * lps
    my_idx = [my_lo + (i % (my_hi - my_lo)) for i in range(need)]

    def step_device(k):
        rows = lat_dev[[my_idx[k * lps + j] for j in range(lps)]]
        st = {}
        res = sx.attfind_sweep(G, clf, rows, noise, precision=args.precision, max_batch=args.max_batch, gather=False,
                               minmax=(minima, maxima), stats=st)
        return res, st["coord_evals"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(args.warmup):
        step_device(k)
    barrier()

    # ---------------- timed region: device resident ----------------
    sampler = ClockSampler(local)
    sampler.start()
    _native.profile_enable(True)
    clf.on = True
    launches0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ncu_range = bool(os.environ.get("SX_NCU_RANGE"))   # `ncu --profile-from-start off`: capture the timed region only
    if ncu_range:
        torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    evals = 0
    for k in range(args.steps):
        _, n = step_device(args.warmup + k)
        evals += n
    e1.record()
    barrier()
    if ncu_range:
        torch.cuda.cudart().cudaProfilerStop()
    ms = e0.elapsed_time(e1)
    launches = _native.launch_count() - launches0
    prof = _native.profile_collect()
    _native.profile_enable(False)
    clf.on = False
    clf_ms = clf.total_ms()
    clocks = sampler.stop()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    tot = torch.tensor([float(evals), float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    value = float(tot[0].item()) / (ms * 1e-3)

    # ---------------- e2e: host buffers, H2D + D2H inside the timed region ----------------
    e2e = None
    if not args.no_e2e:
        lat_pin = lat_all.pin_memory()
        noise_pin = noise_cpu.pin_memory()
        out_pin = torch.empty(lps, 2, S, 2).pin_memory()
        barrier()
        e0.record()
        evals2 = 0
        for k in range(args.steps):
            idx = [my_idx[(args.warmup + args.steps + k) * lps + j] for j in range(lps)]
            rows = lat_pin[idx].pin_memory().to(dev, non_blocking=True)
            nz = noise_pin.to(dev, non_blocking=True)
            st = {}
            res = sx.attfind_sweep(G, clf, rows, nz, precision=args.precision, max_batch=args.max_batch, gather=False,
                                   minmax=(minima, maxima), stats=st)
            out_pin.copy_(res["style_change"], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            evals2 += st["coord_evals"]
        e1.record()
        barrier()
        ms2 = e0.elapsed_time(e1)
        t2 = torch.tensor([ms2], device=dev, dtype=torch.float64)
        tot2 = torch.tensor([float(evals2)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot2, op=dist.ReduceOp.SUM)
        e2e = {"value": float(tot2.item()) / (float(t2.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(lps * 514 * 4 + size * size * 4), "d2h_bytes_per_step": int(lps * 2 * S * 2 * 4)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (Conv2DMod tcgen05 / FFMA) ----------------
    pk, pk_src = peaks()
    conv = {k: v for k, v in prof.items() if k < 32}
    conv_ms = sum(v["ms"] for v in conv.values())
    conv_flops = sum(v["flops"] for v in conv.values())
    conv_launches = sum(v["launches"] for v in conv.values())
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    peak = pk["bf16_tflops_sustained"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "conv_tc_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    bw = {k: v for k, v in prof.items() if k in (32, 33, 34, 37)}
    bw_ms = sum(v["ms"] for v in bw.values())
    bw_bytes = sum(v["bytes"] for v in bw.values())
    all_ms = sum(v["ms"] for v in prof.values())
    roofline = {
        "kernel": "conv_tc_kernel (Conv2DMod implicit GEMM, tcgen05/TMEM/TMA)" if args.precision == "bf16" else "conv_simt_kernel (FFMA)",
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
        "traffic": traffic, "peak_source": pk_src + ", sustained bf16 (kernel timed inside a long step)",
        "launches": conv_launches, "avg_launch_ms": conv_ms / conv_launches if conv_launches else None,
        "flops_per_launch": conv_flops / conv_launches if conv_launches else None,
        "share_of_step": conv_ms / ms if ms else None,
        "per_layer_tflops": {str(k): (v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else None) for k, v in sorted(conv.items())},
        "hbm_kernels": {"achieved_gbs": bw_bytes / (bw_ms * 1e-3) / 1e9 if bw_ms > 0 else None, "peak_gbs": pk["hbm_gbs"],
                        "frac": (bw_bytes / (bw_ms * 1e-3) / 1e9) / pk["hbm_gbs"] if bw_ms > 0 else None,
                        "share_of_step": bw_ms / ms if ms else None},
        "native_kernels_share_of_step": all_ms / ms if ms else None,
        "classifier_share_of_step": clf_ms / ms if ms else None,
        "kernel_ms_per_step": {{32: "modulate", 33: "upsample2x_modulate", 34: "torgb", 35: "demod", 37: "rgb_prev_up_blur"}.get(k, f"conv{k}"):
                               round(v["ms"] / max(1, args.steps), 3) for k, v in sorted(prof.items())},
    }

    # ---------------- CPU baseline (reported only) ----------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        torch.backends.cudnn.allow_tf32 = False
        rate, dt, nc = cpu_reference_rate(args, sd, model_cal_cpu, noise_cpu, lat_all, (minima.cpu(), maxima.cpu()),
                                          args.cpu_sample_coords)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"1 latent x {nc} style coords (every {S // nc}th) x 2 directions + base image, batch-1 full forwards, {dt:.1f} s"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {"workload": f"StylEx {size}px FFHQ-shaped generator (capacity 16, S={S}) + {kind}-18@224 classifier; "
                               f"AttFind sweep, {lps} latent(s)/rank/step x {S} coords x 2 directions",
                   "coord_evals_per_step": int(2 * S * lps * world), "max_batch": args.max_batch,
                   "generator_dtype": args.precision, "classifier_dtype": args.classifier_dtype + " (PyTorch, channels_last)", "classifier_mode": clf_mode, "classifier_preprocess": pre_mode,
                   "prefix_reuse": True, "l2": f"inputs larger than L2: every {args.max_batch}-eval batch streams >{args.max_batch * 16.8e6 / 1e9:.1f} GB of activations (L2 = 126 MB)",
                   "pool_latents": pool_n, "parallelism": f"latent-sharded x{world}, no data-path collective"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(tot[1].item()), "roofline": roofline, "cpu_baseline": cpu,
        "conv_flops_per_full_image": conv_flops_per_image(plan.pairs),
    }
    emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _claim_stdout():
    """Everything third parties print to stdout (e.g. NCCL's version banner) goes to stderr; the ONE JSON line is written
    to the real stdout by emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT, line)


if __name__ == "__main__":
    a = parse()
    _claim_stdout()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
