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
cpu_baseline = the reference's own loop on this box's host cores, on a bounded sample of the same workload: the VERBATIM
         notebook cell 5 (`attfind_extraction`, NB:269-417) over the reference's own Generator / classifier wrapper when
         `oracle/make_ref.sh` staged the reference files under oracle/_ref (kind "reference"), else the oracle port
         (kind "port").  `--impl reference` times exactly the same thing, step by step.
job    = (default on, `--no-job` to skip) ONE WHOLE AttFind job timed by the wall clock through the public API: host
         latents -> H2D -> sweep of `--job-latents` latents per rank (128 x 8 ranks = BASELINE config 3; 256 x 1 at 64px =
         config 2) -> the NCCL all-gather of the effects -> class split + greedy top-k on device -> picks on the host.
         Reported beside the step rate so the gather, the selection, the per-job setup and the tail are all inside.
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

UNIT = "coord-evals/s"


def metric_name(size):
    return f"attfind_coord_evals_per_sec_{size}px"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--image-size", type=int, default=256)
    ap.add_argument("--classifier", default="resnet", choices=["resnet", "mobilenet"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--classifier-dtype", default=None, choices=["bf16", "fp32"],
                    help="default: bf16 for the ResNet-18 wrapper, fp32 for MobileNetV2 (0.05 GFLOP per image, and its bf16 logits "
                         "are off by ~1: profiles/README.md config 1)")
    ap.add_argument("--classifier-mode", default="fused", choices=["fused", "eager"],
                    help="fused: BatchNorm folded + PyTorch's fused cuDNN conv+bias+ReLU ops; eager: the module as is")
    ap.add_argument("--stem", default="native", choices=["native", "s2d", "plain"],
                    help="fused classifier stem: the 7x7/stride-2 conv re-expressed as a 4x4/stride-1 conv on the space-to-depth input "
                         "(native: in the tcgen05 kernel sx_stem_s2d_conv_relu, with the max-pool fused; s2d: through cuDNN), or as is")
    ap.add_argument("--maxpool", default="native", choices=["native", "torch"],
                    help="fused classifier stem max-pool: sx_maxpool3x3s2_nhwc (bit-identical) or F.max_pool2d")
    ap.add_argument("--preprocess", default="native", choices=["native", "torch"],
                    help="classifier input pipeline: one native kernel or torchvision resize + Normalize")
    ap.add_argument("--latents-per-step", type=int, default=1)
    ap.add_argument("--pool", type=int, default=16, help="latents per rank prepared up front (minima/maxima pool)")
    ap.add_argument("--max-batch", type=int, default=None,
                    help="perturbed images per launch; default attfind.default_eval_batch(image size): 256 at 256 px, 1024 at 64 px")
    ap.add_argument("--cpu-sample-coords", type=int, default=60, help="style coordinates in the CPU-baseline sample")
    ap.add_argument("--cpu-sample-images", type=int, default=2, help="images in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-job", action="store_true", help="skip the whole-job leg")
    ap.add_argument("--no-verify", action="store_true", help="whole-job leg without the fp32 verification of the top-k candidates")
    ap.add_argument("--job-latents", type=int, default=0,
                    help="latents PER RANK of the whole-job leg (default: 128 at 256px = config 3 on 8 ranks; 256 at 64px = config 2)")
    ap.add_argument("--out", default=None, help="also append the JSON line to this file")
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
def cpu_sample_sindices(S, n_coords):
    step = max(1, S // n_coords)
    return list(range(0, S, step))[:n_coords]


class CpuReference:
    """The reference's CPU path on a bounded sample: `n_images` images x `n_coords` strided style coordinates x 2
    directions, batch 1, one FULL generator forward + one classifier forward per coord-eval.

    kind "reference": the unmodified reference files staged by oracle/make_ref.sh -- `attfind_extraction` of notebook
    cell 5 exec'd verbatim (phase A included: encoder stand-in that returns the preset w, classifier, generator), the
    reference `Generator` and `ResNet`/`MobileNet.classify_images`.  kind "port": oracle.stylex_oracle.attfind_sweep."""

    def __init__(self, args, sd, model_cpu, noise):
        import torch
        from oracle import ref_loader as RL
        from oracle import stylex_oracle as O
        self.args, self.sd, self.noise = args, sd, noise
        self.S = O.num_style_coords(sd)
        self.kind = "reference" if RL.available() else "port"
        self.O, self.RL = O, RL
        if self.kind == "reference":
            self.G = RL.reference_generator(sd, args.image_size)
            self.clf = RL.reference_classifier(args.classifier, model_cpu, args.image_size)
        else:
            import stylex_b200 as sx
            self.clf = sx.make_classifier(args.classifier, model_cpu, args.image_size)
        g = torch.Generator().manual_seed(1234)
        self.images = torch.rand(max(1, args.cpu_sample_images), 3, args.image_size, args.image_size, generator=g)

    def describe(self, n_coords):
        a = self.args
        n_img = self.images.shape[0]
        what = ("VERBATIM notebook cell 5 (attfind_extraction) over the reference Generator + classify_images"
                if self.kind == "reference" else "oracle port of the notebook loop")
        return (f"{what}: {n_img} image(s) x {n_coords} style coords (every {max(1, self.S // n_coords)}th) x 2 directions + "
                f"{n_img} base image(s), batch-1 full forwards, {a.classifier}{'-18@224' if a.classifier == 'resnet' else 'V2'} fp32, "
                f"{a.image_size}px")

    def run(self, latents, n_coords):
        """-> (coord-evals/s, seconds, coord-evals)"""
        import torch
        sind = cpu_sample_sindices(self.S, n_coords)
        n_img = self.images.shape[0]
        w = latents[:n_img, :512].clone()
        t0 = time.perf_counter()
        if self.kind == "reference":
            calls = {"i": 0}

            def encoder(batch):                      # stands in for stylex.encoder (NB:306): returns the preset w
                i = calls["i"] % n_img
                calls["i"] += 1
                return w[i]
            self.RL.run_reference_attfind(self.G, self.clf, self.images, encoder, self.noise, self.S, shift_size=1.0,
                                          sindex_subset=sind, results_folder="mem://bench")
        else:
            lat = torch.cat([w, self.clf.classify_images(self.images)], dim=1)
            self.O.attfind_sweep(self.sd, self.clf.classify_images, lat, self.noise, sindices=sind)
        dt = time.perf_counter() - t0
        evals = n_img * (2 * len(sind) + 1)
        return evals / dt, dt, evals


def calibrated_cpu_model(args, sd, model, noise):
    """the seeded classifier calibrated on 8 generated images (CPU path of the --impl reference arm)."""
    from stylex_b200 import synthetic
    from oracle import stylex_oracle as O
    import stylex_b200 as sx
    L = len(O.generator_layout(sd))
    lat = synthetic.make_latents(8, 42)
    clf = sx.make_classifier(args.classifier, model, args.image_size)
    calib = O.generator_forward(sd, O.styles_def_to_tensor([(lat, L)]), noise)
    synthetic.calibrate_classifier(model, clf.preprocess, calib, chunk=8)
    return model, lat


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (see CpuReference)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd, model, noise = build_workload(args.image_size, args.classifier)
    model, lat = calibrated_cpu_model(args, sd, model, noise)
    ref = CpuReference(args, sd, model, noise)
    n_coords = max(2, args.cpu_sample_coords)
    for _ in range(args.warmup):
        ref.run(lat, 2)
    t0 = time.perf_counter()
    evals = 0
    for _ in range(args.steps):
        _, _, n = ref.run(lat, n_coords)
        evals += n
    dt = time.perf_counter() - t0
    value = evals / dt
    sample = "per step: " + ref.describe(n_coords)
    emit({
        "impl": "reference", "metric": metric_name(args.image_size), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.image_size, ref.S, args.classifier), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": ref.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }, args)


def workload_name(size, S, kind):
    shape = {64: "CelebA-shaped", 256: "FFHQ-shaped"}.get(size, "")
    clf = "ResNet-18@224" if kind == "resnet" else f"MobileNetV2@{size}"
    return f"StylEx {size}px {shape} generator (capacity 16, S={S}) + {clf} 2-way classifier, random-init + calibrated; AttFind sweep"


# ------------------------------------------------------------------------------------------------------------------
def run_job(args, G, clf, clf_exact, noise, rank, world, dev, step_rate):
    """ONE whole AttFind job through the public API, timed by the wall clock (max over ranks): pinned host latents -> H2D
    -> `attfind_sweep(rank, world)` in the throughput mode (base images / logits / minima / maxima of ALL latents on every
    rank, this rank's contiguous shard swept) -> ONE all-gather of the effects -> `attfind_verify_topk` (the candidate
    columns re-evaluated in the parity mode, sharded the same way, so that the picks are the reference's exact top-k) ->
    class split + greedy top-k on device, merged ranking -> picks on the host.
    `job-latents` per rank: 128 x 8 ranks = BASELINE config 3; 256 x 1 at 64px = config 2."""
    import torch
    import torch.distributed as dist
    import stylex_b200 as sx
    from stylex_b200 import synthetic

    per_rank = args.job_latents or (128 if args.image_size >= 256 else 256)
    n_total = per_rank * world
    lat_pin = synthetic.make_latents(n_total, 4242).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    t0 = time.perf_counter()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    lat = lat_pin.to(dev, non_blocking=True)
    st = {}
    res = sx.attfind_sweep(G, clf, lat, noise, precision=args.precision, max_batch=args.max_batch, rank=rank,
                           world_size=world, gather=False, stats=st)
    ev[1].record()
    local = res["style_change"]
    if world > 1:
        from stylex_b200.dist import gather_effects
        effects = gather_effects(local, n_total, world)
    else:
        effects = local
    res["style_change"] = effects
    ev[2].record()
    fast = sx.attfind_select(effects, res["base_prob"], 5, 0.5)                    # what the throughput mode alone would pick
    ev[3].record()
    if args.no_verify:
        picks, merged, scores = fast
        vinfo = None
    else:
        picks, merged, scores, vinfo = sx.attfind_verify_topk(G, clf_exact, lat, noise, res, 5, 0.5, precision="fp32",
                                                              max_batch=128, rank=rank, world_size=world)
    ev[4].record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t = torch.tensor([wall] + [ev[i].elapsed_time(ev[i + 1]) for i in range(4)], device=dev, dtype=torch.float64)
    sig = torch.tensor([d * 100000 + s_ for c in (0, 1) for d, s_ in picks[c]] + [d * 100000 + s_ for d, s_ in merged] + [-1] * 10,
                       device=dev, dtype=torch.int64)[:20]
    agree = True
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sigs = [torch.empty_like(sig) for _ in range(world)]
        dist.all_gather(sigs, sig)
        agree = all(bool(torch.equal(sigs[0], x)) for x in sigs)
    wall = float(t[0].item())
    S = G.num_style_coords
    evals = n_total * 2 * S
    labels = torch.argmax(res["base_prob"], dim=1)
    out = {
        "what": "whole job, wall clock, max over ranks: H2D latents -> base images/logits + minima/maxima of all latents -> sharded "
                "bf16 sweep -> NCCL all-gather of effects -> top-k candidates re-evaluated in fp32 (sharded, small all-gathers) -> "
                "class split + greedy top-5 per class on device -> merged picks on host",
        "latents_total": n_total, "latents_per_rank": per_rank, "coord_evals": evals, "wall_s": wall,
        "value": evals / wall, "unit": UNIT, "ratio_to_step_rate": (evals / wall) / step_rate if step_rate else None,
        "sweep_ms": float(t[1].item()), "gather_ms": float(t[2].item()), "select_ms": float(t[3].item()),
        "verify_ms": float(t[4].item()),
        "effects_bytes_gathered": int(effects.numel() * 4) if world > 1 else 0,
        "class_sizes": [int((labels == 0).sum()), int((labels == 1).sum())],
        "picks": {str(c): [list(p) for p in picks[c]] for c in (0, 1)}, "merged": [list(p) for p in merged],
        "picks_agree_across_ranks": agree,
        "throughput_mode_picks": {str(c): [list(p) for p in fast[0][c]] for c in (0, 1)},
        "throughput_mode_picks_equal_exact": fast[0] == picks and fast[1] == merged,
    }
    if vinfo is not None:
        out["verify"] = {k: v for k, v in vinfo.items() if k not in ("style_change", "base_prob")}
        out["verify"]["exact_fraction_of_coord_evals"] = vinfo["exact_evals"] / evals
    return out


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
    torch.backends.cudnn.allow_tf32 = False          # fp32 work (calibration, the parity-mode verification) is real fp32;
    torch.backends.cuda.matmul.allow_tf32 = False    # the bf16 throughput classifier is not affected
    size, kind = args.image_size, args.classifier
    if args.max_batch is None:
        args.max_batch = sx.attfind.default_eval_batch(size)
    if args.classifier_dtype is None:
        args.classifier_dtype = "bf16" if kind == "resnet" else "fp32"
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
    # validated against the eager module on real generated images before it is trusted (classifiers.configure_throughput)
    info = clf.configure_throughput(calib[:8], dtype=torch.bfloat16 if args.classifier_dtype == "bf16" else torch.float32,
                                    fused=args.classifier_mode == "fused", stem=args.stem, maxpool=args.maxpool,
                                    preprocess=args.preprocess)
    clf_mode, pre_mode = info["classifier_mode"], info["preprocess"]
    clf_inner = clf

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

    # ---------------- whole job: host latents -> sweep -> all-gather -> selection -> picks (wall clock) ----------------
    job = None
    if not args.no_job:
        clf_exact = sx.make_classifier(kind, copy.deepcopy(model_cal_cpu).to(dev), size)      # parity mode: fp32 eager
        job = run_job(args, G, clf_inner, clf_exact, noise, rank, world, dev, value)

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
        ref = CpuReference(args, sd, model_cal_cpu, noise_cpu)
        ref.run(lat_all, 2)                                                    # warm-up (thread pools, oneDNN primitives)
        rate, dt, _ = ref.run(lat_all, max(2, args.cpu_sample_coords))
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": ref.kind,
               "sample": ref.describe(max(2, args.cpu_sample_coords)) + f", {dt:.1f} s"}

    out = {
        "metric": metric_name(size), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {"workload": workload_name(size, S, kind) + f", {lps} latent(s)/rank/step x {S} coords x 2 directions",
                   "coord_evals_per_step": int(2 * S * lps * world), "max_batch": args.max_batch,
                   "generator_dtype": args.precision, "classifier_dtype": info.get("dtype", args.classifier_dtype) + " (PyTorch, channels_last)", "classifier_mode": clf_mode, "classifier_preprocess": pre_mode,
                   "prefix_reuse": True, "l2": f"inputs larger than L2: every {args.max_batch}-eval batch streams >{args.max_batch * 16.8e6 * (size / 256) ** 2 / 1e9:.1f} GB of activations (L2 = 126 MB)",
                   "pool_latents": pool_n, "parallelism": f"latent-sharded x{world}, no data-path collective"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(tot[1].item()), "roofline": roofline, "cpu_baseline": cpu,
        "job": job, "conv_flops_per_full_image": conv_flops_per_image(plan.pairs),
    }
    emit(out, args)
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


def emit(obj, args=None):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT, line)
    if args is not None and getattr(args, "out", None):
        with open(args.out, "ab") as f:
            f.write(line)


if __name__ == "__main__":
    a = parse()
    _claim_stdout()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
