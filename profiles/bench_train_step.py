#!/usr/bin/env python
"""BASELINE config 5: StylEx training step (generator + encoder + discriminator + classifier loss) at 1/2/4/8 B200.

    python profiles/bench_train_step.py --image-size 256 --batch 32 --steps 6 --warmup 2 [--precision bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        profiles/bench_train_step.py --gpus N ...

One step = ``training.TrainStep.train_step`` (discriminator phase + generator phase of ST:1249-1506, gradient accumulation 2
so that both the noise branch and the encoder branch -- reconstruction + classifier-KL losses -- run; gradient penalty every
4th step as in the reference).  DDP over S / G / D and the encoder (NCCL).  Weak scaling: ``--batch`` images per GPU and
accumulation pass.  Timed with CUDA events, max over ranks; one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import stylex_b200 as sx
from stylex_b200 import _native, dist as sxd, synthetic, training as T

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--image-size", type=int, default=256)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--accumulate", type=int, default=2)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
ap.add_argument("--out", default=None)
ap.add_argument("--channels-last", action="store_true", help="encoder / discriminator / classifier in torch.channels_last")
ap.add_argument("--profile", action="store_true", help="print the top CUDA kernels of one step (torch.profiler) instead of timing")
a = ap.parse_args()

rank, world, local = sxd.init_from_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
_native.device_check()
torch.backends.cudnn.benchmark = True
torch.manual_seed(1234 + rank)
size = a.image_size
st = sx.StylEx(size, rank=local)
st.G.precision = a.precision
model = synthetic.make_classifier_model("resnet", 42).to(dev)
clf = sx.make_classifier("resnet", model, size)
if a.channels_last:
    st.encoder.to(memory_format=torch.channels_last)
    st.D.to(memory_format=torch.channels_last)
    model.to(memory_format=torch.channels_last)
ts = T.TrainStep(st, clf, batch_size=a.batch, gradient_accumulate_every=a.accumulate, ddp=world > 1, rank=local)
pool = torch.rand(4 * a.batch, 3, size, size, device=dev)


def loader():
    i = 0
    while True:
        j = (i * a.batch) % (3 * a.batch)
        yield pool[j: j + a.batch].clone()
        i += 1


it = loader()
for _ in range(a.warmup):
    ts.train_step(it)
if a.profile:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ts.train_step(it)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
    sys.exit(0)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
launches0 = _native.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
logs = [ts.train_step(it) for _ in range(a.steps)]
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
t = torch.tensor([ms], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t.item())
if rank == 0:
    pairs = st.G.plan().pairs
    g_fwd = sum(2.0 * 9 * (ci * co + co * co) * (4 << l) ** 2 for l, (ci, co) in enumerate(pairs))      # Conv2DMod FLOPs / image
    # per accumulation pass: D phase 1 generator forward; G phase forward + backward (2x) = 4 forward-equivalents per image
    g_flops = 4 * g_fwd * a.batch * a.accumulate
    images = a.batch * a.accumulate * world * 2          # images through the generator per step (both phases), all ranks
    line = {
        "metric": "stylex_train_step_images_per_sec_%dpx" % size, "value": images / (ms * 1e-3), "unit": "generated images/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "scaling": "weak", "dtype": a.precision,
        "config": {"workload": "BASELINE config 5: StylEx %dpx training step (D phase + G phase, hinge + L1 reconstruction + classifier KL, "
                               "GP every 4th step), batch %d/GPU x accumulate %d, ResNet-18 classifier, DDP over S/G/D/encoder" % (size, a.batch, a.accumulate),
                   "generator_precision": a.precision,
                   "encoder_discriminator_classifier": "PyTorch / cuDNN fp32 (TF32 convolutions, torch default)" + (", channels_last" if a.channels_last else "")},
        "generator_conv_tflops_per_gpu": g_flops / (ms * 1e-3) / 1e12,
        "native_launches_per_step": (_native.launch_count() - launches0) / a.steps,
        "losses_last": {k: v for k, v in logs[-1].items()},
    }
    s = json.dumps(line)
    print(s, flush=True)
    if a.out:
        with open(a.out, "a") as f:
            f.write(s + "\n")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
