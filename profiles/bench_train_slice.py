"""Training-step slice: Generator forward + backward of an image loss, native kernels
(fp32: FFMA implicit GEMMs for conv / dgrad / wgrad, gather adjoints for the bandwidth ops), CUDA-event timed, one JSON
line per configuration.  The CPU line is the oracle port (torch autograd through the literal restatement, fp32) on the
host cores, bounded to one small batch.

    python profiles/bench_train_slice.py [--size 64|256] [--batch B] [--steps K] [--warmup W] [--no-cpu]

FLOPs: forward 2*k^2*Ci*Co*H*W per conv call per sample; backward = dgrad + wgrad = 2x that.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import stylex_b200 as sx
from stylex_b200 import _native, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=64)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"],
                help="bf16: the 3x3 modulated convs run forward / dgrad / wgrad on the tcgen05 kernels where they take the shape")
a = ap.parse_args()

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
sd = synthetic.make_generator_state(a.size, seed=42)
G = sx.Generator(a.size, 514).to(dev)
G.load_state_dict(sd, strict=False)
G.train()
G.precision = a.precision
pairs = synthetic.generator_pairs(a.size)
fwd_flops = 0.0
for l, (ci, co) in enumerate(pairs):
    hw = (4 << l) ** 2
    fwd_flops += 2.0 * hw * (9 * ci * co + 9 * co * co + 3 * co)
lat = synthetic.make_latents(a.batch, 1).to(dev)
noise = synthetic.make_noise(a.size, 42).to(dev)
styles = sx.styles_def_to_tensor([(lat, G.num_layers)])
go = torch.randn(a.batch, 3, a.size, a.size, device=dev)


def step():
    for p in G.parameters():
        p.grad = None
    rgb = G(styles, noise)
    (rgb * go).sum().backward()
    return rgb


for _ in range(a.warmup):
    step()
torch.cuda.synchronize()
l0 = _native.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
line = {"metric": "generator_fwd_bwd_images_per_sec", "value": a.batch / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms,
        "config": {"workload": f"StylEx {a.size}px generator forward + backward of <rgb, g>, batch {a.batch}, {a.precision} native kernels"},
        "dtype": "f32" if a.precision == "fp32" else "bf16", "gpu_launches": (_native.launch_count() - l0) // a.steps,
        "tflops": 3.0 * fwd_flops * a.batch / (ms * 1e-3) / 1e12, "mem_gb": torch.cuda.max_memory_allocated() / 1e9}
if not a.no_cpu:
    from oracle import stylex_oracle as O
    nb = min(a.batch, 2)
    t0 = time.perf_counter()
    O.generator_grads(sd, styles[:nb].cpu(), noise.cpu(), go[:nb].cpu(), dtype=torch.float32)
    dt = time.perf_counter() - t0
    line["cpu_baseline"] = {"value": nb / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                            "sample": f"{nb} images, torch autograd through the oracle restatement, fp32, {dt:.1f} s"}
print(json.dumps(line))
