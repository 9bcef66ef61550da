"""Profiling driver (run under ncu on the GPU box): full 256px generator forwards through the plan, bf16.

    ncu --set full --clock-control none --import-source on -k regex:conv_tc -s <skip> -c <n> -o gpurun_out/prof \
        python profiles/prof_generator.py --batch 64 --iters 3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import stylex_b200 as sx
from stylex_b200 import synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
G = sx.Generator(a.size, 514).to(dev)
G.load_state_dict(synthetic.make_generator_state(a.size, seed=42), strict=False)
G.precision = a.precision
plan = G.plan()
lat = synthetic.make_latents(a.batch, 42).to(dev)
noise = synthetic.make_noise(a.size, 42).to(dev)
styles = plan.styles(sx.styles_def_to_tensor([(lat, G.num_layers)]).contiguous())
for _ in range(a.iters):
    img = plan.forward(styles, noise, precision=a.precision)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    img = plan.forward(styles, noise, precision=a.precision)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
print(f"generator {a.size}px {a.precision} batch {a.batch}: {ms:.3f} ms/forward = {ms / a.batch * 1e3:.1f} us/image, "
      f"{17.67e9 * a.batch / (ms * 1e-3) / 1e12:.1f} TFLOP/s (conv FLOPs only)")
