"""Per-layer timing of the 256px generator plan (bf16) at a sweep-sized batch: CUDA events around every launch
(sx_profile_enable), full forwards.  Used for A/B runs of kernel variants (SX_HALO_VARIANT, SX_HALO_DEBUG, ...):

    python profiles/exp_layers.py --batch 128 --iters 5
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import stylex_b200 as sx
from stylex_b200 import _native, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--tag", default="")
a = ap.parse_args()
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
G = sx.Generator(a.size, 514).to(dev)
G.load_state_dict(synthetic.make_generator_state(a.size, seed=42), strict=False)
G.precision = "bf16"
plan = G.plan()
lat = synthetic.make_latents(a.batch, 42).to(dev)
noise = synthetic.make_noise(a.size, 42).to(dev)
styles = plan.styles(sx.styles_def_to_tensor([(lat, G.num_layers)]).contiguous())
for _ in range(2):
    img = plan.forward(styles, noise, precision="bf16")
torch.cuda.synchronize()
_native.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    img = plan.forward(styles, noise, precision="bf16")
e1.record()
torch.cuda.synchronize()
prof = _native.profile_collect()
_native.profile_enable(False)
ms = e0.elapsed_time(e1) / a.iters
names = {32: "modulate", 33: "upsample", 34: "torgb", 35: "demod", 37: "rgb_prev"}
parts = []
for k, v in sorted(prof.items()):
    t = v["ms"] / a.iters
    tf = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 and v["flops"] > 0 else 0
    parts.append(f"{names.get(k, 'c%d' % k)}={t:.3f}ms" + (f"/{tf:.0f}TF" if tf else ""))
print(f"[{a.tag}] {a.size}px b{a.batch}: {ms:.3f} ms/fwd checksum={float(img.float().abs().mean()):.6f} | " + " ".join(parts), flush=True)
