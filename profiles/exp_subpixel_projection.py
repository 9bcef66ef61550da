"""Measured proxy for the sub-pixel form of the fused-upsample layers: conv3x3(upsample2x(x), W) is a
plain 3x3 conv at the LOW resolution with 4*Co outputs, so its tensor-core cost is that of these three plain layers
(same FLOPs as c12 / c10 / c8 of the 256 px generator).  Run under ncu for the kernel times:

    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_tc --csv python profiles/exp_subpixel_projection.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import stylex_b200 as sx

torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
B = 256
for ci, co, h, name in ((64, 128, 128, "c12 as 64->128@128"), (128, 256, 64, "c10 as 128->256@64"), (256, 512, 32, "c8 as 256->512@32")):
    m = sx.Conv2DMod(ci, co, 3, precision="bf16").to(dev)
    x = torch.randn(B, ci, h, h, device=dev)
    y = torch.randn(B, ci, device=dev) * 0.3
    for _ in range(3):
        out = m(x, y)
    torch.cuda.synchronize()
    print(name, "flops/sample", 2 * 9 * ci * co * h * h, "checksum", float(out.abs().mean()))
