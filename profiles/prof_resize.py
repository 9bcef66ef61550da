"""ncu target: the classifier preprocessing kernels at the sweep's batch size (separable and direct s2d variants).
    ncu --set full --import-source on --clock-control none -k regex:resize_aa -o gpurun_out/resize python profiles/prof_resize.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import stylex_b200 as sx

dev = torch.device("cuda:0")
model = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(3 * 224 * 224, 2)).to(dev)
clf = sx.make_classifier("resnet", model, 256)
clf.to(dev).set_compute(torch.bfloat16, channels_last=True)
x = torch.rand(256, 3, 256, 256, device=dev)
for _ in range(2):
    a = clf._native_pre(x, s2d=True)
os.environ["SX_RESIZE_DIRECT"] = "1"
b = clf._native_pre(x, s2d=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, env in (("direct", "1"), ("separable", None)):
    if env:
        os.environ["SX_RESIZE_DIRECT"] = env
    else:
        os.environ.pop("SX_RESIZE_DIRECT", None)
    e0.record()
    for _ in range(10):
        clf._native_pre(x, s2d=True)
    e1.record()
    torch.cuda.synchronize()
    print(name, "%.1f us per launch (batch 256, 256 -> 224)" % (e0.elapsed_time(e1) * 100))
print("equal", torch.equal(a, b))
