"""Time sx_stem_s2d_conv_relu alone (CUDA events, 256 x [115,115,16] bf16 = the AttFind batch): variants and the
SX_STEM_DEBUG knobs are read once per process, so each configuration is its own `python profiles/exp_stem.py` run."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stylex_b200 as sx  # noqa: E402
from stylex_b200.classifiers import FusedResNetInference  # noqa: E402

dev = torch.device("cuda:0")
b = int(os.environ.get("B", 256))
g = torch.Generator().manual_seed(0)
x = torch.randn(b, 16, 115, 115, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
wt = (torch.randn(64, 16, 4, 4, generator=g) * 0.1).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
bias = torch.randn(64, generator=g).to(dev).to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for pool in (True, False):
    f = FusedResNetInference.__new__(FusedResNetInference)
    f.dtype, f.stem_s2d, f.native_stem = torch.bfloat16, (wt, bias), None
    f.enable_native_stem(fuse_pool=pool)
    for _ in range(3):
        f._stem_native(x)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f._stem_native(x)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"variant={os.environ.get('SX_STEM_VARIANT', '0')} debug={os.environ.get('SX_STEM_DEBUG', '0')} pool={pool}: "
          f"median {ts[len(ts) // 2]:.4f} ms  min {ts[0]:.4f} ms", flush=True)
if os.environ.get("CUDNN"):
    for _ in range(3):
        y = torch.cudnn_convolution_relu(x, wt, bias, (1, 1), (0, 0), (1, 1), 1)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = torch.cudnn_convolution_relu(x, wt, bias, (1, 1), (0, 0), (1, 1), 1)
        z = torch.nn.functional.max_pool2d(y, 3, 2, 1)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"cudnn_convolution_relu + max_pool2d: median {ts[len(ts) // 2]:.4f} ms", flush=True)
