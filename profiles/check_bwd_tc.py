"""Bisect / time the tensor-core backward of one Conv2DMod shape (run each variant in its own process):
    python profiles/check_bwd_tc.py B Ci Co HW K        # env SX_BWD_NO_TC_DGRAD / SX_BWD_NO_TC_WGRAD select the kernels"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stylex_b200 import _native
from oracle import stylex_oracle as O

b, ci, co, hw, k = (int(v) for v in sys.argv[1:6])
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
x = torch.randn(b, ci, hw, hw, generator=g)
y = torch.randn(b, ci, generator=g) * 0.5
w = torch.randn(co, ci, k, k, generator=g) * (2.0 / (ci * k * k)) ** 0.5
go = torch.randn(b, co, hw, hw, generator=g)
lib = _native.lib()
xd, yd, wd, god = x.to(dev), y.to(dev), w.to(dev), go.to(dev)
ws = torch.empty(lib.sx_conv2dmod_workspace_bytes(b, ci, co, hw, hw, k, 0) + 256, dtype=torch.uint8, device=dev)
out = torch.empty(b, co, hw, hw, device=dev)
_native.check(lib.sx_conv2dmod_fwd(xd.data_ptr(), wd.data_ptr(), yd.data_ptr(), out.data_ptr(), b, ci, co, hw, hw, k, 1, 1e-8, 0,
                                   ws.data_ptr(), ws.numel(), _native.stream_ptr()), "fwd")
gx, gy, gw = torch.empty_like(xd), torch.empty_like(yd), torch.empty_like(wd)
wb = torch.empty(lib.sx_conv2dmod_bwd_workspace_bytes(b, ci, co, hw, hw, k, 1), dtype=torch.uint8, device=dev)


def run():
    _native.check(lib.sx_conv2dmod_bwd(xd.data_ptr(), wd.data_ptr(), yd.data_ptr(), out.data_ptr(), god.data_ptr(), gx.data_ptr(),
                                       gw.data_ptr(), gy.data_ptr(), b, ci, co, hw, hw, k, 1, 1e-8, 1, wb.data_ptr(), wb.numel(),
                                       _native.stream_ptr()), "bwd")


run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    run()
e1.record()
torch.cuda.synchronize()
if b * ci * hw * hw <= 1 << 22:
    _, gx_ref, gy_ref, gw_ref = O.modconv_grads(x, w, y, go, demod=True)
    rel = lambda a, r: float((a.double().cpu() - r).abs().max()) / max(1.0, float(r.abs().max()))
    errs = "rel err gx %.2e gy %.2e gw %.2e" % (rel(gx, gx_ref), rel(gy, gy_ref), rel(gw, gw_ref))
else:
    errs = "(too large for the CPU oracle)"
flops = 2 * 2.0 * k * k * ci * co * hw * hw * b
ms = e0.elapsed_time(e1) / 5
print("variant dgrad_tc=%d wgrad_tc=%d  B=%d %d->%d @%d k=%d: %.3f ms per backward (%.1f TFLOP/s incl. prep)  %s" % (
    "SX_BWD_NO_TC_DGRAD" not in os.environ, "SX_BWD_NO_TC_WGRAD" not in os.environ, b, ci, co, hw, k, ms, flops / ms / 1e9, errs), flush=True)
