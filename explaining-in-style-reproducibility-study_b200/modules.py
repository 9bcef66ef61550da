"""Drop-in replacements for the reference's generator-side ``nn.Module`` classes.

Same constructor arguments, attributes, forward signatures and **state-dict keys** as
``stylex/stylex_train.py`` of the reference (``Blur`` :144-153, ``RGBBlock`` :604-629, ``Conv2DMod``
:632-667, ``GeneratorBlock`` :670-718, ``Generator`` :747-840), so a reference checkpoint's
``["StylEx"]["G.*"]`` entries load with ``load_state_dict``.  Every forward runs hand-written sm_100a
kernels through the C ABI (``include/stylex_b200.h``); there is no PyTorch / CPU fallback.

Two levels:

* module level (``Conv2DMod.forward(x, y)``, ``GeneratorBlock.forward(x, prev_rgb, istyle, inoise)``,
  ``RGBBlock.forward``, ``Blur.forward``): tensors cross the boundary in the reference's NCHW fp32
  layout, one native op per reference op -- this is the compatibility surface.
* plan level (``Generator.forward`` and the AttFind sweep): the whole synthesis network runs inside
  the library on NHWC activations with fused epilogues (``GeneratorPlan``) -- this is the fast path.

Autograd (first slice): at module level every op is a ``torch.autograd.Function`` whose
forward AND first-order backward are native kernels (``csrc/conv_bwd.cuh``, ``csrc/bwd_ops.cuh``), fp32, deterministic.
``Generator.forward`` takes that path when gradients are being recorded; under ``torch.no_grad()`` (AttFind, rendering)
it runs the fused plan, which records nothing.  With ``precision = "bf16"`` the 3x3 modulated convolutions run forward,
dgrad and wgrad on the tcgen05 kernels (bf16 operands, fp32 accumulation).

Double backward (the path-length penalty ST:306-316 differentiates d(image)/d(styles) again; the gradient penalty ST:296-303
does the same through the discriminator): ``Blur`` and the bilinear upsample are linear maps, so each is a pair of Functions
(operator / adjoint) that are each other's backward -- differentiable to any order.  With ``Generator.double_backward = True``
the generator runs ``_forward_dd``: Conv2DMod as ``d * conv(W, x * (s + 1))`` where the convolution with batch-shared weights
is ``ConvSharedFunction`` (native FFMA kernel) whose backward is again a ``ConvSharedFunction`` (dgrad = conv with the flipped,
transposed weights) plus ``ConvWgradFunction`` (native wgrad), whose own backward is two ``ConvSharedFunction``s: the pair is
closed under differentiation.  The cheap glue (modulation, demodulation coefficients, noise, leaky-ReLU, the style affines)
is plain torch in that mode.  The default training path stays the fused first-order Functions above.
"""
from __future__ import annotations

import ctypes
from functools import partial
from math import log2
from typing import Optional

import torch
from torch import nn

from . import _native as N


def exists(val):
    return val is not None


def leaky_relu(p=0.2):
    return nn.LeakyReLU(p, inplace=True)


def image_noise(n, im_size, device):
    """reference stylex_train.py:336-337 (U[0,1) drawn on the CPU, then moved)."""
    return torch.FloatTensor(n, im_size, im_size, 1).uniform_(0., 1.).to(_dev(device))


def styles_def_to_tensor(styles_def):
    """reference stylex_train.py:352-353."""
    return torch.cat([t[:, None, :].expand(-1, n, -1) for t, n in styles_def], dim=1)


def _dev(device):
    return torch.device("cuda", device) if isinstance(device, int) else torch.device(device)


def _prec(module) -> int:
    return N.PRECISIONS[getattr(module, "precision", "fp32")]


_op_ws = N.Workspace()


def _noise_arg(inoise: torch.Tensor):
    """[B|1, S, S, 1] (reference layout) -> dense fp32 [nb, S, S]."""
    if inoise.dim() != 4 or inoise.shape[-1] != 1 or inoise.shape[1] != inoise.shape[2]:
        raise ValueError(f"input_noise must be [B|1, S, S, 1], got {tuple(inoise.shape)}")
    return N.f32c(inoise), inoise.shape[0], inoise.shape[1]


# ---------------------------------------------------------------------------------------------
# L1 ops
# ---------------------------------------------------------------------------------------------
def _wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _blur_fwd(x):
    x = N.f32c(x)
    b, c, h, w = x.shape
    out = torch.empty_like(x)
    N.check(N.lib().sx_blur3x3_reflect(x.data_ptr(), out.data_ptr(), b, c, h, w, N.stream_ptr()), "sx_blur3x3_reflect")
    return out


def _blur_bwd(g):
    g = N.f32c(g)
    b, c, h, w = g.shape
    gx = torch.empty_like(g)
    N.check(N.lib().sx_blur3x3_reflect_bwd(g.data_ptr(), gx.data_ptr(), b, c, h, w, N.stream_ptr()), "sx_blur3x3_reflect_bwd")
    return gx


def _up_fwd(x):
    x = N.f32c(x)
    b, c, h, w = x.shape
    out = torch.empty(b, c, 2 * h, 2 * w, device=x.device, dtype=torch.float32)
    N.check(N.lib().sx_upsample2x_bilinear(x.data_ptr(), out.data_ptr(), b, c, h, w, N.stream_ptr()), "sx_upsample2x_bilinear")
    return out


def _up_bwd(g):
    g = N.f32c(g)
    b, c, h2, w2 = g.shape
    gx = torch.empty(b, c, h2 // 2, w2 // 2, device=g.device, dtype=torch.float32)
    N.check(N.lib().sx_upsample2x_bilinear_bwd(g.data_ptr(), gx.data_ptr(), b, c, h2 // 2, w2 // 2, N.stream_ptr()),
            "sx_upsample2x_bilinear_bwd")
    return gx


class BlurFunction(torch.autograd.Function):
    """Blur.forward ST:144-153.  A linear map: its backward is the adjoint (reflected border taps fold back inside), itself a
    Function whose backward is this one -- differentiable to any order (the gradient penalty ST:296-303 needs the second)."""

    @staticmethod
    def forward(ctx, x):
        return _blur_fwd(x.detach())

    @staticmethod
    def backward(ctx, g):
        return BlurAdjointFunction.apply(g)


class BlurAdjointFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g):
        return _blur_bwd(g.detach())

    @staticmethod
    def backward(ctx, gg):
        return BlurFunction.apply(gg)


class Upsample2xFunction(torch.autograd.Function):
    """nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False) ST:679; backward = the adjoint, whose backward is
    the upsample again (linear map: differentiable to any order)."""

    @staticmethod
    def forward(ctx, x):
        return _up_fwd(x.detach())

    @staticmethod
    def backward(ctx, g):
        return Upsample2xAdjointFunction.apply(g)


class Upsample2xAdjointFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g):
        return _up_bwd(g.detach())

    @staticmethod
    def backward(ctx, gg):
        return Upsample2xFunction.apply(gg)


class LinearFunction(torch.autograd.Function):
    """nn.Linear (to_style1/2, RGBBlock.to_style) forward / backward on the native kernels."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        xd, wd = N.f32c(x.detach()), N.f32c(weight.detach())
        bd = None if bias is None else N.f32c(bias.detach())
        out = torch.empty(xd.shape[0], wd.shape[0], device=xd.device, dtype=torch.float32)
        N.check(N.lib().sx_linear_fwd(xd.data_ptr(), wd.data_ptr(), N.ptr(bd), out.data_ptr(), xd.shape[0], xd.shape[1],
                                      wd.shape[0], N.stream_ptr()), "sx_linear_fwd")
        ctx.save_for_backward(xd, wd)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g = N.f32c(g)
        gx, gw = torch.empty_like(x), torch.empty_like(w)
        gb = torch.empty(w.shape[0], device=w.device, dtype=torch.float32) if ctx.has_bias else None
        N.check(N.lib().sx_linear_bwd(x.data_ptr(), w.data_ptr(), g.data_ptr(), gx.data_ptr(), gw.data_ptr(), N.ptr(gb),
                                      x.shape[0], x.shape[1], w.shape[0], N.stream_ptr()), "sx_linear_bwd")
        return gx, gw, gb


class NoiseLReLUFunction(torch.autograd.Function):
    """leaky_relu_0.2(x + to_noise(inoise[:, :H, :W]).permute(0,3,2,1)) ST:696-698,705,714 and its backward: grad_x and
    the gradients of the to_noise Linear(1, C) (the noise map itself is data)."""

    @staticmethod
    def forward(ctx, x, inoise, nw, nbias):
        xd = N.f32c(x.detach())
        nz, nb, ns = _noise_arg(inoise)
        b, c, h, w = xd.shape
        out = torch.empty_like(xd)
        nwd, nbd = N.f32c(nw.detach()), N.f32c(nbias.detach())
        N.check(N.lib().sx_noise_lrelu(xd.data_ptr(), nz.data_ptr(), nwd.data_ptr(), nbd.data_ptr(), out.data_ptr(), b, c, h, w,
                                       nb, ns, N.stream_ptr()), "sx_noise_lrelu")
        ctx.save_for_backward(out, nz)
        ctx.noise = (nb, ns)
        ctx.w_shape = tuple(nw.shape)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        out, nz = ctx.saved_tensors
        nb, ns = ctx.noise
        g = N.f32c(g)
        b, c, h, w = out.shape
        lib = N.lib()
        gx = torch.empty_like(out)
        gnw = torch.empty(c, device=out.device, dtype=torch.float32)
        gnb = torch.empty_like(gnw)
        ws = _op_ws.get(lib.sx_noise_lrelu_bwd_workspace_bytes(b, c), out.device)
        N.check(lib.sx_noise_lrelu_bwd(out.data_ptr(), g.data_ptr(), nz.data_ptr(), gx.data_ptr(), gnw.data_ptr(), gnb.data_ptr(),
                                       b, c, h, w, nb, ns, ws.data_ptr(), ws.numel(), N.stream_ptr()), "sx_noise_lrelu_bwd")
        return gx, None, gnw.reshape(ctx.w_shape), gnb


class RGBTailFunction(torch.autograd.Function):
    """RGBBlock tail ST:623-627: out = blur(upsample2x(rgb + prev)) (or rgb + prev for the last block), one fused native
    kernel; backward = blur adjoint -> upsample adjoint, and the same gradient flows to both summands."""

    @staticmethod
    def forward(ctx, rgb, prev, up):
        x = N.f32c(rgb.detach())
        pv = None if prev is None else N.f32c(prev.detach())
        b, c, h, w = x.shape
        out = torch.empty(b, c, h * (2 if up else 1), w * (2 if up else 1), device=x.device, dtype=torch.float32)
        N.check(N.lib().sx_rgb_add_upsample_blur(x.data_ptr(), N.ptr(pv), out.data_ptr(), b, c, h, w, 1 if up else 0,
                                                 N.stream_ptr()), "sx_rgb_add_upsample_blur")
        ctx.up, ctx.has_prev = up, prev is not None
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        gs = _up_bwd(_blur_bwd(g)) if ctx.up else N.f32c(g)
        return gs, (gs if ctx.has_prev else None), None


class Blur(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer('f', torch.Tensor([1, 2, 1]))

    def forward(self, x):
        N.require_cuda(x)
        N.device_check()
        return BlurFunction.apply(x)


def upsample2x(x: torch.Tensor) -> torch.Tensor:
    """nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False) on the native kernel."""
    N.require_cuda(x)
    N.device_check()
    return Upsample2xFunction.apply(x)


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    N.require_cuda(x, weight, bias)
    N.device_check()
    return LinearFunction.apply(x, weight, bias)


def noise_lrelu(x: torch.Tensor, inoise: torch.Tensor, to_noise: nn.Linear) -> torch.Tensor:
    """leaky_relu_0.2(x + to_noise(inoise[:, :H, :W]).permute(0,3,2,1)) -- reference :696-698,705."""
    N.require_cuda(x, inoise)
    N.device_check()
    return NoiseLReLUFunction.apply(x, inoise, to_noise.weight, to_noise.bias)


class Conv2DMod(nn.Module):
    def __init__(self, in_chan, out_chan, kernel, demod=True, stride=1, dilation=1, eps=1e-8, **kwargs):
        super().__init__()
        self.filters = out_chan
        self.demod = demod
        self.kernel = kernel
        self.stride = stride
        self.dilation = dilation
        self.weight = nn.Parameter(torch.randn((out_chan, in_chan, kernel, kernel)))
        self.eps = eps
        self.precision = kwargs.get("precision", "fp32")
        nn.init.kaiming_normal_(self.weight, a=0, mode='fan_in', nonlinearity='leaky_relu')

    def _get_same_padding(self, size, kernel, dilation, stride):
        return ((size - 1) * (stride - 1) + dilation * (kernel - 1)) // 2

    def forward(self, x, y):
        if self.stride != 1 or self.dilation != 1:
            raise NotImplementedError("stylex_b200 Conv2DMod: only stride=1, dilation=1 (all the reference ever uses)")
        N.require_cuda(x, y, self.weight)
        N.device_check()
        b, c = x.shape[:2]
        co, ci, k, _ = self.weight.shape
        if c != ci or y.shape != (b, ci):
            raise ValueError(f"Conv2DMod: x {tuple(x.shape)} / style {tuple(y.shape)} do not match weight {tuple(self.weight.shape)}")
        if _wants_grad(x, y, self.weight):
            # training: the native forward + the native first-order backward (sx_conv2dmod_bwd).  precision "bf16" runs the
            # forward, the dgrad and the wgrad on the tensor cores where the kernels take the shape (3x3, tensor-core channel
            # counts, power-of-two maps) and this conv in fp32 otherwise (the 1x1 ToRGB convs)
            prec = _prec(self)
            if prec == N.PREC_BF16 and not _tc_shape_ok(ci, co, x.shape[2], x.shape[3], k):
                prec = N.PREC_FP32
            return Conv2DModFunction.apply(x, y, self.weight, bool(self.demod), float(self.eps), prec)
        return _conv2dmod_forward(x, y, self.weight.detach(), bool(self.demod), float(self.eps), _prec(self))


def _tc_shape_ok(ci, co, h, w, k) -> bool:
    """mirror of tc::tc_shape_supported (csrc/conv_tc.cuh): shapes the bf16 tcgen05 conv kernels take."""
    return (k == 3 and ci % 32 == 0 and (co in (32, 64, 128) or co % 256 == 0) and h == w and w >= 4 and (w & (w - 1)) == 0)


def _conv2dmod_forward(x, y, w, demod, eps, prec):
    x, y, w = N.f32c(x), N.f32c(y), N.f32c(w)
    b, ci, h, wd = x.shape
    co, _, k, _ = w.shape
    lib = N.lib()
    nbytes = lib.sx_conv2dmod_workspace_bytes(b, ci, co, h, wd, k, prec)
    ws = _op_ws.get(nbytes, x.device)
    out = torch.empty(b, co, h, wd, device=x.device, dtype=torch.float32)
    N.check(lib.sx_conv2dmod_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), out.data_ptr(), b, ci, co, h, wd, k,
                                 1 if demod else 0, eps, prec, ws.data_ptr(), ws.numel(), N.stream_ptr()), "sx_conv2dmod_fwd")
    return out


class Conv2DModFunction(torch.autograd.Function):
    """Conv2DMod.forward (ST:647-667) with its first-order backward on the native kernels (``csrc/conv_bwd.cuh``; with
    ``prec`` = bf16 the forward, the dgrad and the wgrad run on the tcgen05 kernels, ``csrc/wgrad_tc.cuh``):
    grad_x (dgrad through the shared weights), grad_weight (wgrad over all pixels of the batch + the demodulation term)
    and grad_style (modulation + demodulation terms).  Not twice differentiable: the path-length and gradient penalties
    of the training step (ST:296-316) need a double backward that is not built."""

    @staticmethod
    def forward(ctx, x, y, weight, demod, eps, prec=N.PREC_FP32):
        xd, yd, wdt = N.f32c(x.detach()), N.f32c(y.detach()), N.f32c(weight.detach())
        out = _conv2dmod_forward(xd, yd, wdt, demod, eps, prec)
        ctx.save_for_backward(xd, yd, wdt, out)
        ctx.demod, ctx.eps, ctx.prec = demod, eps, prec
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        x, y, w, out = ctx.saved_tensors
        N.require_cuda(grad_out)
        g = N.f32c(grad_out)
        b, ci, h, wd = x.shape
        co, _, k, _ = w.shape
        lib = N.lib()
        gx = torch.empty_like(x)
        gy = torch.empty_like(y)
        gw = torch.empty_like(w)
        ws = _op_ws.get(lib.sx_conv2dmod_bwd_workspace_bytes(b, ci, co, h, wd, k, ctx.prec), x.device)
        N.check(lib.sx_conv2dmod_bwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), out.data_ptr(), g.data_ptr(), gx.data_ptr(),
                                     gw.data_ptr(), gy.data_ptr(), b, ci, co, h, wd, k, 1 if ctx.demod else 0, ctx.eps,
                                     ctx.prec, ws.data_ptr(), ws.numel(), N.stream_ptr()), "sx_conv2dmod_bwd")
        return gx, gy, gw, None, None, None


class ConvSharedFunction(torch.autograd.Function):
    """y = conv2d(x, W, padding=(k-1)//2) with weights shared by the whole batch (what Conv2DMod becomes once the modulation
    sits on the activations), NCHW fp32, native FFMA kernel.  Closed under differentiation together with
    ``ConvWgradFunction``: grad_x = conv(g, flip(W)^T) is another ConvSharedFunction, grad_W = wgrad(x, g)."""

    @staticmethod
    def forward(ctx, x, w):
        N.require_cuda(x, w)
        xd, wd = N.f32c(x.detach()), N.f32c(w.detach())
        ctx.save_for_backward(x, w)
        zero_style = torch.zeros(xd.shape[0], xd.shape[1], device=xd.device, dtype=torch.float32)
        return _conv2dmod_forward(xd, zero_style, wd, False, 1e-8, N.PREC_FP32)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = ConvSharedFunction.apply(g, w.flip(2, 3).transpose(0, 1))
        if ctx.needs_input_grad[1]:
            gw = ConvWgradFunction.apply(x, g, w.shape[2])
        return gx, gw


class ConvWgradFunction(torch.autograd.Function):
    """gW[o,i,ky,kx] = sum_{b,y,x} g[b,o,y,x] * x[b,i,y+ky-p,x+kx-p] (native wgrad kernels).  Bilinear in (x, g): its backward
    w.r.t. x is conv(g, flip(gW_bar)^T), w.r.t. g is conv(x, gW_bar) -- both ConvSharedFunctions."""

    @staticmethod
    def forward(ctx, x, g, k):
        N.require_cuda(x, g)
        xd, gd = N.f32c(x.detach()), N.f32c(g.detach())
        b, ci, h, wd = xd.shape
        co = gd.shape[1]
        ctx.save_for_backward(x, g)
        lib = N.lib()
        zero_style = torch.zeros(b, ci, device=xd.device, dtype=torch.float32)
        w_dummy = torch.zeros(co, ci, k, k, device=xd.device, dtype=torch.float32)   # feeds the (discarded) dgrad
        gx = torch.empty_like(xd)
        gy = torch.empty_like(zero_style)
        gw = torch.empty_like(w_dummy)
        ws = _op_ws.get(lib.sx_conv2dmod_bwd_workspace_bytes(b, ci, co, h, wd, k, N.PREC_FP32), xd.device)
        N.check(lib.sx_conv2dmod_bwd(xd.data_ptr(), w_dummy.data_ptr(), zero_style.data_ptr(), None, gd.data_ptr(), gx.data_ptr(),
                                     gw.data_ptr(), gy.data_ptr(), b, ci, co, h, wd, k, 0, 1e-8, N.PREC_FP32, ws.data_ptr(),
                                     ws.numel(), N.stream_ptr()), "sx_conv2dmod_bwd")
        return gw

    @staticmethod
    def backward(ctx, gw_bar):
        x, g = ctx.saved_tensors
        gx = gg = None
        if ctx.needs_input_grad[0]:
            gx = ConvSharedFunction.apply(g, gw_bar.flip(2, 3).transpose(0, 1))
        if ctx.needs_input_grad[1]:
            gg = ConvSharedFunction.apply(x, gw_bar)
        return gx, gg, None


def conv2dmod_dd(x, y, weight, demod=True, eps=1e-8):
    """Conv2DMod.forward ST:647-667 in its twice-differentiable form: out = d[b,o] * conv(W, x * (y + 1)),
    d = rsqrt(((y + 1)^2) @ (sum_taps W^2)^T + eps)  (the identity the fused kernels use)."""
    s = y + 1
    out = ConvSharedFunction.apply(x * s[:, :, None, None], weight)
    if demod:
        d = torch.rsqrt((s * s) @ (weight * weight).sum(dim=(2, 3)).t() + eps)
        out = out * d[:, :, None, None]
    return out


class RGBBlock(nn.Module):
    def __init__(self, latent_dim, input_channel, upsample, rgba=False):
        super().__init__()
        self.input_channel = input_channel
        self.to_style = nn.Linear(latent_dim, input_channel)

        out_filters = 3 if not rgba else 4
        self.conv = Conv2DMod(input_channel, out_filters, 1, demod=False)

        # kept as modules so the state-dict key `upsample.1.f` (the Blur buffer) exists like in the reference
        self.upsample = nn.Sequential(
            nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False),
            Blur()
        ) if upsample else None

    def forward(self, x, prev_rgb, istyle):
        if getattr(self, "double_backward", False):
            return self._forward_dd(x, prev_rgb, istyle)
        style = linear(istyle, self.to_style.weight, self.to_style.bias)
        x = self.conv(x, style)
        return RGBTailFunction.apply(x, prev_rgb, exists(self.upsample))

    def _forward_dd(self, x, prev_rgb, istyle):
        """ST:618-629 out of twice-differentiable pieces."""
        style = nn.functional.linear(istyle, self.to_style.weight, self.to_style.bias)
        x = conv2dmod_dd(x, style, self.conv.weight, demod=False)
        if exists(prev_rgb):
            x = x + prev_rgb
        if exists(self.upsample):
            x = BlurFunction.apply(Upsample2xFunction.apply(x))
        return x


class GeneratorBlock(nn.Module):
    def __init__(self, latent_dim, input_channels, filters, upsample=True, upsample_rgb=True, rgba=False):
        super().__init__()

        self.input_channels = input_channels
        self.filters = filters

        self.num_style_coords = self.input_channels + self.filters

        self.upsample = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False) if upsample else None

        self.to_style1 = nn.Linear(latent_dim, input_channels)
        self.to_noise1 = nn.Linear(1, filters)
        self.conv1 = Conv2DMod(input_channels, filters, 3)

        self.to_style2 = nn.Linear(latent_dim, filters)
        self.to_noise2 = nn.Linear(1, filters)
        self.conv2 = Conv2DMod(filters, filters, 3)

        self.activation = leaky_relu()
        self.to_rgb = RGBBlock(latent_dim, filters, upsample_rgb, rgba)

    def forward(self, x, prev_rgb, istyle, inoise):
        if getattr(self, "double_backward", False):
            return self._forward_dd(x, prev_rgb, istyle, inoise)
        if exists(self.upsample):
            x = upsample2x(x)
        style1 = linear(istyle, self.to_style1.weight, self.to_style1.bias)
        x = self.conv1(x, style1)
        x = noise_lrelu(x, inoise, self.to_noise1)
        style2 = linear(istyle, self.to_style2.weight, self.to_style2.bias)
        style_coords = torch.cat([style1, style2], dim=-1)
        x = self.conv2(x, style2)
        x = noise_lrelu(x, inoise, self.to_noise2)
        rgb = self.to_rgb(x, prev_rgb, istyle)
        return x, rgb, style_coords

    def _forward_dd(self, x, prev_rgb, istyle, inoise):
        """ST:692-718 out of twice-differentiable pieces (native convolutions / upsample / blur, torch glue)."""
        N.require_cuda(x, istyle, inoise)
        lin = nn.functional.linear
        if exists(self.upsample):
            x = Upsample2xFunction.apply(x)
        inoise = inoise[:, :x.shape[2], :x.shape[3], :]
        noise1 = self.to_noise1(inoise).permute((0, 3, 2, 1))
        noise2 = self.to_noise2(inoise).permute((0, 3, 2, 1))
        style1 = lin(istyle, self.to_style1.weight, self.to_style1.bias)
        x = conv2dmod_dd(x, style1, self.conv1.weight, True, self.conv1.eps)
        x = nn.functional.leaky_relu(x + noise1, 0.2)
        style2 = lin(istyle, self.to_style2.weight, self.to_style2.bias)
        style_coords = torch.cat([style1, style2], dim=-1)
        x = conv2dmod_dd(x, style2, self.conv2.weight, True, self.conv2.eps)
        x = nn.functional.leaky_relu(x + noise2, 0.2)
        self.to_rgb.double_backward = True
        rgb = self.to_rgb(x, prev_rgb, istyle)
        return x, rgb, style_coords


# ---------------------------------------------------------------------------------------------
# plan level
# ---------------------------------------------------------------------------------------------
class GeneratorPlan:
    """Owns the native ``sx_generator_t`` of one ``Generator`` on one device (packed weights, workspaces)."""

    def __init__(self, generator: "Generator"):
        self.G = generator
        self.handle = ctypes.c_void_p()
        self.device = None
        self._versions = None
        self._ws = {}            # precision -> Workspace
        self._cache_ok = {}      # precision -> bool (clean-prefix cache primed on the current workspace)
        self._scratch = {}       # name -> tensor (scratch())
        pairs = [(b.input_channels, b.filters) for b in generator.blocks]
        self.pairs = pairs
        ci = (ctypes.c_int * len(pairs))(*[p[0] for p in pairs])
        co = (ctypes.c_int * len(pairs))(*[p[1] for p in pairs])
        N.check(N.lib().sx_generator_create(ci, co, len(pairs), generator.latent_dim, ctypes.byref(self.handle)),
                "sx_generator_create")
        self.S = N.lib().sx_generator_num_style_coords(self.handle)
        self.row = N.lib().sx_generator_style_row(self.handle)
        # style-coordinate offset of every conv (2*block + {0,1}) and its width
        self.conv_coords = []
        off = 0
        for cin, cout in pairs:
            self.conv_coords.append((off, cin))
            off += cin
            self.conv_coords.append((off, cout))
            off += cout

    def __del__(self):
        try:
            if self.handle:
                N.lib().sx_generator_destroy(self.handle)
        except Exception:
            pass

    def _param_list(self):
        G = self.G
        ps = [G.initial_block, G.initial_conv.weight, G.initial_conv.bias]
        for b in G.blocks:
            ps += [b.to_style1.weight, b.to_style1.bias, b.to_noise1.weight, b.to_noise1.bias, b.conv1.weight,
                   b.to_style2.weight, b.to_style2.bias, b.to_noise2.weight, b.to_noise2.bias, b.conv2.weight,
                   b.to_rgb.to_style.weight, b.to_rgb.to_style.bias, b.to_rgb.conv.weight]
        return ps

    def sync(self):
        """(re)pack the weights when the module's parameters changed (load_state_dict, .to(), optimiser step)."""
        ps = self._param_list()
        N.require_cuda(*ps, same_device=False)      # packing runs under the parameters' own device (below)
        versions = tuple((p.data_ptr(), p._version) for p in ps)
        if versions == self._versions:
            return
        dev = ps[0].device
        with torch.cuda.device(dev):
            N.device_check()
        keep = [N.f32c(p.detach()) for p in ps]
        blocks = (N.sx_block_params * len(self.G.blocks))()
        names = [f[0] for f in N.sx_block_params._fields_]
        for i in range(len(self.G.blocks)):
            for j, name in enumerate(names):
                setattr(blocks[i], name, keep[3 + 13 * i + j].data_ptr())
        with torch.cuda.device(dev):
            N.check(N.lib().sx_generator_load(self.handle, keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr(),
                                              blocks, N.stream_ptr()), "sx_generator_load")
        self.device = dev
        self._versions = versions
        self._cache_ok = {}

    def reserve(self, max_batch: int, precision) -> None:
        """size the workspace for ``max_batch`` up front (growing it later would drop the clean-prefix cache)."""
        prec = N.PRECISIONS[precision]
        self.sync()
        need = N.lib().sx_generator_workspace_bytes(self.handle, max_batch, prec)
        ws = self._ws.setdefault(prec, N.Workspace())
        old = None if ws.buf is None else ws.buf.data_ptr()
        buf = ws.get(need, self.device)
        if buf.data_ptr() != old:
            self._cache_ok[prec] = False

    def scratch(self, name: str, shape, device) -> torch.Tensor:
        """a named fp32 scratch tensor owned by the plan, re-used across calls while shape and device stay the same"""
        t = self._scratch.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.device != torch.device(device):
            t = torch.empty(tuple(shape), device=device, dtype=torch.float32)
            self._scratch[name] = t
        return t

    def styles(self, w: torch.Tensor) -> torch.Tensor:
        """w [B, L, latent] -> styles [B, row] = [style coords (S) | ToRGB styles]."""
        self.sync()
        N.require_cuda(w)
        w = N.f32c(w)
        if w.dim() != 3 or w.shape[1] != len(self.pairs) or w.shape[2] != self.G.latent_dim:
            raise ValueError(f"styles must be [B, {len(self.pairs)}, {self.G.latent_dim}], got {tuple(w.shape)}")
        out = torch.empty(w.shape[0], self.row, device=w.device, dtype=torch.float32)
        N.check(N.lib().sx_generator_styles(self.handle, w.data_ptr(), out.data_ptr(), w.shape[0], N.stream_ptr()),
                "sx_generator_styles")
        return out

    def forward(self, styles_all: torch.Tensor, input_noise: torch.Tensor, start_conv: int = 0, save_cache: bool = False,
                precision="fp32", out: Optional[torch.Tensor] = None) -> torch.Tensor:
        prec = N.PRECISIONS[precision]
        self.sync()
        N.require_cuda(styles_all, input_noise)
        if styles_all.dtype != torch.float32 or not styles_all.is_contiguous() or styles_all.shape[1] != self.row:
            raise ValueError(f"styles_all must be contiguous fp32 [B, {self.row}]")
        B = styles_all.shape[0]
        nz, nb, ns = _noise_arg(input_noise)
        S = self.G.image_size
        if ns != S:
            raise ValueError(f"noise map is {ns}x{ns}, generator image_size is {S}")
        self.reserve(B, prec)
        if start_conv > 0 and not self._cache_ok.get(prec, False):
            raise RuntimeError("AttFind suffix forward without a primed clean-prefix cache (call forward(save_cache=True) first)")
        ws = self._ws[prec].buf
        if out is None:
            out = torch.empty(B, 3, S, S, device=styles_all.device, dtype=torch.float32)
        N.check(N.lib().sx_generator_forward(self.handle, styles_all.data_ptr(), nz.data_ptr(), nb, out.data_ptr(), B,
                                             start_conv, 1 if save_cache else 0, prec, ws.data_ptr(), ws.numel(),
                                             N.stream_ptr()), "sx_generator_forward")
        if save_cache:
            self._cache_ok[prec] = True
        return out


class Generator(nn.Module):
    def __init__(self, image_size, latent_dim, network_capacity=16, transparent=False, attn_layers=[], no_const=False,
                 fmap_max=512):
        super().__init__()
        if transparent or attn_layers or no_const:
            raise NotImplementedError("stylex_b200 Generator: transparent / attn_layers / no_const are default-off "
                                      "extras of the reference that are outside the AttFind hot path (SURVEY.md section 2)")
        self.image_size = image_size
        self.latent_dim = latent_dim
        self.num_layers = int(log2(image_size) - 1)
        self.precision = "fp32"
        self.double_backward = False     # True: the twice-differentiable composition (path-length penalty), see module docstring

        filters = [network_capacity * (2 ** (i + 1)) for i in range(self.num_layers)][::-1]

        set_fmap_max = partial(min, fmap_max)
        filters = list(map(set_fmap_max, filters))
        init_channels = filters[0]
        filters = [init_channels, *filters]

        in_out_pairs = zip(filters[:-1], filters[1:])
        self.no_const = no_const

        self.initial_block = nn.Parameter(torch.randn((1, init_channels, 4, 4)))
        self.initial_conv = nn.Conv2d(filters[0], filters[0], 3, padding=1)
        self.blocks = nn.ModuleList([])
        self.attns = nn.ModuleList([])

        for ind, (in_chan, out_chan) in enumerate(in_out_pairs):
            not_first = ind != 0
            not_last = ind != (self.num_layers - 1)
            self.attns.append(None)
            self.blocks.append(GeneratorBlock(latent_dim, in_chan, out_chan, upsample=not_first, upsample_rgb=not_last,
                                              rgba=transparent))
        self._plan = None

    @property
    def num_style_coords(self):
        return sum(b.num_style_coords for b in self.blocks)

    def __getstate__(self):
        """copy.deepcopy / torch.save(model): the native plan (a ctypes handle + device workspaces) is not part of the
        module's state; the copy rebuilds it lazily on its first forward."""
        state = self.__dict__.copy()
        state["_plan"] = None
        return state

    def plan(self) -> GeneratorPlan:
        if self._plan is None:
            object.__setattr__(self, "_plan", GeneratorPlan(self))
        return self._plan

    def forward(self, styles, input_noise, get_style_coords=False):
        """styles [B, L, latent], input_noise [B|1, S, S, 1] -> rgb [B,3,S,S] (, style_coords [B, S_total])."""
        if _wants_grad(styles, *self.parameters()):
            return self._forward_autograd(styles, input_noise, get_style_coords)
        plan = self.plan()
        styles_all = plan.styles(styles)
        rgb = plan.forward(styles_all, input_noise, precision=self.precision)
        if get_style_coords:
            return rgb, styles_all[:, :plan.S].clone()
        return rgb

    def _forward_autograd(self, styles, input_noise, get_style_coords=False):
        """ST:794-825 op by op at module level: every op is an autograd Function over native kernels, so the graph torch
        records gives d(rgb)/d(every generator parameter, styles).  ``initial_conv`` is a plain batch-invariant
        convolution (ST:802,806) and goes through PyTorch like the encoder's convolutions."""
        N.require_cuda(styles, input_noise)
        batch_size = styles.shape[0]
        dd = bool(getattr(self, "double_backward", False))
        for block in self.blocks:      # the 3x3 convs follow the generator's precision (tensor cores in "bf16")
            block.conv1.precision = block.conv2.precision = self.precision
            block.double_backward = block.to_rgb.double_backward = dd
        x = self.initial_conv(self.initial_block).expand(batch_size, -1, -1, -1)
        rgb = None
        coords = []
        for style, block in zip(styles.transpose(0, 1), self.blocks):
            x, rgb, sc = block(x, rgb, style, input_noise)
            coords.append(sc)
        if get_style_coords:
            return rgb, torch.cat(coords, dim=1)
        return rgb

    def forward_from_styles(self, styles_all, input_noise):
        """StyleSpace entry point: styles_all [B, style_row] as returned by ``plan().styles``."""
        return self.plan().forward(styles_all, input_noise, precision=self.precision)
