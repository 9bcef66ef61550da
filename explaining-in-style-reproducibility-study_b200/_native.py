"""ctypes binding of ``libstylex_b200.so`` (C ABI in ``include/stylex_b200.h``).

PyTorch is plumbing here: it owns device memory and the stream; every kernel is ours.  There is no
fallback of any kind -- if the library is missing, not built for this GPU, or a call fails, a
``RuntimeError`` is raised (BASELINE north_star: "no Triton, no multi-backend dispatch and no CPU
fallback").
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from ctypes import POINTER, Structure, byref, c_char_p, c_double, c_float, c_int, c_size_t, c_ulonglong, c_void_p

import torch

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# SX_LIB_PATH: load another build of the SAME sources (profiling builds with -DSX_HALO_DEBUG_KNOBS, A/B experiments)
LIB_PATH = os.environ.get("SX_LIB_PATH") or os.path.join(PKG_DIR, "libstylex_b200.so")
CSRC = os.path.join(PKG_DIR, "csrc")
HEADER = os.path.join(os.path.dirname(PKG_DIR), "include", "stylex_b200.h")

PREC_FP32, PREC_BF16 = 0, 1
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, torch.float32: PREC_FP32, torch.bfloat16: PREC_BF16,
              PREC_FP32: PREC_FP32, PREC_BF16: PREC_BF16}

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared", "-cudart", "static"]


class sx_block_params(Structure):
    _fields_ = [(n, c_void_p) for n in (
        "to_style1_w", "to_style1_b", "to_noise1_w", "to_noise1_b", "conv1_w", "to_style2_w", "to_style2_b",
        "to_noise2_w", "to_noise2_b", "conv2_w", "rgb_style_w", "rgb_style_b", "rgb_conv_w")]


# name -> (restype, argtypes); mirrors include/stylex_b200.h one to one
SIGNATURES = {
    "sx_version": (c_int, []),
    "sx_last_error": (c_char_p, []),
    "sx_device_check": (c_int, []),
    "sx_launch_count": (c_ulonglong, []),
    "sx_conv2dmod_workspace_bytes": (c_size_t, [c_int] * 7),
    "sx_conv2dmod_fwd": (c_int, [c_void_p] * 4 + [c_int] * 7 + [c_float, c_int, c_void_p, c_size_t, c_void_p]),
    "sx_conv2dmod_bwd_workspace_bytes": (c_size_t, [c_int] * 7),
    "sx_conv2dmod_bwd": (c_int, [c_void_p] * 8 + [c_int] * 7 + [c_float, c_int, c_void_p, c_size_t, c_void_p]),
    "sx_upsample2x_bilinear": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "sx_blur3x3_reflect": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "sx_noise_lrelu": (c_int, [c_void_p] * 5 + [c_int] * 6 + [c_void_p]),
    "sx_rgb_add_upsample_blur": (c_int, [c_void_p] * 3 + [c_int] * 5 + [c_void_p]),
    "sx_rgb_prefill_upsample_blur": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "sx_linear_fwd": (c_int, [c_void_p] * 4 + [c_int] * 3 + [c_void_p]),
    "sx_linear_bwd": (c_int, [c_void_p] * 6 + [c_int] * 3 + [c_void_p]),
    "sx_noise_lrelu_bwd_workspace_bytes": (c_size_t, [c_int, c_int]),
    "sx_noise_lrelu_bwd": (c_int, [c_void_p] * 6 + [c_int] * 6 + [c_void_p, c_size_t, c_void_p]),
    "sx_upsample2x_bilinear_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "sx_blur3x3_reflect_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "sx_resize_aa_normalize": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_float),
                                       POINTER(c_float), c_void_p]),
    "sx_resize_aa_normalize_s2d": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_float),
                                           POINTER(c_float), c_void_p]),
    "sx_maxpool3x3s2_nhwc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "sx_stem_s2d_conv_relu": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "sx_generator_create": (c_int, [POINTER(c_int), POINTER(c_int), c_int, c_int, POINTER(c_void_p)]),
    "sx_generator_destroy": (None, [c_void_p]),
    "sx_generator_load": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(sx_block_params), c_void_p]),
    "sx_generator_num_style_coords": (c_int, [c_void_p]),
    "sx_generator_style_row": (c_int, [c_void_p]),
    "sx_generator_styles": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "sx_generator_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "sx_generator_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                     c_size_t, c_void_p]),
    "sx_attfind_minmax": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "sx_attfind_make_styles": (c_int, [c_void_p] * 4 + [c_int, c_int, c_int, c_float, c_void_p]),
    "sx_attfind_make_styles_pairs": (c_int, [c_void_p, ctypes.c_longlong, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                             c_int, c_float, c_void_p]),
    "sx_attfind_scatter_effects": (c_int, [c_void_p] * 3 + [c_int] * 4 + [c_void_p]),
    "sx_attfind_select_workspace_bytes": (c_size_t, [c_int, c_int]),
    "sx_attfind_select": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_double, c_int, c_void_p, c_void_p,
                                  c_size_t, c_void_p]),
    "sx_tc_selftest": (c_int, [c_float, POINTER(c_float)]),
    "sx_profile_enable": (c_int, [c_int]),
    "sx_profile_collect": (c_int, [POINTER(c_double), c_int, POINTER(c_int)]),
}

_lib = None


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/api.cu for sm_100a into the in-tree shared library (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [HEADER]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB_PATH, os.path.join(CSRC, "api.cu")]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed ({r.returncode}):\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """The loaded library; raises if it has not been built (never falls back to anything)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                "stylex_b200 has no CPU / PyTorch fallback for the generator and AttFind kernels.")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.sx_version() != 102:
            raise RuntimeError(f"{LIB_PATH}: version {l.sx_version()} does not match the Python binding (102); rebuild")
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().sx_last_error().decode(errors="replace")
        raise RuntimeError(f"stylex_b200 native call {what} failed (code {rc}): {msg}")


def require_cuda(*tensors: torch.Tensor, same_device: bool = True) -> None:
    """Every native call launches on the CURRENT device's current stream (``stream_ptr``): the tensors it is handed must
    live on that device -- a tensor of another GPU would be dereferenced from the wrong context.  Callers that take a
    reference-style ``cuda_rank`` / ``rank`` argument (``attfind_extraction``, ``StylEx``) make that device current
    themselves; anything else raises here instead of faulting on the device."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("stylex_b200 ops run on CUDA tensors only (there is no CPU fallback); got a "
                               f"{t.device} tensor")
        if cur is None:
            cur = torch.cuda.current_device()
        if same_device and t.device.index != cur:
            raise RuntimeError(f"stylex_b200 native call: tensor on {t.device} but the current CUDA device is cuda:{cur}; "
                               f"wrap the call in `with torch.cuda.device({t.device.index})` (or torch.cuda.set_device)")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def f32c(t: torch.Tensor) -> torch.Tensor:
    """contiguous fp32 view/copy (the C ABI takes dense fp32 tensors in the reference's layout)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def profile_enable(on: bool) -> None:
    check(lib().sx_profile_enable(1 if on else 0), "sx_profile_enable")


def profile_collect() -> dict:
    """{kind: {"launches", "ms", "flops", "bytes"}} since profile_enable(True); synchronises the recorded events."""
    rows = (c_double * (5 * 64))()
    n = c_int(0)
    check(lib().sx_profile_collect(rows, 64, byref(n)), "sx_profile_collect")
    return {int(rows[5 * i]): {"launches": int(rows[5 * i + 1]), "ms": rows[5 * i + 2], "flops": rows[5 * i + 3],
                               "bytes": rows[5 * i + 4]} for i in range(n.value)}


def launch_count() -> int:
    return int(lib().sx_launch_count())


_device_ok = {}


def device_check() -> None:
    dev = torch.cuda.current_device()
    if dev not in _device_ok:
        check(lib().sx_device_check(), "sx_device_check")
        _device_ok[dev] = True


class Workspace:
    """A grow-only 256-byte aligned device scratch buffer owned by torch's allocator."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes: int, device) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != torch.device(device):
            self.buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
            assert self.buf.data_ptr() % 256 == 0
        return self.buf
