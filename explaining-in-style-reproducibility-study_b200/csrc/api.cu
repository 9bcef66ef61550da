// api.cu -- the extern "C" surface declared in include/stylex_b200.h.  One translation unit: every kernel
// header is included here and compiled for sm_100a only.
#include <type_traits>
#include <vector>

#include "attfind.cuh"
#include "bandwidth.cuh"
#include "common.cuh"
#include "bwd_ops.cuh"
#include "conv_bwd.cuh"
#include "conv_simt.cuh"
#include "conv_tc.cuh"
#include "conv_tc_halo.cuh"
#include "generator.cuh"
#include "preprocess.cuh"
#include "stem_tc.cuh"
#include "wgrad_tc.cuh"

using namespace sx;

static inline cudaStream_t S(sx_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int sx_version(void) { return SX_VERSION; }
const char* sx_last_error(void) { return last_error().c_str(); }
unsigned long long sx_launch_count(void) { return launch_counter().load(); }

int sx_device_check(void) {
  int dev = 0;
  SX_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  SX_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  SX_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) return fail(SX_EUNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev, major, minor);
  return SX_OK;
}

// -------------------------------------------------------------------------------------------------
// L1 ops, NCHW fp32 boundary
// -------------------------------------------------------------------------------------------------
struct Conv2dModWs {
  size_t xmod, wpk, wsq, dcoef, total;
};
static Conv2dModWs conv2dmod_ws(int B, int Ci, int Co, int H, int W, int k, int precision) {
  const size_t es = precision == SX_PREC_BF16 ? 2 : 4;
  Conv2dModWs w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  w.xmod = take((size_t)B * H * W * Ci * es);
  w.wpk = take((size_t)k * k * Ci * Co * es);
  w.wsq = take((size_t)Ci * Co * 4);
  w.dcoef = take((size_t)B * Co * 4);
  w.total = off;
  return w;
}

size_t sx_conv2dmod_workspace_bytes(int B, int Ci, int Co, int H, int W, int k, int precision) {
  if (B < 0 || Ci < 1 || Co < 1 || H < 1 || W < 1 || k < 1) return 0;
  return conv2dmod_ws(B, Ci, Co, H, W, k, precision).total;
}

int sx_conv2dmod_fwd(const float* x, const float* weight, const float* style, float* out, int B, int Ci, int Co, int H, int W,
                     int k, int demod, float eps, int precision, void* workspace, size_t ws_bytes, sx_stream_t stream) {
  SX_REQUIRE(B >= 0 && Ci >= 1 && Co >= 1 && H >= 1 && W >= 1, "bad shape B=%d Ci=%d Co=%d H=%d W=%d", B, Ci, Co, H, W);
  if (B == 0) return SX_OK;
  SX_REQUIRE(x && weight && style && out && workspace, "null argument");
  SX_REQUIRE(k == 1 || k == 3, "kernel size %d not supported (1 or 3)", k);
  SX_REQUIRE(precision == SX_PREC_FP32 || precision == SX_PREC_BF16, "precision=%d", precision);
  SX_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  if (B == 0) return SX_OK;
  if (precision == SX_PREC_BF16 && !tc::tc_shape_supported(Ci, Co, H, W, k))
    return fail(SX_EUNSUPPORTED, "bf16 tcgen05 Conv2DMod needs k=3, Ci%%32==0, Co in {32,64,128,256n}, square power-of-two H=W>=4 (got Ci=%d Co=%d H=%d W=%d k=%d)", Ci, Co, H, W, k);
  const Conv2dModWs L = conv2dmod_ws(B, Ci, Co, H, W, k, precision);
  if (ws_bytes < L.total) return fail(SX_ENOMEM, "workspace %zu bytes < required %zu", ws_bytes, L.total);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  cudaStream_t st = S(stream);
  float* wsq = reinterpret_cast<float*>(ws + L.wsq);
  float* dcoef = reinterpret_cast<float*>(ws + L.dcoef);
  const int HW = H * W;
  dim3 tgrid((HW + 31) / 32, (Ci + 31) / 32, B);
  ConvEpilogue ep{};
  ep.dcoef = demod ? dcoef : nullptr;
  ep.dcoef_stride = Co;
  ep.out = out;
  ep.out_nchw_f32 = 1;
  if (precision == SX_PREC_FP32) {
    float* xmod = reinterpret_cast<float*>(ws + L.xmod);
    float* wpk = reinterpret_cast<float*>(ws + L.wpk);
    pack_weights_kernel<<<ew_grid((long long)Ci * Co, 256), 256, 0, st>>>(weight, wpk, nullptr, wsq, Co, Ci, k * k);
    SX_CHECK_LAUNCH();
    nchw_to_nhwc_modulate_kernel<float><<<tgrid, 256, 0, st>>>(x, style, xmod, Ci, HW);
    SX_CHECK_LAUNCH();
  } else {
    __nv_bfloat16* xmod = reinterpret_cast<__nv_bfloat16*>(ws + L.xmod);
    __nv_bfloat16* wbf = reinterpret_cast<__nv_bfloat16*>(ws + L.wpk);
    pack_weights_kernel<<<ew_grid((long long)Ci * Co, 256), 256, 0, st>>>(weight, nullptr, wbf, wsq, Co, Ci, k * k);
    SX_CHECK_LAUNCH();
    nchw_to_nhwc_modulate_kernel<__nv_bfloat16><<<tgrid, 256, 0, st>>>(x, style, xmod, Ci, HW);
    SX_CHECK_LAUNCH();
  }
  if (demod) {
    DemodParams dp{};
    dp.conv[0] = DemodConv{wsq, Ci, Co, 0, 0};
    dp.first_conv = 0;
    dp.num_convs = 1;
    dp.styles = style;
    dp.style_stride = Ci;
    dp.dcoef = dcoef;
    dp.dcoef_stride = Co;
    dp.eps = eps;
    SX_TRY(launch_demod(dp, B, Ci, Co, st));
  }
  if (precision == SX_PREC_FP32) {
    ConvSimtParams p;
    p.x = reinterpret_cast<float*>(ws + L.xmod);
    p.x_bstride = (long long)HW * Ci;
    p.wpk = reinterpret_cast<float*>(ws + L.wpk);
    p.B = B; p.Ci = Ci; p.Co = Co; p.H = H; p.W = W; p.KS = k; p.ep = ep;
    return launch_conv_simt(p, st);
  }
  return tc::launch_conv_bf16(reinterpret_cast<__nv_bfloat16*>(ws + L.xmod), reinterpret_cast<__nv_bfloat16*>(ws + L.wpk), B, Ci, Co,
                            H, W, k, ep, st);
}

// ---- Conv2DMod backward (first order) -- conv_bwd.cuh, wgrad_tc.cuh ------------------------------------------------
// SX_PREC_FP32: FFMA implicit GEMMs (the <= 1e-4 parity mode).  SX_PREC_BF16: the two big contractions run on the tensor
// cores where the kernels take the shape -- dgrad on the tcgen05 conv kernels with flipped / transposed weights, wgrad as a
// tcgen05 GEMM over K = pixels (wgrad_tc.cuh) -- and on the FFMA kernels otherwise (k = 1, 4x4 maps, odd channel counts);
// the small per-sample reductions stay fp32 in both modes.
struct Conv2dModBwdWs {
  size_t gz, xm, wT, wsq, dcoef, gdot, gm1, msq, partial, gz_bf, wT_bf, gzp_bf, xmp_bf, total;
  int splits, splits_tc;
  bool dgrad_tc, wgrad_tc;
};
static Conv2dModBwdWs conv2dmod_bwd_ws(int B, int Ci, int Co, int H, int W, int k, int precision) {
  Conv2dModBwdWs w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t HW = (size_t)H * W;
  // SX_BWD_NO_TC_DGRAD / SX_BWD_NO_TC_WGRAD: keep one contraction on the FFMA kernels (A/B measurements, bisecting)
  w.dgrad_tc = precision == SX_PREC_BF16 && tc::tc_shape_supported(Co, Ci, H, W, k) && !getenv("SX_BWD_NO_TC_DGRAD");
  w.wgrad_tc = precision == SX_PREC_BF16 && tc::wgrad_tc_supported(Ci, Co, H, W, k) && !getenv("SX_BWD_NO_TC_WGRAD");
  w.splits = wgrad_splits(B, Ci, Co, H, W, k);
  w.splits_tc = w.wgrad_tc ? tc::wgrad_tc_splits(B, Ci, Co, H, W, k) : 0;
  w.gz = take((size_t)B * HW * Co * 4);
  w.xm = take((size_t)B * HW * Ci * 4);
  w.wT = take((size_t)k * k * Ci * Co * 4);
  w.wsq = take((size_t)Ci * Co * 4);
  w.dcoef = take((size_t)B * Co * 4);
  w.gdot = take((size_t)B * Co * 4);
  w.gm1 = take((size_t)B * Ci * 4);
  w.msq = take((size_t)B * Ci * 4);
  const int smax = w.splits > w.splits_tc ? w.splits : w.splits_tc;
  w.partial = take((size_t)smax * k * k * Co * Ci * 4);
  if (w.dgrad_tc) {
    w.gz_bf = take((size_t)B * HW * Co * 2);
    w.wT_bf = take((size_t)k * k * Ci * Co * 2);
  }
  if (w.wgrad_tc) {
    w.gzp_bf = take((size_t)B * HW * Co * 2);
    w.xmp_bf = take((size_t)k * B * HW * Ci * 2);   // k copies shifted along x (wgrad_tc.cuh)
  }
  w.total = off;
  return w;
}

size_t sx_conv2dmod_bwd_workspace_bytes(int B, int Ci, int Co, int H, int W, int k, int precision) {
  if (B < 1 || Ci < 1 || Co < 1 || H < 1 || W < 1 || k < 1) return 0;
  return conv2dmod_bwd_ws(B, Ci, Co, H, W, k, precision).total;
}

int sx_conv2dmod_bwd(const float* x, const float* weight, const float* style, const float* out, const float* grad_out,
                     float* grad_x, float* grad_weight, float* grad_style, int B, int Ci, int Co, int H, int W, int k,
                     int demod, float eps, int precision, void* workspace, size_t ws_bytes, sx_stream_t stream) {
  SX_REQUIRE(B >= 0 && Ci >= 1 && Co >= 1 && H >= 1 && W >= 1, "bad shape B=%d Ci=%d Co=%d H=%d W=%d", B, Ci, Co, H, W);
  SX_REQUIRE(k == 1 || k == 3, "kernel size %d not supported (1 or 3)", k);
  SX_REQUIRE(precision == SX_PREC_FP32 || precision == SX_PREC_BF16, "precision=%d", precision);
  SX_REQUIRE(grad_weight, "null argument");
  cudaStream_t st = S(stream);
  if (B == 0) {   // empty batch: the weight gradient is zero, nothing else to write
    SX_CUDA(cudaMemsetAsync(grad_weight, 0, (size_t)Co * Ci * k * k * 4, st));
    return SX_OK;
  }
  SX_REQUIRE(x && weight && style && grad_out && grad_x && grad_style && workspace, "null argument");
  SX_REQUIRE(!demod || out, "demod backward needs the forward output");
  SX_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  const Conv2dModBwdWs L = conv2dmod_bwd_ws(B, Ci, Co, H, W, k, precision);
  if (ws_bytes < L.total) return fail(SX_ENOMEM, "workspace %zu bytes < required %zu", ws_bytes, L.total);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  auto F = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  auto H16 = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  float *gz = F(L.gz), *xm = F(L.xm), *wT = F(L.wT), *wsq = F(L.wsq), *dcoef = F(L.dcoef), *gdot = F(L.gdot), *gm1 = F(L.gm1),
        *msq = F(L.msq), *partial = F(L.partial);
  const int HW = H * W, taps = k * k;
  const long long n_g = (long long)B * Co * HW, n_x = (long long)B * Ci * HW;
  // packed operands: wsq[i][o], flipped / transposed weights for the dgrad
  pack_weights_kernel<<<ew_grid((long long)Ci * Co, 256), 256, 0, st>>>(weight, nullptr, nullptr, wsq, Co, Ci, taps);
  SX_CHECK_LAUNCH();
  if (demod) {
    DemodParams dp{};
    dp.conv[0] = DemodConv{wsq, Ci, Co, 0, 0};
    dp.first_conv = 0; dp.num_convs = 1; dp.styles = style; dp.style_stride = Ci; dp.dcoef = dcoef; dp.dcoef_stride = Co; dp.eps = eps;
    SX_TRY(launch_demod(dp, B, Ci, Co, st));
  }
  const float* dscale = demod ? dcoef : nullptr;
  const dim3 grid_g((HW + 31) / 32, (Co + 31) / 32, B), grid_x((HW + 31) / 32, (Ci + 31) / 32, B);
  // ---- dgrad: gxm = conv(W^T flipped, g * d) -> grad_x (NCHW fp32)
  {
    ConvEpilogue ep{};
    ep.out = grad_x; ep.out_nchw_f32 = 1;
    if (L.dgrad_tc) {
      pack_weights_dgrad_bf16_kernel<<<ew_grid((long long)Ci * Co, 256), 256, 0, st>>>(weight, H16(L.wT_bf), Co, Ci, taps);
      SX_CHECK_LAUNCH();
      nchw_to_nhwc_scale_kernel<__nv_bfloat16><<<grid_g, 256, 0, st>>>(grad_out, dscale, H16(L.gz_bf), Co, HW);
      SX_CHECK_LAUNCH();
      SX_TRY(tc::launch_conv_bf16(H16(L.gz_bf), H16(L.wT_bf), B, Co, Ci, H, W, k, ep, st));
    } else {
      pack_weights_dgrad_kernel<<<ew_grid((long long)Ci * Co, 256), 256, 0, st>>>(weight, wT, Co, Ci, taps);
      SX_CHECK_LAUNCH();
      nchw_to_nhwc_scale_kernel<float><<<grid_g, 256, 0, st>>>(grad_out, dscale, gz, Co, HW);
      SX_CHECK_LAUNCH();
      ConvSimtParams p;
      p.x = gz; p.x_bstride = (long long)HW * Co; p.wpk = wT;
      p.B = B; p.Ci = Co; p.Co = Ci; p.H = H; p.W = W; p.KS = k;
      p.ep = ep;
      SX_TRY(launch_conv_simt(p, st));
    }
  }
  // gm1 = <gxm, x> per plane, then grad_x = gxm * (style + 1) in place
  plane_dot_kernel<<<(unsigned)((long long)B * Ci), 256, 0, st>>>(grad_x, x, style, gm1, HW);
  SX_CHECK_LAUNCH();
  if (demod) {
    plane_dot_kernel<<<(unsigned)((long long)B * Co), 256, 0, st>>>(const_cast<float*>(grad_out), out, nullptr, gdot, HW);
    SX_CHECK_LAUNCH();
    const long long n = (long long)B * (Ci > Co ? Ci : Co);
    demod_grad_prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(gdot, dcoef, (long long)B * Co, style, msq, (long long)B * Ci);
    SX_CHECK_LAUNCH();
  }
  style_grad_kernel<<<(unsigned)(((long long)B * Ci * 32 + 255) / 256), 256, 0, st>>>(gm1, style, demod ? gdot : nullptr, wsq, grad_style, B, Ci, Co);
  SX_CHECK_LAUNCH();
  // ---- wgrad over the modulated activations
  int splits = L.splits;
  if (L.wgrad_tc) {
    splits = L.splits_tc;
    nchw_scale_bf16_kernel<<<ew_grid(n_g, 256), 256, 0, st>>>(grad_out, dscale, 0.f, H16(L.gzp_bf), HW, n_g);
    SX_CHECK_LAUNCH();
    nchw_scale_shift_bf16_kernel<<<ew_grid(n_x, 256), 256, 0, st>>>(x, style, H16(L.xmp_bf), HW, W, k, n_x);
    SX_CHECK_LAUNCH();
    SX_TRY(tc::launch_wgrad_tc(H16(L.gzp_bf), H16(L.xmp_bf), partial, B, Ci, Co, H, W, k, splits, st));
  } else {
    if (L.dgrad_tc) {   // the FFMA wgrad reads fp32 NHWC g * d, which the tensor-core dgrad did not produce
      nchw_to_nhwc_scale_kernel<float><<<grid_g, 256, 0, st>>>(grad_out, dscale, gz, Co, HW);
      SX_CHECK_LAUNCH();
    }
    nchw_to_nhwc_modulate_kernel<float><<<grid_x, 256, 0, st>>>(x, style, xm, Ci, HW);
    SX_CHECK_LAUNCH();
    WgradParams p;
    p.gz = gz; p.xm = xm; p.partial = partial;
    p.B = B; p.H = H; p.W = W; p.Ci = Ci; p.Co = Co; p.KS = k; p.splits = splits;
    const long long M = (long long)B * HW;
    p.pix_per_split = ((M + splits - 1) / splits + 15) / 16 * 16;
    dim3 grid((unsigned)(((Co + 63) / 64) * ((Ci + 63) / 64)), (unsigned)taps, (unsigned)splits);
    wgrad_simt_kernel<<<grid, 256, 0, st>>>(p);
    SX_CHECK_LAUNCH();
  }
  wgrad_reduce_kernel<<<ew_grid((long long)Co * Ci, 256), 256, 0, st>>>(partial, splits, weight, demod ? gdot : nullptr, msq, grad_weight,
                                                                      B, Co, Ci, taps);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

int sx_upsample2x_bilinear(const float* x, float* out, int B, int C, int H, int W, sx_stream_t stream) {
  SX_REQUIRE(B >= 0 && C >= 0 && H >= 1 && W >= 1, "bad shape");
  const long long planes = (long long)B * C;
  if (planes == 0) return SX_OK;
  SX_REQUIRE(x && out, "null argument");
  upsample2x_nchw_kernel<<<ew_grid(planes * 4 * H * W, 256), 256, 0, S(stream)>>>(x, out, planes, H, W);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

int sx_blur3x3_reflect(const float* x, float* out, int B, int C, int H, int W, sx_stream_t stream) {
  SX_REQUIRE(x && out, "null argument");
  SX_REQUIRE(B >= 0 && C >= 0 && H >= 2 && W >= 2, "blur with reflect border needs H, W >= 2");
  const long long planes = (long long)B * C;
  if (planes == 0) return SX_OK;
  blur_nchw_kernel<<<ew_grid(planes * H * W, 256), 256, 0, S(stream)>>>(x, out, planes, H, W);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

int sx_noise_lrelu(const float* x, const float* inoise, const float* noise_w, const float* noise_b, float* out, int B, int C,
                   int H, int W, int noise_batch, int noise_size, sx_stream_t stream) {
  SX_REQUIRE(x && inoise && noise_w && noise_b && out, "null argument");
  SX_REQUIRE(noise_batch == 1 || noise_batch == B, "noise batch %d must be 1 or B=%d", noise_batch, B);
  SX_REQUIRE(H <= noise_size && W <= noise_size, "noise map %d smaller than the activation %dx%d", noise_size, H, W);
  const long long total = (long long)B * C * H * W;
  if (total == 0) return SX_OK;
  noise_lrelu_nchw_kernel<<<ew_grid(total, 256), 256, 0, S(stream)>>>(x, inoise, noise_w, noise_b, out, B, C, H, W, noise_batch,
                                                                      noise_size);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

int sx_rgb_add_upsample_blur(const float* rgb, const float* prev, float* out, int B, int C, int H, int W, int upsample,
                             sx_stream_t stream) {
  SX_REQUIRE(rgb && out, "null argument");
  SX_REQUIRE(B >= 0 && C >= 0 && H >= 1 && W >= 1, "bad shape");
  const long long planes = (long long)B * C;
  if (planes == 0) return SX_OK;
  cudaStream_t st = S(stream);
  if (!upsample) {
    rgb_tail_nchw_kernel<<<ew_grid(planes * H * W, 256), 256, 0, st>>>(rgb, prev, out, planes * H * W);
  } else {
    rgb_tail_up_blur_nchw_kernel<<<ew_grid(planes * 4 * H * W, 256), 256, 0, st>>>(rgb, prev, out, planes, H, W);
  }
  SX_CHECK_LAUNCH();
  return SX_OK;
}

int sx_rgb_prefill_upsample_blur(const float* prev, int prev_batch, float* out, int B, int h, int w, sx_stream_t stream) {
  SX_REQUIRE(prev && out, "null argument");
  SX_REQUIRE(B >= 0 && h >= 2 && w >= 2, "bad shape");
  SX_REQUIRE(prev_batch == 1 || prev_batch == B, "prev batch %d must be 1 or B=%d", prev_batch, B);
  return launch_rgb_prev_up_blur(prev, prev_batch == 1 ? 0 : (long long)3 * h * w, out, B, 2 * h, 2 * w, S(stream));
}

int sx_linear_fwd(const float* x, const float* weight, const float* bias, float* out, int B, int in_f, int out_f,
                  sx_stream_t stream) {
  SX_REQUIRE(x && weight && out, "null argument");
  SX_REQUIRE(B >= 0 && in_f >= 1 && out_f >= 1, "bad shape");
  if (B == 0) return SX_OK;
  linear_kernel<<<ew_grid((long long)B * out_f * 32, 256, 16), 256, 0, S(stream)>>>(x, weight, bias, out, B, in_f, out_f);
  SX_CHECK_LAUNCH();
  return SX_OK;
}


// ---- backward of the bandwidth ops (bwd_ops.cuh) ---------------------------------------------------------------------
int sx_linear_bwd(const float* x, const float* weight, const float* grad_out, float* grad_x, float* grad_weight, float* grad_bias,
                  int B, int in_f, int out_f, sx_stream_t stream) {
  SX_REQUIRE(B >= 0 && in_f >= 1 && out_f >= 1, "bad shape");
  SX_REQUIRE(grad_weight, "null argument");
  SX_REQUIRE(B == 0 || (x && weight && grad_out && grad_x), "null argument");
  cudaStream_t st = S(stream);
  linear_bwd_w_kernel<<<ew_grid((long long)out_f * in_f, 256), 256, 0, st>>>(grad_out, x, grad_weight, grad_bias, B, in_f, out_f);
  SX_CHECK_LAUNCH();
  if (B == 0) return SX_OK;
  linear_bwd_x_kernel<<<ew_grid((long long)B * in_f, 256), 256, 0, st>>>(grad_out, weight, grad_x, B, in_f, out_f);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

size_t sx_noise_lrelu_bwd_workspace_bytes(int B, int C) { return B < 0 || C < 0 ? 0 : align_up((size_t)2 * B * C * 4, 256); }

int sx_noise_lrelu_bwd(const float* out, const float* grad_out, const float* inoise, float* grad_x, float* grad_noise_w,
                       float* grad_noise_b, int B, int C, int H, int W, int noise_batch, int noise_size, void* workspace,
                       size_t ws_bytes, sx_stream_t stream) {
  SX_REQUIRE(B >= 0 && C >= 1 && H >= 1 && W >= 1, "bad shape");
  SX_REQUIRE(grad_noise_w && grad_noise_b, "null argument");
  cudaStream_t st = S(stream);
  if (B == 0) {
    SX_CUDA(cudaMemsetAsync(grad_noise_w, 0, (size_t)C * 4, st));
    SX_CUDA(cudaMemsetAsync(grad_noise_b, 0, (size_t)C * 4, st));
    return SX_OK;
  }
  SX_REQUIRE(out && grad_out && inoise && grad_x && workspace, "null argument");
  SX_REQUIRE(noise_batch == 1 || noise_batch == B, "noise batch %d must be 1 or B=%d", noise_batch, B);
  SX_REQUIRE(H <= noise_size && W <= noise_size, "noise map %d smaller than the activation %dx%d", noise_size, H, W);
  SX_REQUIRE(ws_bytes >= sx_noise_lrelu_bwd_workspace_bytes(B, C), "workspace too small");
  float* pw = reinterpret_cast<float*>(workspace);
  float* pb = pw + (size_t)B * C;
  noise_lrelu_bwd_kernel<<<(unsigned)((long long)B * C), 256, 0, st>>>(out, grad_out, inoise, grad_x, pw, pb, C, H, W, noise_batch, noise_size);
  SX_CHECK_LAUNCH();
  noise_param_reduce_kernel<<<(C + 127) / 128, 128, 0, st>>>(pw, pb, grad_noise_w, grad_noise_b, B, C);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

int sx_upsample2x_bilinear_bwd(const float* grad_out, float* grad_x, int B, int C, int H, int W, sx_stream_t stream) {
  SX_REQUIRE(B >= 0 && C >= 0 && H >= 1 && W >= 1, "bad shape");
  const long long planes = (long long)B * C;
  if (planes == 0) return SX_OK;
  SX_REQUIRE(grad_out && grad_x, "null argument");
  upsample2x_bwd_nchw_kernel<<<ew_grid(planes * H * W, 256), 256, 0, S(stream)>>>(grad_out, grad_x, planes, H, W);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

int sx_blur3x3_reflect_bwd(const float* grad_out, float* grad_x, int B, int C, int H, int W, sx_stream_t stream) {
  SX_REQUIRE(B >= 0 && C >= 0 && H >= 2 && W >= 2, "blur with reflect border needs H, W >= 2");
  const long long planes = (long long)B * C;
  if (planes == 0) return SX_OK;
  SX_REQUIRE(grad_out && grad_x, "null argument");
  blur_bwd_nchw_kernel<<<ew_grid(planes * H * W, 256), 256, 0, S(stream)>>>(grad_out, grad_x, planes, H, W);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

int sx_resize_aa_normalize(const float* in, void* out, int out_bf16, int B, int IH, int IW, int OH, int OW, int normalize,
                           const float* mean3, const float* std3, sx_stream_t stream) {
  SX_REQUIRE(in && out, "null argument");
  SX_REQUIRE(B >= 0 && IH >= 1 && IW >= 1 && OH >= 1 && OW >= 1, "bad shape");
  SX_REQUIRE(!normalize || (mean3 && std3), "normalize needs mean and std");
  Norm3 nm{};
  nm.on = normalize ? 1 : 0;
  for (int c = 0; c < 3; ++c) {
    nm.mean[c] = normalize ? mean3[c] : 0.f;
    nm.std[c] = normalize ? std3[c] : 1.f;
  }
  if (out_bf16) return launch_resize_aa_normalize<__nv_bfloat16>(in, reinterpret_cast<__nv_bfloat16*>(out), B, IH, IW, OH, OW, nm, S(stream));
  return launch_resize_aa_normalize<float>(in, reinterpret_cast<float*>(out), B, IH, IW, OH, OW, nm, S(stream));
}

int sx_resize_aa_normalize_s2d(const float* in, void* out, int out_bf16, int B, int IH, int IW, int OH, int OW, int normalize,
                               const float* mean3, const float* std3, sx_stream_t stream) {
  SX_REQUIRE(in && out, "null argument");
  SX_REQUIRE(B >= 0 && IH >= 1 && IW >= 1 && OH >= 2 && OW >= 2, "bad shape");
  SX_REQUIRE(!normalize || (mean3 && std3), "normalize needs mean and std");
  Norm3 nm{};
  nm.on = normalize ? 1 : 0;
  for (int c = 0; c < 3; ++c) {
    nm.mean[c] = normalize ? mean3[c] : 0.f;
    nm.std[c] = normalize ? std3[c] : 1.f;
  }
  if (out_bf16) return launch_resize_aa_normalize_s2d<__nv_bfloat16>(in, reinterpret_cast<__nv_bfloat16*>(out), B, IH, IW, OH, OW, nm, S(stream));
  return launch_resize_aa_normalize_s2d<float>(in, reinterpret_cast<float*>(out), B, IH, IW, OH, OW, nm, S(stream));
}

int sx_maxpool3x3s2_nhwc(const void* in, void* out, int is_bf16, int B, int H, int W, int C, sx_stream_t stream) {
  SX_REQUIRE(B >= 0 && H >= 1 && W >= 1 && C >= 1, "bad shape");
  if (B == 0) return SX_OK;
  SX_REQUIRE(in && out, "null argument");
  if (is_bf16)
    return launch_maxpool3x3s2_nhwc<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), B, H, W, C, S(stream));
  return launch_maxpool3x3s2_nhwc<float>(reinterpret_cast<const float*>(in), reinterpret_cast<float*>(out), B, H, W, C, S(stream));
}

int sx_stem_s2d_conv_relu(const void* x, const void* w_taps, const float* bias, void* out, int B, int Hin, int Win, int fuse_pool,
                          sx_stream_t stream) {
  SX_REQUIRE(B >= 0 && Hin >= 4 && Win >= 4, "bad shape");
  if (B == 0) return SX_OK;
  SX_REQUIRE(x && w_taps && bias && out, "null argument");
  SX_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_taps) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(out) & 15) == 0, "stem: 16-byte aligned buffers required");
  return tc::launch_stem_s2d(reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(w_taps), bias,
                             reinterpret_cast<__nv_bfloat16*>(out), B, Hin, Win, fuse_pool, S(stream));
}

// -------------------------------------------------------------------------------------------------
// generator plan
// -------------------------------------------------------------------------------------------------
int sx_generator_create(const int* ci, const int* co, int num_blocks, int latent_dim, sx_generator_t** out) {
  return generator_create(ci, co, num_blocks, latent_dim, out);
}
void sx_generator_destroy(sx_generator_t* g) { delete g; }

int sx_generator_load(sx_generator_t* g, const float* initial_block, const float* initial_conv_w, const float* initial_conv_b,
                      const sx_block_params* blocks, sx_stream_t stream) {
  return generator_load(g, initial_block, initial_conv_w, initial_conv_b, blocks, S(stream));
}
int sx_generator_num_style_coords(const sx_generator_t* g) { return g ? g->S : 0; }
int sx_generator_style_row(const sx_generator_t* g) { return g ? g->style_row : 0; }

int sx_generator_styles(const sx_generator_t* g, const float* w, float* styles, int B, sx_stream_t stream) {
  return generator_styles(g, w, styles, B, S(stream));
}
size_t sx_generator_workspace_bytes(const sx_generator_t* g, int max_batch, int precision) {
  if (!g) return 0;
  return gen_workspace(g, max_batch, precision).total;
}
int sx_generator_forward(sx_generator_t* g, const float* styles, const float* inoise, int noise_batch, float* rgb_out, int B,
                         int start_conv, int save_cache, int precision, void* workspace, size_t workspace_bytes,
                         sx_stream_t stream) {
  return generator_forward(g, styles, inoise, noise_batch, rgb_out, B, start_conv, save_cache, precision, workspace,
                           workspace_bytes, S(stream));
}

// -------------------------------------------------------------------------------------------------
// AttFind
// -------------------------------------------------------------------------------------------------
int sx_attfind_minmax(const float* style_coords, int N, int Sc, int row_stride, float* minima, float* maxima, sx_stream_t stream) {
  SX_REQUIRE(style_coords && minima && maxima, "null argument");
  SX_REQUIRE(N >= 1 && Sc >= 1 && row_stride >= Sc, "bad shape N=%d S=%d stride=%d (no images pass the threshold check)", N, Sc, row_stride);
  minmax_kernel<<<(Sc + 255) / 256, 256, 0, S(stream)>>>(style_coords, N, Sc, row_stride, minima, maxima);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

int sx_attfind_make_styles(const float* base_row, const float* minima, const float* maxima, float* out, int style_row,
                           int first_sindex, int num_coords, float shift_size, sx_stream_t stream) {
  SX_REQUIRE(base_row && minima && maxima && out, "null argument");
  SX_REQUIRE(style_row >= 1 && first_sindex >= 0 && num_coords >= 0 && first_sindex + num_coords <= style_row, "bad range");
  if (num_coords == 0) return SX_OK;
  make_styles_kernel<<<2 * num_coords, 256, 0, S(stream)>>>(base_row, minima, maxima, out, style_row, first_sindex, shift_size);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

int sx_attfind_make_styles_pairs(const float* styles_all, long long row_stride, const float* minima, const float* maxima, float* out,
                                 int style_row, int Sc, const int* latent_idx, const int* columns, int count, float shift_size,
                                 sx_stream_t stream) {
  SX_REQUIRE(styles_all && minima && maxima && out, "null argument");
  SX_REQUIRE(style_row >= 1 && Sc >= 1 && Sc <= style_row && row_stride >= style_row && count >= 0, "bad range");
  if (count == 0) return SX_OK;
  SX_REQUIRE(latent_idx && columns, "null index list");
  make_styles_pairs_kernel<<<count, 256, 0, S(stream)>>>(styles_all, row_stride, minima, maxima, out, style_row, Sc, latent_idx,
                                                         columns, shift_size);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

int sx_attfind_scatter_effects(const float* logits, const float* base_logits, float* effects, int n, int Sc, int first_sindex,
                               int num_coords, sx_stream_t stream) {
  SX_REQUIRE(logits && base_logits && effects, "null argument");
  SX_REQUIRE(n >= 0 && Sc >= 1 && first_sindex >= 0 && num_coords >= 0 && first_sindex + num_coords <= Sc, "bad range");
  if (num_coords == 0) return SX_OK;
  scatter_effects_kernel<<<(2 * num_coords + 255) / 256, 256, 0, S(stream)>>>(logits, base_logits, effects, n, Sc, first_sindex,
                                                                              2 * num_coords);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

struct SelectWs {
  size_t colmean, images_effect, row_class, picked, num_picked, total;
};
static SelectWs select_ws(int N, int Sc, int k) {
  SelectWs w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  w.colmean = take((size_t)2 * Sc * 8);
  w.images_effect = take((size_t)N * 8);
  w.row_class = take((size_t)N * 4);
  w.picked = take((size_t)k * 4);
  w.num_picked = take(16);
  w.total = off;
  return w;
}
size_t sx_attfind_select_workspace_bytes(int N, int Sc) {
  if (N < 0 || Sc < 1) return 0;
  return select_ws(N, Sc, 4096).total;
}
}  // extern "C"

template <typename E>
static int select_run(const E* effects, const float* base_logits, int N, int Sc, int k, double max_image_effect, int class_index,
                      int* picks, SelectState stt, cudaStream_t st) {
  select_init_kernel<<<(N + 255) / 256, 256, 0, st>>>(base_logits, N, class_index, stt);
  SX_CHECK_LAUNCH();
  for (int r = 0; r < k; ++r) {
    colmean_kernel<E><<<(2 * Sc + 127) / 128, 128, 0, st>>>(effects, N, Sc, class_index, max_image_effect, stt);
    SX_CHECK_LAUNCH();
    argmax_update_kernel<E><<<1, 1024, 0, st>>>(effects, N, Sc, class_index, stt, picks, r);
    SX_CHECK_LAUNCH();
  }
  return SX_OK;
}

extern "C" {

int sx_attfind_select(const void* effects, int effects_f64, const float* base_logits, int N, int Sc, int k,
                      double max_image_effect, int class_index, int* picks, void* workspace, size_t workspace_bytes,
                      sx_stream_t stream) {
  SX_REQUIRE(effects && picks && workspace, "null argument");
  SX_REQUIRE(N >= 1 && Sc >= 1 && k >= 1 && k <= 4096, "bad shape N=%d S=%d k=%d", N, Sc, k);
  SX_REQUIRE(class_index == 0 || class_index == 1, "class_index=%d", class_index);
  SX_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  const SelectWs L = select_ws(N, Sc, 4096);
  if (workspace_bytes < L.total) return fail(SX_ENOMEM, "workspace %zu bytes < required %zu", workspace_bytes, L.total);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  SelectState stt;
  stt.colmean = reinterpret_cast<double*>(ws + L.colmean);
  stt.images_effect = reinterpret_cast<double*>(ws + L.images_effect);
  stt.row_class = reinterpret_cast<int*>(ws + L.row_class);
  stt.picked = reinterpret_cast<int*>(ws + L.picked);
  stt.num_picked = reinterpret_cast<int*>(ws + L.num_picked);
  stt.best = nullptr;
  if (effects_f64)
    return select_run<double>(reinterpret_cast<const double*>(effects), base_logits, N, Sc, k, max_image_effect, class_index, picks, stt, S(stream));
  return select_run<float>(reinterpret_cast<const float*>(effects), base_logits, N, Sc, k, max_image_effect, class_index, picks, stt, S(stream));
}

// -------------------------------------------------------------------------------------------------
// per-kernel timing for bench.py
// -------------------------------------------------------------------------------------------------
int sx_profile_enable(int on) {
  Profiler& p = profiler();
  std::lock_guard<std::mutex> g(p.m);
  for (auto& r : p.recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  p.recs.clear();
  p.on = on != 0;
  return SX_OK;
}

int sx_profile_collect(double* rows, int max_rows, int* n_rows) {
  SX_REQUIRE(rows && n_rows && max_rows >= 1, "bad argument");
  Profiler& p = profiler();
  std::lock_guard<std::mutex> g(p.m);
  std::map<int, std::vector<double>> agg;  // kind -> launches, ms, flops, bytes
  for (auto& r : p.recs) {
    SX_CUDA(cudaEventSynchronize(r.b));
    float ms = 0.f;
    SX_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    auto& a = agg[r.kind];
    if (a.empty()) a.assign(4, 0.0);
    a[0] += 1; a[1] += ms; a[2] += r.flops; a[3] += r.bytes;
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  p.recs.clear();
  int n = 0;
  for (auto& kv : agg) {
    if (n >= max_rows) break;
    rows[5 * n] = kv.first;
    for (int j = 0; j < 4; ++j) rows[5 * n + 1 + j] = kv.second[j];
    ++n;
  }
  *n_rows = n;
  return SX_OK;
}

// -------------------------------------------------------------------------------------------------
// diagnostics: tcgen05 kernel vs FFMA kernel on small problems (host synchronous)
// -------------------------------------------------------------------------------------------------
static int selftest_case(int B, int Ci, int Co, int H, float* max_err) {
  const int W = H, k = 3;
  const size_t nx = (size_t)B * Ci * H * W, nw = (size_t)Co * Ci * 9, ns = (size_t)B * Ci, no = (size_t)B * Co * H * W;
  std::vector<float> hx(nx), hw(nw), hs(ns), o32(no), o16(no);
  uint32_t seed = 12345u + B * 7 + Ci * 13 + Co * 17 + H;
  auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
  auto bf = [](float v) { return __bfloat162float(__float2bfloat16_rn(v)); };
  for (auto& v : hs) v = rnd();
  // choose x so that x * (style + 1) is exactly representable in bf16 products' inputs: round after modulation
  // happens on the device for the bf16 path; keep values coarse (multiples of 1/64) so both paths see the same
  for (auto& v : hx) v = roundf(rnd() * 64.f) / 64.f;
  for (auto& v : hw) v = bf(rnd() * 0.2f);
  for (auto& v : hs) v = roundf(v * 4.f) / 4.f;  // (style+1) in {0.5, 0.75, ..., 1.5}: products stay bf16-exact
  float *dx, *dw, *ds, *do32, *do16;
  void* ws;
  const size_t wsb = sx_conv2dmod_workspace_bytes(B, Ci, Co, H, W, k, SX_PREC_FP32);
  SX_CUDA(cudaMalloc(&dx, nx * 4));
  SX_CUDA(cudaMalloc(&dw, nw * 4));
  SX_CUDA(cudaMalloc(&ds, ns * 4));
  SX_CUDA(cudaMalloc(&do32, no * 4));
  SX_CUDA(cudaMalloc(&do16, no * 4));
  SX_CUDA(cudaMalloc(&ws, wsb));
  SX_CUDA(cudaMemcpy(dx, hx.data(), nx * 4, cudaMemcpyHostToDevice));
  SX_CUDA(cudaMemcpy(dw, hw.data(), nw * 4, cudaMemcpyHostToDevice));
  SX_CUDA(cudaMemcpy(ds, hs.data(), ns * 4, cudaMemcpyHostToDevice));
  int rc = sx_conv2dmod_fwd(dx, dw, ds, do32, B, Ci, Co, H, W, k, 1, 1e-8f, SX_PREC_FP32, ws, wsb, nullptr);
  if (rc == SX_OK) rc = sx_conv2dmod_fwd(dx, dw, ds, do16, B, Ci, Co, H, W, k, 1, 1e-8f, SX_PREC_BF16, ws, wsb, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  if (rc == SX_OK && e != cudaSuccess) rc = fail(SX_ECUDA, "selftest B=%d Ci=%d Co=%d H=%d: %s", B, Ci, Co, H, cudaGetErrorString(e));
  if (rc == SX_OK) {
    cudaMemcpy(o32.data(), do32, no * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(o16.data(), do16, no * 4, cudaMemcpyDeviceToHost);
    float m = 0.f;
    for (size_t i = 0; i < no; ++i) {
      const float d = fabsf(o32[i] - o16[i]);
      if (!(d <= m)) m = d;  // NaN propagates into m
    }
    if (!(m <= *max_err)) *max_err = m;
  }
  cudaFree(dx); cudaFree(dw); cudaFree(ds); cudaFree(do32); cudaFree(do16); cudaFree(ws);
  return rc;
}

// fused-upsample conv (low-resolution input, upsample inside the kernel) vs upsample kernel + plain halo conv.
// Inputs are coarse (multiples of 1/64), so every bilinear sum is exact in fp32 and both paths round the same
// values to bf16: the two outputs must agree to the last bit.
static int selftest_ups_case(int B, int Ci, int Co, int H, float* max_err) {
  const int Hs = H / 2;
  const size_t nlow = (size_t)B * Hs * Hs * Ci, nfull = (size_t)B * H * H * Ci, nw = (size_t)Co * 9 * Ci, no = (size_t)B * H * H * Co;
  std::vector<__nv_bfloat16> hx(nlow), hw(nw), oa(no), ob(no);
  uint32_t seed = 777u + B * 7 + Ci * 13 + Co * 17 + H;
  auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
  for (auto& v : hx) v = __float2bfloat16_rn(roundf(rnd() * 64.f) / 64.f);
  for (auto& v : hw) v = __float2bfloat16_rn(rnd() * 0.2f);
  __nv_bfloat16 *dlow, *dfull, *dw, *da, *db;
  SX_CUDA(cudaMalloc(&dlow, nlow * 2));
  SX_CUDA(cudaMalloc(&dfull, nfull * 2));
  SX_CUDA(cudaMalloc(&dw, nw * 2));
  SX_CUDA(cudaMalloc(&da, no * 2));
  SX_CUDA(cudaMalloc(&db, no * 2));
  SX_CUDA(cudaMemcpy(dlow, hx.data(), nlow * 2, cudaMemcpyHostToDevice));
  SX_CUDA(cudaMemcpy(dw, hw.data(), nw * 2, cudaMemcpyHostToDevice));
  SX_CUDA(cudaMemset(da, 0xFF, no * 2));
  SX_CUDA(cudaMemset(db, 0x7F, no * 2));
  ConvEpilogue ep{};
  ep.act = 1;
  int rc = launch_upsample2x_modulate<__nv_bfloat16>(dlow, (long long)Hs * Hs * Ci, nullptr, 0, dfull, B, Hs, Hs, Ci, nullptr);
  bool handled = false;
  ep.out = da;
  if (rc == SX_OK) rc = tc::launch_conv_halo(dfull, dw, B, Ci, Co, H, H, ep, nullptr, &handled);
  if (rc == SX_OK && !handled) rc = fail(SX_EUNSUPPORTED, "selftest(ups): halo kernel does not cover Ci=%d Co=%d H=%d", Ci, Co, H);
  ep.out = db;
  if (rc == SX_OK) rc = tc::launch_conv_halo_ups(dlow, dw, B, Ci, Co, H, H, ep, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  if (rc == SX_OK && e != cudaSuccess) rc = fail(SX_ECUDA, "selftest(ups) B=%d Ci=%d Co=%d H=%d: %s", B, Ci, Co, H, cudaGetErrorString(e));
  if (rc == SX_OK) {
    cudaMemcpy(oa.data(), da, no * 2, cudaMemcpyDeviceToHost);
    cudaMemcpy(ob.data(), db, no * 2, cudaMemcpyDeviceToHost);
    float m = 0.f;
    for (size_t i = 0; i < no; ++i) {
      const float d = fabsf(__bfloat162float(oa[i]) - __bfloat162float(ob[i]));
      if (!(d <= m)) m = d;
    }
    // same MMA order in both kernels => bit-identical; when exactly one of the two runs the column-parity form (another
    // fp32 summation order) the rounded bf16 outputs may differ by an ulp
    const float allow = (tc::halo_par() & ~1) ? 0.02f : 0.f;
    if (m > allow) rc = fail(SX_ECUDA, "selftest(ups) B=%d Ci=%d Co=%d H=%d: fused-upsample conv differs from upsample + conv by %g", B, Ci, Co, H, m);
    if (!(m <= *max_err)) *max_err = m;
  }
  cudaFree(dlow); cudaFree(dfull); cudaFree(dw); cudaFree(da); cudaFree(db);
  return rc;
}

int sx_tc_selftest(float tol, float* max_err_out) {
  SX_TRY(sx_device_check());
  float max_err = 0.f;
  const int cases[][4] = {{2, 64, 64, 16}, {3, 32, 32, 8}, {5, 128, 256, 4}, {1, 64, 32, 128}, {2, 512, 512, 8}, {1, 64, 128, 32}};
  for (auto& c : cases) SX_TRY(selftest_case(c[0], c[1], c[2], c[3], &max_err));
  const int ups_cases[][4] = {{2, 64, 32, 64}, {3, 64, 32, 32}, {1, 128, 64, 32}, {2, 256, 128, 32}, {1, 64, 64, 64}};
  for (auto& c : ups_cases) SX_TRY(selftest_ups_case(c[0], c[1], c[2], c[3], &max_err));
  if (max_err_out) *max_err_out = max_err;
  if (!(max_err <= tol)) return fail(SX_ECUDA, "tcgen05 selftest: max-abs error %g > tol %g", max_err, tol);
  return SX_OK;
}

}  // extern "C"
