// generator.cuh -- the generator plan: packed weights + the launch sequence of Generator.forward
// (reference ST:794-825) from StyleSpace, with clean-prefix reuse for the AttFind sweep (NB:346-387).
//
// Per block l (Ci -> Co at H = 4 << l), all activations NHWC in T (float | bf16):
//   xin  = [upsample2x(y2 of block l-1)] * (s1 + 1)                      modulate / upsample2x_modulate
//   y1m  = lrelu(d1 * conv3x3(W1, xin) + noise1) * (s2 + 1)               conv kernel, fused epilogue
//   y2   = lrelu(d2 * conv3x3(W2, y1m) + noise2)                          conv kernel, fused epilogue
//   rgb  = conv1x1(Wrgb * (sr + 1), y2) + blur(upsample2x(rgb of l-1))    torgb kernel
// d = rsqrt(sum_i (s_i+1)^2 * sum_taps W[o,i]^2 + eps)                     demod kernel (one launch, all convs)
#pragma once

#include <vector>

#include "bandwidth.cuh"
#include "common.cuh"
#include "conv_simt.cuh"
#include "conv_tc.cuh"
#include "conv_tc_halo.cuh"

struct sx_generator {
  int num_blocks = 0, latent = 0, image_size = 0;
  int ci[SX_MAX_BLOCKS], co[SX_MAX_BLOCKS];
  int soff1[SX_MAX_BLOCKS], soff2[SX_MAX_BLOCKS], soffr[SX_MAX_BLOCKS];  // columns of the style row
  int doff1[SX_MAX_BLOCKS], doff2[SX_MAX_BLOCKS];                         // columns of the demod row
  int S = 0, style_row = 0, demod_row = 0;
  bool loaded = false;
  // device memory owned by the handle
  std::vector<void*> allocs;
  float* affine_w = nullptr;   // [style_row][latent]
  float* affine_b = nullptr;   // [style_row]
  int* col_layer = nullptr;    // [style_row] block index of each style column
  float* x0 = nullptr;         // [4][4][C0] folded initial conv, fp32
  __nv_bfloat16* x0_bf = nullptr;
  struct Conv {
    float* wpk = nullptr;            // [9][Ci][Co] fp32
    __nv_bfloat16* wbf = nullptr;    // [Co][9*Ci]  bf16
    float* wsq = nullptr;            // [Ci][Co]
    float* noise_w = nullptr;        // [Co]
    float* noise_b = nullptr;        // [Co]
  } conv[2 * SX_MAX_BLOCKS];
  float* wrgb[SX_MAX_BLOCKS];        // [3][Co]

  ~sx_generator() {
    for (void* p : allocs) cudaFree(p);
  }
};

namespace sx {

template <typename T>
inline int gen_alloc(sx_generator* g, T** out, size_t count) {
  void* p = nullptr;
  SX_CUDA(cudaMalloc(&p, count * sizeof(T) + 256));
  g->allocs.push_back(p);
  *out = reinterpret_cast<T*>(p);
  return SX_OK;
}

inline int generator_create(const int* ci, const int* co, int nb, int latent, sx_generator** out) {
  SX_REQUIRE(out && ci && co, "null argument");
  SX_REQUIRE(nb >= 1 && nb <= SX_MAX_BLOCKS, "num_blocks=%d out of range [1,%d]", nb, SX_MAX_BLOCKS);
  SX_REQUIRE(latent >= 1, "latent_dim=%d", latent);
  sx_generator* g = new sx_generator();
  g->num_blocks = nb;
  g->latent = latent;
  g->image_size = 4 << (nb - 1);
  int s = 0, d = 0;
  for (int l = 0; l < nb; ++l) {
    if (ci[l] < 1 || co[l] < 1 || (ci[l] % 4) || (co[l] % 4) || (l > 0 && ci[l] != co[l - 1])) {
      delete g;
      return fail(SX_EINVAL, "block %d: channels (%d -> %d) must be positive multiples of 4 and chain", l, ci[l], co[l]);
    }
    g->ci[l] = ci[l];
    g->co[l] = co[l];
    g->soff1[l] = s; s += ci[l];
    g->soff2[l] = s; s += co[l];
    g->doff1[l] = d; d += co[l];
    g->doff2[l] = d; d += co[l];
  }
  g->S = s;
  for (int l = 0; l < nb; ++l) { g->soffr[l] = s; s += co[l]; }
  g->style_row = s;
  g->demod_row = d;
  *out = g;
  return SX_OK;
}

inline int generator_load(sx_generator* g, const float* initial_block, const float* iw, const float* ib,
                          const sx_block_params* bp, cudaStream_t st) {
  SX_REQUIRE(g && initial_block && iw && ib && bp, "null argument");
  for (void* p : g->allocs) cudaFree(p);
  g->allocs.clear();
  g->loaded = false;
  const int nb = g->num_blocks, lat = g->latent;
  SX_TRY(gen_alloc(g, &g->affine_w, (size_t)g->style_row * lat));
  SX_TRY(gen_alloc(g, &g->affine_b, (size_t)g->style_row));
  SX_TRY(gen_alloc(g, &g->col_layer, (size_t)g->style_row));
  std::vector<int> col_layer(g->style_row);
  auto d2d = [&](void* dst, const void* src, size_t bytes) { return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st); };
  for (int l = 0; l < nb; ++l) {
    const sx_block_params& b = bp[l];
    SX_REQUIRE(b.to_style1_w && b.to_style1_b && b.to_noise1_w && b.to_noise1_b && b.conv1_w && b.to_style2_w && b.to_style2_b &&
                   b.to_noise2_w && b.to_noise2_b && b.conv2_w && b.rgb_style_w && b.rgb_style_b && b.rgb_conv_w,
               "block %d: null parameter pointer", l);
    const int ci = g->ci[l], co = g->co[l];
    SX_CUDA(d2d(g->affine_w + (size_t)g->soff1[l] * lat, b.to_style1_w, (size_t)ci * lat * 4));
    SX_CUDA(d2d(g->affine_b + g->soff1[l], b.to_style1_b, (size_t)ci * 4));
    SX_CUDA(d2d(g->affine_w + (size_t)g->soff2[l] * lat, b.to_style2_w, (size_t)co * lat * 4));
    SX_CUDA(d2d(g->affine_b + g->soff2[l], b.to_style2_b, (size_t)co * 4));
    SX_CUDA(d2d(g->affine_w + (size_t)g->soffr[l] * lat, b.rgb_style_w, (size_t)co * lat * 4));
    SX_CUDA(d2d(g->affine_b + g->soffr[l], b.rgb_style_b, (size_t)co * 4));
    for (int i = 0; i < ci; ++i) col_layer[g->soff1[l] + i] = l;
    for (int i = 0; i < co; ++i) col_layer[g->soff2[l] + i] = l;
    for (int i = 0; i < co; ++i) col_layer[g->soffr[l] + i] = l;
    for (int j = 0; j < 2; ++j) {
      sx_generator::Conv& c = g->conv[2 * l + j];
      const int cin = j == 0 ? ci : co;
      const float* W = j == 0 ? b.conv1_w : b.conv2_w;
      SX_TRY(gen_alloc(g, &c.wpk, (size_t)9 * cin * co));
      SX_TRY(gen_alloc(g, &c.wbf, (size_t)9 * cin * co));
      SX_TRY(gen_alloc(g, &c.wsq, (size_t)cin * co));
      SX_TRY(gen_alloc(g, &c.noise_w, (size_t)co));
      SX_TRY(gen_alloc(g, &c.noise_b, (size_t)co));
      pack_weights_kernel<<<ew_grid((long long)cin * co, 256), 256, 0, st>>>(W, c.wpk, c.wbf, c.wsq, co, cin, 9);
      SX_CHECK_LAUNCH();
      SX_CUDA(d2d(c.noise_w, j == 0 ? b.to_noise1_w : b.to_noise2_w, (size_t)co * 4));
      SX_CUDA(d2d(c.noise_b, j == 0 ? b.to_noise1_b : b.to_noise2_b, (size_t)co * 4));
    }
    SX_TRY(gen_alloc(g, &g->wrgb[l], (size_t)3 * co));
    SX_CUDA(d2d(g->wrgb[l], b.rgb_conv_w, (size_t)3 * co * 4));
  }
  SX_CUDA(cudaMemcpyAsync(g->col_layer, col_layer.data(), col_layer.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  const int c0 = g->ci[0];
  SX_TRY(gen_alloc(g, &g->x0, (size_t)16 * c0));
  SX_TRY(gen_alloc(g, &g->x0_bf, (size_t)16 * c0));
  initial_conv_kernel<<<(16 * c0 + 127) / 128, 128, 0, st>>>(initial_block, iw, ib, g->x0, c0);
  SX_CHECK_LAUNCH();
  convert_kernel<__nv_bfloat16><<<ew_grid(16 * c0, 256), 256, 0, st>>>(g->x0, g->x0_bf, 16 * c0);
  SX_CHECK_LAUNCH();
  SX_CUDA(cudaStreamSynchronize(st));  // col_layer (host vector) must outlive the copy
  g->loaded = true;
  return SX_OK;
}

inline int generator_styles(const sx_generator* g, const float* w, float* styles, int B, cudaStream_t st) {
  SX_REQUIRE(g && g->loaded, "generator not loaded");
  SX_REQUIRE(B >= 0, "B=%d", B);
  if (B == 0) return SX_OK;
  SX_REQUIRE(w && styles, "null argument");
  const long long warps = (long long)B * g->style_row;
  styles_affine_kernel<<<ew_grid(warps * 32, 256, 16), 256, 0, st>>>(w, g->affine_w, g->affine_b, g->col_layer, styles, B,
                                                                  g->num_blocks, g->latent, g->style_row);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

// ---- workspace layout ------------------------------------------------------------------------------
struct GenWorkspace {
  // clean-prefix cache of ONE latent (batch 1): raw (un-modulated) input of every conv + every block's rgb
  size_t cache_in[2 * SX_MAX_BLOCKS];  // byte offsets
  size_t cache_rgb[SX_MAX_BLOCKS];
  size_t dcoef, xin, y1m, y2, rgb_a, rgb_b, total;
};

inline GenWorkspace gen_workspace(const sx_generator* g, int maxB, int precision) {
  const size_t es = precision == SX_PREC_BF16 ? 2 : 4;
  GenWorkspace w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  size_t max_xin = 0, max_y = 0;
  for (int l = 0; l < g->num_blocks; ++l) {
    const size_t hw = (size_t)(4 << l) * (4 << l);
    w.cache_in[2 * l] = take(hw * g->ci[l] * es);
    w.cache_in[2 * l + 1] = take(hw * g->co[l] * es);
    w.cache_rgb[l] = take(hw * 3 * 4);
    max_xin = max_xin > hw * g->ci[l] ? max_xin : hw * g->ci[l];
    max_y = max_y > hw * g->co[l] ? max_y : hw * g->co[l];
  }
  const size_t B = maxB < 1 ? 1 : maxB;
  w.dcoef = take(B * g->demod_row * 4);
  w.xin = take(B * max_xin * es);
  w.y1m = take(B * max_y * es);
  w.y2 = take(B * max_y * es);
  const size_t img = (size_t)g->image_size * g->image_size * 3 * 4;
  w.rgb_a = take(B * img);
  w.rgb_b = take(B * img / 4);  // rgb of every block but the last is at most a quarter of the image
  w.total = off;
  return w;
}

template <typename T>
inline int run_conv(const sx_generator* g, int conv_idx, const T* x, int B, int Ci, int Co, int H, const ConvEpilogue& ep,
                    cudaStream_t st);
template <>
inline int run_conv<float>(const sx_generator* g, int conv_idx, const float* x, int B, int Ci, int Co, int H,
                           const ConvEpilogue& ep, cudaStream_t st) {
  ConvSimtParams p;
  p.x = x; p.x_bstride = (long long)H * H * Ci; p.wpk = g->conv[conv_idx].wpk;
  p.B = B; p.Ci = Ci; p.Co = Co; p.H = H; p.W = H; p.KS = 3; p.ep = ep;
  ProfScope ps(conv_idx, 2.0 * 9 * Ci * Co * (double)H * H * B, 4.0 * B * H * H * (Ci + Co), st);
  return launch_conv_simt(p, st);
}
template <>
inline int run_conv<__nv_bfloat16>(const sx_generator* g, int conv_idx, const __nv_bfloat16* x, int B, int Ci, int Co, int H,
                                   const ConvEpilogue& ep, cudaStream_t st) {
  ProfScope ps(conv_idx, 2.0 * 9 * Ci * Co * (double)H * H * B, 2.0 * B * H * H * (Ci + Co), st);
  return tc::launch_conv_bf16(x, g->conv[conv_idx].wbf, B, Ci, Co, H, H, 3, ep, st);
}

template <typename T>
inline int run_conv_ups(const sx_generator* g, int conv_idx, const T* xlow, int B, int Ci, int Co, int H, const ConvEpilogue& ep,
                        cudaStream_t st);
template <>
inline int run_conv_ups<float>(const sx_generator*, int, const float*, int, int, int, int, const ConvEpilogue&, cudaStream_t) {
  return fail(SX_EUNSUPPORTED, "internal: the fused-upsample conv is a bf16 kernel");
}
template <>
inline int run_conv_ups<__nv_bfloat16>(const sx_generator* g, int conv_idx, const __nv_bfloat16* xlow, int B, int Ci, int Co, int H,
                                       const ConvEpilogue& ep, cudaStream_t st) {
  return tc::launch_conv_halo_ups(xlow, g->conv[conv_idx].wbf, B, Ci, Co, H, H, ep, st);
}

template <typename T>
int generator_forward_t(sx_generator* g, const float* styles, const float* inoise, int noise_batch, float* rgb_out, int B,
                        int start_conv, int save_cache, uint8_t* ws, const GenWorkspace& L, cudaStream_t st) {
  const int nb = g->num_blocks, row = g->style_row;
  const int S = g->image_size;
  float* dcoef = reinterpret_cast<float*>(ws + L.dcoef);
  T* xin = reinterpret_cast<T*>(ws + L.xin);
  T* y1m = reinterpret_cast<T*>(ws + L.y1m);
  T* y2 = reinterpret_cast<T*>(ws + L.y2);
  auto cache_in = [&](int c) { return reinterpret_cast<T*>(ws + L.cache_in[c]); };
  auto cache_rgb = [&](int l) { return reinterpret_cast<float*>(ws + L.cache_rgb[l]); };
  const T* x0 = std::is_same<T, float>::value ? reinterpret_cast<const T*>(g->x0) : reinterpret_cast<const T*>(g->x0_bf);

  // demodulation coefficients of every conv that runs
  {
    DemodParams dp{};
    int max_ci = 0, max_co = 0;
    for (int l = 0; l < nb; ++l) {
      dp.conv[2 * l] = DemodConv{g->conv[2 * l].wsq, g->ci[l], g->co[l], g->soff1[l], g->doff1[l]};
      dp.conv[2 * l + 1] = DemodConv{g->conv[2 * l + 1].wsq, g->co[l], g->co[l], g->soff2[l], g->doff2[l]};
    }
    for (int c = start_conv; c < 2 * nb; ++c) {
      max_ci = max_ci > dp.conv[c].ci ? max_ci : dp.conv[c].ci;
      max_co = max_co > dp.conv[c].co ? max_co : dp.conv[c].co;
    }
    dp.first_conv = start_conv;
    dp.num_convs = 2 * nb - start_conv;
    dp.styles = styles;
    dp.style_stride = row;
    dp.dcoef = dcoef;
    dp.dcoef_stride = g->demod_row;
    dp.eps = 1e-8f;
    ProfScope ps(35, 0, 0, st);
    SX_TRY(launch_demod(dp, B, max_ci, max_co, st));
  }

  const int start_block = start_conv / 2;
  // ToRGB is fused into the conv2 epilogue when the kernel has the whole channel row of a pixel in one thread
  // and one sample per M tile: bf16 path, Co <= 256, H >= 16.
  static const bool fuse_off = getenv("SX_DISABLE_RGB_FUSION") != nullptr;
  auto rgb_fused = [&](int l) { return !fuse_off && std::is_same<T, __nv_bfloat16>::value && g->co[l] <= 256 && (4 << l) >= 16; };
  // The bilinear 2x upsample in front of conv1 is fused into the conv kernel (its A-operand producers) where the
  // halo kernel covers the shape: the previous block's conv2 then stores its output already modulated by this block's
  // style1 (its ToRGB must be fused, because nothing else sees the un-modulated tensor) and the upsampled tensor is
  // never materialised.  The clean-prefix cache of such a conv1 holds the LOW-RESOLUTION raw tensor.
  static const bool ups_off = getenv("SX_DISABLE_UPS_FUSION") != nullptr;
  auto ups_fused = [&](int l) {
    return !ups_off && std::is_same<T, __nv_bfloat16>::value && l >= 1 && rgb_fused(l - 1) &&
           tc::halo_ups_supported(g->ci[l], g->co[l], 4 << l, 4 << l);
  };
  const float* prev_rgb = nullptr;  // rgb of the previous block, [Bp,3,H/2,H/2]
  long long prev_bstride = 0;
  if (start_block > 0) {
    prev_rgb = cache_rgb(start_block - 1);
    prev_bstride = 0;  // one cached sample broadcast to the batch
  }
  float* rgb_ping = reinterpret_cast<float*>(ws + L.rgb_a);
  float* rgb_pong = reinterpret_cast<float*>(ws + L.rgb_b);

  for (int l = start_block; l < nb; ++l) {
    const int H = 4 << l, ci = g->ci[l], co = g->co[l];
    const long long HW = (long long)H * H;
    const bool run_conv1 = start_conv <= 2 * l;
    ConvEpilogue ep{};
    ep.noise = inoise; ep.noise_batch = noise_batch; ep.noise_size = S; ep.act = 1; ep.dcoef_stride = g->demod_row;
    const bool ups = ups_fused(l);
    if (run_conv1 && ups) {
      // ---- conv1 with the upsample fused: its input is the low-resolution tensor modulated by (style1 + 1)
      const long long HWs = HW / 4;
      const T* xlow = y2;   // written (modulated) by the previous block's conv2 epilogue
      if (l == start_block) {
        ProfScope ps(32, 0, (double)sizeof(T) * B * HWs * ci, st);
        SX_TRY(launch_modulate<T>(cache_in(2 * l), 0, styles + g->soff1[l], row, xin, B, HWs, ci, st));
        xlow = xin;
      }
      ep.dcoef = dcoef + g->doff1[l];
      ep.noise_w = g->conv[2 * l].noise_w; ep.noise_b = g->conv[2 * l].noise_b;
      ep.next_style = styles + g->soff2[l]; ep.next_style_stride = row;
      ep.out = y1m; ep.out_nchw_f32 = 0;
      ep.out_raw = save_cache ? cache_in(2 * l + 1) : nullptr;
      {
        ProfScope ps(2 * l, 2.0 * 9 * ci * co * (double)HW * B, 2.0 * B * (HWs * ci + HW * co), st);
        SX_TRY(run_conv_ups<T>(g, 2 * l, xlow, B, ci, co, H, ep, st));
      }
    } else if (run_conv1) {
      // ---- input of conv1, modulated by (style1 + 1)
      if (l == start_block) {
        const T* src = (l == 0) ? x0 : cache_in(2 * l);
        if (l == 0 && save_cache) {
          SX_CUDA(cudaMemcpyAsync(cache_in(0), x0, (size_t)HW * ci * sizeof(T), cudaMemcpyDeviceToDevice, st));
        }
        ProfScope ps(32, 0, (double)sizeof(T) * B * HW * ci, st);
        SX_TRY(launch_modulate<T>(src, 0, styles + g->soff1[l], row, xin, B, HW, ci, st));
      } else if (save_cache) {
        // keep the raw upsampled tensor for the sweep, then modulate it
        SX_TRY(launch_upsample2x_modulate<T>(y2, HW / 4 * ci, nullptr, 0, cache_in(2 * l), B, H / 2, H / 2, ci, st));
        SX_TRY(launch_modulate<T>(cache_in(2 * l), HW * ci, styles + g->soff1[l], row, xin, B, HW, ci, st));
      } else {
        ProfScope ps(33, 0, (double)sizeof(T) * B * HW * ci * 1.25, st);
        SX_TRY(launch_upsample2x_modulate<T>(y2, HW / 4 * ci, styles + g->soff1[l], row, xin, B, H / 2, H / 2, ci, st));
      }
      // ---- conv1: y1m = lrelu(d1 * conv + noise1) * (style2 + 1)
      ep.dcoef = dcoef + g->doff1[l];
      ep.noise_w = g->conv[2 * l].noise_w; ep.noise_b = g->conv[2 * l].noise_b;
      ep.next_style = styles + g->soff2[l]; ep.next_style_stride = row;
      ep.out = y1m; ep.out_nchw_f32 = 0;
      ep.out_raw = save_cache ? cache_in(2 * l + 1) : nullptr;
      SX_TRY(run_conv<T>(g, 2 * l, xin, B, ci, co, H, ep, st));
    } else {
      // sweep starts at conv2 of this block: its raw input comes from the cache
      ProfScope ps(32, 0, (double)sizeof(T) * B * HW * co, st);
      SX_TRY(launch_modulate<T>(cache_in(2 * l + 1), 0, styles + g->soff2[l], row, y1m, B, HW, co, st));
    }
    // ---- conv2: y2 = lrelu(d2 * conv + noise2)   [+ fused ToRGB on the tcgen05 path]
    float* rgb_dst = (l == nb - 1) ? rgb_out : ((l & 1) ? rgb_pong : rgb_ping);
    if (l != nb - 1 && (size_t)B * 3 * HW * 4 > ((l & 1) ? (L.total - L.rgb_b) : (L.rgb_b - L.rgb_a)))
      return fail(SX_ENOMEM, "internal: rgb scratch too small");
    const bool fuse_rgb = rgb_fused(l);
    const bool next_ups = l + 1 < nb && ups_fused(l + 1);
    ep.dcoef = dcoef + g->doff2[l];
    ep.noise_w = g->conv[2 * l + 1].noise_w; ep.noise_b = g->conv[2 * l + 1].noise_b;
    ep.next_style = nullptr; ep.next_style_stride = 0;
    ep.out = y2; ep.out_raw = nullptr;
    if (next_ups) {   // y2 leaves this conv modulated for the next block's conv1; the raw copy feeds the prefix cache
      ep.next_style = styles + g->soff1[l + 1]; ep.next_style_stride = row;
      ep.out_raw = save_cache ? cache_in(2 * l + 2) : nullptr;
    }
    if (fuse_rgb) {
      if (prev_rgb) {
        ProfScope ps(37, 0, (double)B * HW * 15, st);
        SX_TRY(launch_rgb_prev_up_blur(prev_rgb, prev_bstride, rgb_dst, B, H, H, st));
      }
      ep.rgb_style = styles + g->soffr[l]; ep.rgb_style_stride = row; ep.rgb_w = g->wrgb[l];
      ep.rgb_out = rgb_dst; ep.rgb_accumulate = prev_rgb ? 1 : 0;
      if (l == nb - 1) ep.out = nullptr;   // nothing reads the last feature map
    }
    SX_TRY(run_conv<T>(g, 2 * l + 1, y1m, B, co, co, H, ep, st));
    ep.rgb_style = nullptr; ep.rgb_out = nullptr;
    // ---- ToRGB (+ upsample/blur of the previous rgb), standalone where it is not fused
    if (!fuse_rgb) {
      ProfScope ps(34, 2.0 * 3 * co * (double)HW * B, (double)B * HW * (sizeof(T) * co + 12 + (prev_rgb ? 3 : 0)), st);
      SX_TRY(launch_torgb<T>(y2, styles + g->soffr[l], row, g->wrgb[l], prev_rgb, prev_bstride, rgb_dst, B, H, H, co, st));
    }
    if (save_cache)
      SX_CUDA(cudaMemcpyAsync(cache_rgb(l), rgb_dst, (size_t)3 * HW * 4, cudaMemcpyDeviceToDevice, st));
    prev_rgb = rgb_dst;
    prev_bstride = 3 * HW;
  }
  return SX_OK;
}

inline int generator_forward(sx_generator* g, const float* styles, const float* inoise, int noise_batch, float* rgb_out, int B,
                             int start_conv, int save_cache, int precision, void* workspace, size_t ws_bytes, cudaStream_t st) {
  SX_REQUIRE(g && g->loaded, "generator not loaded");
  SX_REQUIRE(B >= 0, "B=%d", B);
  if (B == 0) return SX_OK;
  SX_REQUIRE(styles && inoise && rgb_out && workspace, "null argument");
  SX_REQUIRE(noise_batch == 1 || noise_batch == B, "noise batch %d must be 1 or B=%d", noise_batch, B);
  SX_REQUIRE(start_conv >= 0 && start_conv < 2 * g->num_blocks, "start_conv=%d out of range", start_conv);
  SX_REQUIRE(!(save_cache && (B != 1 || start_conv != 0)), "save_cache needs B == 1 and start_conv == 0");
  SX_REQUIRE(precision == SX_PREC_FP32 || precision == SX_PREC_BF16, "precision=%d", precision);
  SX_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  if (B == 0) return SX_OK;
  const GenWorkspace L = gen_workspace(g, B, precision);
  if (ws_bytes < L.total) return fail(SX_ENOMEM, "workspace %zu bytes < required %zu (B=%d)", ws_bytes, L.total, B);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  if (precision == SX_PREC_BF16) {
    for (int l = 0; l < g->num_blocks; ++l)
      if (!tc::tc_shape_supported(g->ci[l], g->co[l], 4 << l, 4 << l, 3) || !tc::tc_shape_supported(g->co[l], g->co[l], 4 << l, 4 << l, 3))
        return fail(SX_EUNSUPPORTED, "bf16 tcgen05 path: block %d (%d -> %d) has an unsupported channel count", l, g->ci[l], g->co[l]);
    return generator_forward_t<__nv_bfloat16>(g, styles, inoise, noise_batch, rgb_out, B, start_conv, save_cache, ws, L, st);
  }
  return generator_forward_t<float>(g, styles, inoise, noise_batch, rgb_out, B, start_conv, save_cache, ws, L, st);
}

}  // namespace sx
