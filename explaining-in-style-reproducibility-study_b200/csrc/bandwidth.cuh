// bandwidth.cuh -- the HBM-bound kernels of the generator: modulate, bilinear 2x upsample, ToRGB
// (+ fused upsample/blur of the previous rgb), noise/leaky-ReLU, layout changes, and the small
// per-batch tables (style affine, demodulation coefficients).  All loads/stores are 16-byte vectors on
// the contiguous (channel for NHWC, x for NCHW) axis; grids are sized in multiples of the SM count.
#pragma once

#include "common.cuh"

namespace sx {

// =================================================================================================
// NHWC plan kernels (T = float | __nv_bfloat16)
// =================================================================================================

// out[b, p, c] = in[b|0, p, c] * (style[b, c] + 1)          (style may be null: plain broadcast copy)
// grid = (chunks of one sample, B): all index arithmetic is 32-bit (64-bit div/mod made the first version
// of these kernels instruction-bound at 19 % of HBM peak).
template <typename T>
__global__ void __launch_bounds__(256) modulate_kernel(const T* __restrict__ in, long long in_bstride,
                                                       const float* __restrict__ style, int style_stride,
                                                       T* __restrict__ out, unsigned per_b /* P * C/V */, unsigned cv) {
  constexpr int V = Elem<T>::kVec;
  using vec_t = typename Elem<T>::vec_t;
  const unsigned b = blockIdx.y;
  const vec_t* src = reinterpret_cast<const vec_t*>(in + (long long)b * in_bstride);
  vec_t* dst = reinterpret_cast<vec_t*>(out) + (size_t)b * per_b;
  const float* s = style ? style + (long long)b * style_stride : nullptr;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < per_b; i += gridDim.x * blockDim.x) {
    vec_t v = __ldg(src + i);
    if (s) {
      const unsigned c = (i % cv) * V;
      float f[V];
      unpack(v, f);
      const float4* sp = reinterpret_cast<const float4*>(s + c);
#pragma unroll
      for (int j = 0; j < V / 4; ++j) {
        const float4 m = __ldg(sp + j);
        f[4 * j] *= m.x + 1.f; f[4 * j + 1] *= m.y + 1.f; f[4 * j + 2] *= m.z + 1.f; f[4 * j + 3] *= m.w + 1.f;
      }
      pack(f, v);
    }
    dst[i] = v;
  }
}

template <typename T>
int launch_modulate(const T* in, long long in_bstride, const float* style, int style_stride, T* out, int B,
                    long long P, int C, cudaStream_t st) {
  SX_REQUIRE(C % Elem<T>::kVec == 0, "modulate: C=%d not a multiple of %d", C, Elem<T>::kVec);
  const long long per_b = P * (C / Elem<T>::kVec);
  SX_REQUIRE(per_b < (1ll << 31), "modulate: sample too large");
  if (per_b == 0 || B == 0) return SX_OK;
  const int gx = (int)((per_b + 255) / 256 < 4 * num_sms() ? (per_b + 255) / 256 : 4 * num_sms());
  dim3 grid(gx, B);
  modulate_kernel<T><<<grid, 256, 0, st>>>(in, in_bstride, style, style_stride, out, (unsigned)per_b, (unsigned)(C / Elem<T>::kVec));
  SX_CHECK_LAUNCH();
  return SX_OK;
}

// torch's upsample_bilinear2d(align_corners=False, scale 2) source index computation
__device__ __forceinline__ void bilinear_src(int o, int n_in, int& i0, int& i1, float& l0, float& l1) {
  float src = (o + 0.5f) * 0.5f - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = (int)src;
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  l1 = src - i0;
  l0 = 1.f - l1;
}

constexpr int UPS_ROWS = 4;  // INPUT rows per CTA of the upsample kernel (8 output rows)

// out[b, 2H, 2W, C] = upsample2x(in[b|0]) * (style[b, c] + 1)      (ST:679,693-694 fused with the
// activation-side modulation of the next conv1).  H, W are the INPUT sizes.
// One thread = one 2x2 output quad x V channels.  Interior quads read their 3x3 input neighbourhood once (9 vector
// loads for 4 outputs) and use the constant bilinear weights (0.25, 0.75) in torch's evaluation order; border quads
// (clamped source indices) take the generic per-pixel path.  The first version (1 output per thread, 4 loads, per-pixel
// weight computation) was instruction-bound at 28 % of the HBM roofline.
template <typename T>
__device__ __forceinline__ void ups_store(T* __restrict__ dst, const float* __restrict__ o, const float* __restrict__ m, bool mod) {
  constexpr int V = Elem<T>::kVec;
  using vec_t = typename Elem<T>::vec_t;
  float r[V];
#pragma unroll
  for (int k = 0; k < V; ++k) r[k] = mod ? to_f(from_f<T>(o[k])) * m[k] : o[k];   // round to storage type, then modulate
  vec_t v;
  pack(r, v);
  *reinterpret_cast<vec_t*>(dst) = v;
}

template <typename T>
__global__ void __launch_bounds__(256) upsample2x_modulate_kernel(const T* __restrict__ in, long long in_bstride,
                                                                  const float* __restrict__ style, int style_stride,
                                                                  T* __restrict__ out, int H, int W, int C) {
  constexpr int V = Elem<T>::kVec;
  using vec_t = typename Elem<T>::vec_t;
  const unsigned cv = C / V;
  const int OW = 2 * W;
  const int b = blockIdx.z;
  const T* src = in + (long long)b * in_bstride;
  const float* s = style ? style + (long long)b * style_stride : nullptr;
  T* dstb = out + (size_t)b * 4 * H * W * C;
  const unsigned rowq = W * cv;
  const int i_end = min((int)(blockIdx.y + 1) * UPS_ROWS, H);
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < rowq; t += gridDim.x * blockDim.x) {
    const unsigned j = t / cv;
    const unsigned c = (t - j * cv) * V;
    float m[V];
    if (s) {
#pragma unroll
      for (int k = 0; k < V; k += 4) {
        const float4 mv = __ldg(reinterpret_cast<const float4*>(s + c + k));
        m[k] = mv.x + 1.f; m[k + 1] = mv.y + 1.f; m[k + 2] = mv.z + 1.f; m[k + 3] = mv.w + 1.f;
      }
    }
    const bool jint = j >= 1 && (int)j < W - 1;
    for (int i = blockIdx.y * UPS_ROWS; i < i_end; ++i) {
      T* d0 = dstb + ((size_t)(2 * i) * OW + 2 * j) * C + c;   // (2i, 2j); +C: (2i, 2j+1); +OW*C: next row
      if (jint && i >= 1 && i < H - 1) {
        float he0[V], ho0[V], he1[V], ho1[V], o[V];
        {
          const T* r = src + ((size_t)(i - 1) * W + (j - 1)) * C + c;
          float a[V], bb[V], cc[V];
          unpack(__ldg(reinterpret_cast<const vec_t*>(r)), a);
          unpack(__ldg(reinterpret_cast<const vec_t*>(r + C)), bb);
          unpack(__ldg(reinterpret_cast<const vec_t*>(r + 2 * C)), cc);
#pragma unroll
          for (int k = 0; k < V; ++k) { he0[k] = 0.25f * a[k] + 0.75f * bb[k]; ho0[k] = 0.75f * bb[k] + 0.25f * cc[k]; }
        }
        {
          const T* r = src + ((size_t)i * W + (j - 1)) * C + c;
          float a[V], bb[V], cc[V];
          unpack(__ldg(reinterpret_cast<const vec_t*>(r)), a);
          unpack(__ldg(reinterpret_cast<const vec_t*>(r + C)), bb);
          unpack(__ldg(reinterpret_cast<const vec_t*>(r + 2 * C)), cc);
#pragma unroll
          for (int k = 0; k < V; ++k) { he1[k] = 0.25f * a[k] + 0.75f * bb[k]; ho1[k] = 0.75f * bb[k] + 0.25f * cc[k]; }
        }
#pragma unroll
        for (int k = 0; k < V; ++k) o[k] = 0.25f * he0[k] + 0.75f * he1[k];
        ups_store<T>(d0, o, m, s != nullptr);
#pragma unroll
        for (int k = 0; k < V; ++k) o[k] = 0.25f * ho0[k] + 0.75f * ho1[k];
        ups_store<T>(d0 + C, o, m, s != nullptr);
        {
          const T* r = src + ((size_t)(i + 1) * W + (j - 1)) * C + c;
          float a[V], bb[V], cc[V];
          unpack(__ldg(reinterpret_cast<const vec_t*>(r)), a);
          unpack(__ldg(reinterpret_cast<const vec_t*>(r + C)), bb);
          unpack(__ldg(reinterpret_cast<const vec_t*>(r + 2 * C)), cc);
#pragma unroll
          for (int k = 0; k < V; ++k) { he0[k] = 0.25f * a[k] + 0.75f * bb[k]; ho0[k] = 0.75f * bb[k] + 0.25f * cc[k]; }   // row i+1
        }
#pragma unroll
        for (int k = 0; k < V; ++k) o[k] = 0.75f * he1[k] + 0.25f * he0[k];
        ups_store<T>(d0 + (size_t)OW * C, o, m, s != nullptr);
#pragma unroll
        for (int k = 0; k < V; ++k) o[k] = 0.75f * ho1[k] + 0.25f * ho0[k];
        ups_store<T>(d0 + (size_t)OW * C + C, o, m, s != nullptr);
      } else {
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          const int oy = 2 * i + (q >> 1), ox = 2 * (int)j + (q & 1);
          int y0, y1, x0, x1;
          float ly0, ly1, lx0, lx1;
          bilinear_src(oy, H, y0, y1, ly0, ly1);
          bilinear_src(ox, W, x0, x1, lx0, lx1);
          float f00[V], f01[V], f10[V], f11[V], o[V];
          unpack(__ldg(reinterpret_cast<const vec_t*>(src + ((size_t)y0 * W + x0) * C + c)), f00);
          unpack(__ldg(reinterpret_cast<const vec_t*>(src + ((size_t)y0 * W + x1) * C + c)), f01);
          unpack(__ldg(reinterpret_cast<const vec_t*>(src + ((size_t)y1 * W + x0) * C + c)), f10);
          unpack(__ldg(reinterpret_cast<const vec_t*>(src + ((size_t)y1 * W + x1) * C + c)), f11);
#pragma unroll
          for (int k = 0; k < V; ++k) o[k] = ly0 * (lx0 * f00[k] + lx1 * f01[k]) + ly1 * (lx0 * f10[k] + lx1 * f11[k]);
          ups_store<T>(dstb + ((size_t)oy * OW + ox) * C + c, o, m, s != nullptr);
        }
      }
    }
  }
}

template <typename T>
int launch_upsample2x_modulate(const T* in, long long in_bstride, const float* style, int style_stride, T* out, int B,
                               int H, int W, int C, cudaStream_t st) {
  SX_REQUIRE(C % Elem<T>::kVec == 0, "upsample: C=%d not a multiple of %d", C, Elem<T>::kVec);
  if (B == 0) return SX_OK;
  const int row = W * (C / Elem<T>::kVec);
  const int threads = row >= 256 ? 256 : (row >= 128 ? 128 : 64);
  dim3 grid((row + threads - 1) / threads, (H + UPS_ROWS - 1) / UPS_ROWS, B);
  SX_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "upsample: grid too large");
  upsample2x_modulate_kernel<T><<<grid, threads, 0, st>>>(in, in_bstride, style, style_stride, out, H, W, C);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

// value of blur3x3_reflect(upsample2x(p + q)) at (y, x) of the (2h x 2w) image; p, q: fp32 planes [h, w] (q may be null).
__device__ __forceinline__ float ld2(const float* __restrict__ p, const float* __restrict__ q, int i) {
  float v = __ldg(p + i);
  if (q) v += __ldg(q + i);
  return v;
}
__device__ __forceinline__ float up2(const float* __restrict__ p, const float* __restrict__ q, int h, int w, int y, int x) {
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  bilinear_src(y, h, y0, y1, ly0, ly1);
  bilinear_src(x, w, x0, x1, lx0, lx1);
  return ly0 * (lx0 * ld2(p, q, y0 * w + x0) + lx1 * ld2(p, q, y0 * w + x1)) +
         ly1 * (lx0 * ld2(p, q, y1 * w + x0) + lx1 * ld2(p, q, y1 * w + x1));
}
__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }
__device__ __forceinline__ float up2_blur(const float* __restrict__ p, const float* __restrict__ q, int h, int w, int y, int x) {
  const int H = 2 * h, W = 2 * w;
  float acc = 0.f;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = reflect1(y + dy, H);
    const float ky = dy == 0 ? 2.f : 1.f;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = reflect1(x + dx, W);
      const float kx = dx == 0 ? 2.f : 1.f;
      acc += (ky * kx * (1.f / 16.f)) * up2(p, q, h, w, yy, xx);
    }
  }
  return acc;
}

// RGBBlock (ST:618-629) for one block, fused with the previous block's Upsample+Blur tail:
//   rgb[b, c, y, x] = sum_o y2[b, y, x, o] * (sr[b, o] + 1) * Wrgb[c, o]  +  blur(up2x(prev[b|0]))[c, y, x]
// y2 NHWC (T), prev [Bp, 3, H/2, W/2] fp32 planar (null for the first block), rgb [B,3,H,W] fp32 planar.
// One CTA = one (<=16 x 16)-pixel tile of one sample:
//   phase A  the (TH+2) x (TW+2) window of upsample2x(prev) the blur needs (reflect border) goes to shared memory
//            once per tile (~4 bilinear samples per thread instead of 108 per pixel), then a 9-tap blur from smem;
//   phase B  the 1x1 conv with LPP = min(32, Co/V) lanes per pixel: every lane loads 16 B of consecutive channels, so
//            a warp load instruction reads whole contiguous pixel rows (coalesced), partial sums are shuffle-reduced;
//   phase C  planar fp32 stores, 16 consecutive x per row.
template <typename T>
__global__ void __launch_bounds__(256) torgb_kernel(const T* __restrict__ y2, const float* __restrict__ rgb_style,
                                                    int style_stride, const float* __restrict__ wrgb,
                                                    const float* __restrict__ prev, long long prev_bstride,
                                                    float* __restrict__ rgb, int H, int W, int Co, int TH, int TW) {
  extern __shared__ float smem_f[];
  constexpr int V = Elem<T>::kVec;
  using vec_t = typename Elem<T>::vec_t;
  float* s_w = smem_f;                        // [3][Co]
  float* s_res = s_w + 3 * Co;                // [3][256]
  float* s_up = s_res + 3 * 256;              // [3][(TH+2)*(TW+2)]
  const int b = blockIdx.y;
  const int tiles_x = W / TW;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int y0 = ty * TH, x0 = tx * TW;
  const int npix = TH * TW;
  const int tid = threadIdx.x;
  for (int i = tid; y2 != nullptr && i < 3 * Co; i += blockDim.x) {
    const int o = i % Co;
    s_w[i] = (__ldg(rgb_style + (long long)b * style_stride + o) + 1.f) * __ldg(wrgb + i);
  }
  const int UW = TW + 2, UH = TH + 2;
  if (prev) {
    const float* pp = prev + (long long)b * prev_bstride;
    const int h = H / 2, w = W / 2;
    for (int i = tid; i < 3 * UH * UW; i += blockDim.x) {
      const int c = i / (UH * UW);
      const int r = i - c * UH * UW;
      const int uy = r / UW, ux = r - uy * UW;
      s_up[i] = up2(pp + c * h * w, nullptr, h, w, reflect1(y0 - 1 + uy, H), reflect1(x0 - 1 + ux, W));
    }
  }
  __syncthreads();
  // ---- phase B
  const int lpp = Co / V < 32 ? Co / V : 32;         // lanes per pixel (power of two: Co is 32..512, V 4|8)
  const int ppw = 32 / lpp;                           // pixels per warp iteration
  const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int sub = lane / lpp, cl = lane - sub * lpp;  // pixel slot in the warp, channel lane
  if (y2 == nullptr) {   // "previous rgb only" mode: rgb = blur(up2x(prev)); the 1x1 conv is fused into the conv2 epilogue
    for (int pl = tid; pl < npix; pl += blockDim.x) { s_res[pl] = 0.f; s_res[256 + pl] = 0.f; s_res[512 + pl] = 0.f; }
  }
  for (int p0 = warp * ppw; y2 != nullptr && p0 < npix; p0 += nwarps * ppw) {
    const int pl = p0 + sub;                          // pixel index in the tile
    const bool live = pl < npix;
    const int py = pl / TW, px = pl - py * TW;
    const T* src = y2 + (((long long)b * H + y0 + py) * W + x0 + px) * Co;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int o = cl * V; live && o < Co; o += lpp * V) {
      float f[V];
      unpack(__ldg(reinterpret_cast<const vec_t*>(src + o)), f);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        a0 = fmaf(f[j], s_w[o + j], a0);
        a1 = fmaf(f[j], s_w[Co + o + j], a1);
        a2 = fmaf(f[j], s_w[2 * Co + o + j], a2);
      }
    }
    for (int off = lpp >> 1; off > 0; off >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, off);
      a1 += __shfl_xor_sync(0xffffffffu, a1, off);
      a2 += __shfl_xor_sync(0xffffffffu, a2, off);
    }
    if (cl == 0 && live) {
      s_res[pl] = a0;
      s_res[256 + pl] = a1;
      s_res[512 + pl] = a2;
    }
  }
  __syncthreads();
  // ---- phase C
  const long long HW = (long long)H * W;
  for (int pl = tid; pl < npix; pl += blockDim.x) {
    const int py = pl / TW, px = pl - py * TW;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = s_res[c * 256 + pl];
      if (prev) {
        const float* u = s_up + c * UH * UW + py * UW + px;   // window origin = (y0-1, x0-1)
        v += (1.f / 16.f) * (u[0] + u[2] + u[2 * UW] + u[2 * UW + 2]) + (2.f / 16.f) * (u[1] + u[UW] + u[UW + 2] + u[2 * UW + 1]) +
             (4.f / 16.f) * u[UW + 1];
      }
      rgb[((long long)b * 3 + c) * HW + (long long)(y0 + py) * W + x0 + px] = v;
    }
  }
}

template <typename T>
int launch_torgb(const T* y2, const float* rgb_style, int style_stride, const float* wrgb, const float* prev,
                 long long prev_bstride, float* rgb, int B, int H, int W, int Co, cudaStream_t st) {
  constexpr int V = Elem<T>::kVec;
  SX_REQUIRE(Co % V == 0, "torgb: Co=%d not a multiple of %d", Co, V);
  const int lpp = Co / V < 32 ? Co / V : 32;
  SX_REQUIRE((lpp & (lpp - 1)) == 0, "torgb: Co/%d = %d must be a power of two (or >= 32)", V, Co / V);
  SX_REQUIRE(Co / V < 32 || Co % (32 * V) == 0, "torgb: Co=%d must be a multiple of %d", Co, 32 * V);
  if (B == 0) return SX_OK;
  // 8 x 8 pixel tiles, 8 warps: a warp walks its pixels one after the other (load -> 48 FMAs -> shuffle reduce), so the
  // loads in flight per SM are the number of resident warps.  With 16 x 16 tiles the 16 px / 512-channel launch was 256 CTAs
  // of 32 serial pixels per warp: 835 GB/s (ncu r02b: 13 % of the copy bandwidth, latency-bound).  SX_TORGB_TILE=16: the old tiling.
  static const int tile = getenv("SX_TORGB_TILE") ? atoi(getenv("SX_TORGB_TILE")) : 8;
  const int TW = W < tile ? W : tile, TH = H < tile ? H : tile;
  SX_REQUIRE(W % TW == 0 && H % TH == 0, "torgb: H, W must be multiples of the tile");
  const int npix = TH * TW;
  const int threads = npix >= 64 ? 256 : 128;    // >= 128: the 3*Co-entry weight table fill is a chain of dependent loads
  dim3 grid((W / TW) * (H / TH), B);
  const size_t smem = (size_t)(3 * Co + 3 * 256 + 3 * (TH + 2) * (TW + 2)) * sizeof(float);
  torgb_kernel<T><<<grid, threads, smem, st>>>(y2, rgb_style, style_stride, wrgb, prev, prev_bstride, rgb, H, W, Co, TH, TW);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

// rgb[b] = blur(upsample2x(prev[b|0]))  -- pre-fills the rgb buffer a fused-ToRGB conv2 epilogue accumulates into.
// Interior pixels use the closed form of blur([1,2,1]/4) o bilinear-upsample: per axis a 3-tap filter on the low-res
// neighbourhood {i-1, i, i+1} with weights (5,10,1)/16 for even and (1,10,5)/16 for odd output coordinates
// (9 loads + 9 FMAs per channel); pixels whose neighbourhood touches the border (index clamping of the upsample,
// reflect padding of the blur) take the generic two-step evaluation.
// One thread = one 2x2 output quad (all 3 channels): the four pixels share the low-res neighbourhood, so 27 loads
// serve 12 outputs.  Border quads use the same 3x3 evaluation with the composite weights of their row / column
// computed on the fly (blur taps reflected, bilinear sources clamped) -- the first versions sent them through the
// generic 36-sample path, and with one border lane in half of the warps that path was 90 % of the kernel's time.
__device__ __forceinline__ void blur_up_weights(int o, int n_in, float* wgt /*[3]: low-res o/2-1, o/2, o/2+1*/) {
  const int base = (o >> 1) - 1;
  wgt[0] = wgt[1] = wgt[2] = 0.f;
#pragma unroll
  for (int t = -1; t <= 1; ++t) {
    const int u = reflect1(o + t, 2 * n_in);
    int i0, i1;
    float l0, l1;
    bilinear_src(u, n_in, i0, i1, l0, l1);
    const float kb = t == 0 ? 0.5f : 0.25f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (i0 - base == k) wgt[k] += kb * l0;
      if (i1 - base == k) wgt[k] += kb * l1;
    }
  }
}

__global__ void __launch_bounds__(128) rgb_prev_up_blur_kernel(const float* __restrict__ prev, long long prev_bstride,
                                                               float* __restrict__ rgb, int h, int w) {
  const int H = 2 * h, W = 2 * w;
  const int b = blockIdx.z;
  const int i = blockIdx.y;
  const float* pp = prev + (long long)b * prev_bstride;
  float* dst = rgb + (long long)b * 3 * H * W;
  float cye[3] = {0.3125f, 0.625f, 0.0625f}, cyo[3] = {0.0625f, 0.625f, 0.3125f};   // rows 2i / 2i+1 over low-res rows i-1, i, i+1
  if (i < 1 || i > h - 2) {
    blur_up_weights(2 * i, h, cye);
    blur_up_weights(2 * i + 1, h, cyo);
  }
  const int r0 = max(i - 1, 0), r2 = min(i + 1, h - 1);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < w; j += gridDim.x * blockDim.x) {
    float cxe[3] = {0.3125f, 0.625f, 0.0625f}, cxo[3] = {0.0625f, 0.625f, 0.3125f};
    if (j < 1 || j > w - 2) {
      blur_up_weights(2 * j, w, cxe);
      blur_up_weights(2 * j + 1, w, cxo);
    }
    const int c0 = max(j - 1, 0), c2 = min(j + 1, w - 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* p = pp + (size_t)c * h * w;
      const int rows[3] = {r0, i, r2};
      float re[3], ro[3];   // even / odd output column of the three low-res rows
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float* q = p + (size_t)rows[r] * w;
        const float p0 = __ldg(q + c0), p1 = __ldg(q + j), p2 = __ldg(q + c2);
        re[r] = cxe[0] * p0 + cxe[1] * p1 + cxe[2] * p2;
        ro[r] = cxo[0] * p0 + cxo[1] * p1 + cxo[2] * p2;
      }
      float* d = dst + ((size_t)c * H + 2 * i) * W + 2 * j;
      *reinterpret_cast<float2*>(d) = make_float2(cye[0] * re[0] + cye[1] * re[1] + cye[2] * re[2],
                                                  cye[0] * ro[0] + cye[1] * ro[1] + cye[2] * ro[2]);
      *reinterpret_cast<float2*>(d + W) = make_float2(cyo[0] * re[0] + cyo[1] * re[1] + cyo[2] * re[2],
                                                      cyo[0] * ro[0] + cyo[1] * ro[1] + cyo[2] * ro[2]);
    }
  }
}

// The same arithmetic (horizontal 3-tap per low-res row, then vertical 3-tap, identical operation order) with one
// thread = a 2x2 block of low-res pixels = a 4x4 block of outputs per channel: 48 loads serve 48 outputs (the quad form:
// 27 for 12) and every output row leaves as one 16-byte store (a warp writes 512 contiguous bytes per row).
__global__ void __launch_bounds__(128) rgb_prev_up_blur4_kernel(const float* __restrict__ prev, long long prev_bstride,
                                                                float* __restrict__ rgb, int h, int w) {
  const int H = 2 * h, W = 2 * w;
  const int b = blockIdx.z;
  const int i0 = 2 * blockIdx.y;
  const int jp = blockIdx.x * blockDim.x + threadIdx.x;
  if (2 * jp >= w) return;
  const int j0 = 2 * jp;
  const float* pp = prev + (long long)b * prev_bstride;
  float* dst = rgb + (long long)b * 3 * H * W;
  // weights of output row 2*i0 + r over the low-res rows (i-1, i, i+1), i = i0 + (r >> 1); likewise for columns
  float cy[4][3], cx[4][3];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const bool odd = r & 1;
    cy[r][0] = cx[r][0] = odd ? 0.0625f : 0.3125f;
    cy[r][1] = cx[r][1] = 0.625f;
    cy[r][2] = cx[r][2] = odd ? 0.3125f : 0.0625f;
  }
  if (i0 < 1 || i0 + 1 > h - 2) {
#pragma unroll
    for (int r = 0; r < 4; ++r) blur_up_weights(2 * i0 + r, h, cy[r]);
  }
  if (j0 < 1 || j0 + 1 > w - 2) {
#pragma unroll
    for (int r = 0; r < 4; ++r) blur_up_weights(2 * j0 + r, w, cx[r]);
  }
  int rr[4], cc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    rr[k] = min(max(i0 - 1 + k, 0), h - 1);
    cc[k] = min(max(j0 - 1 + k, 0), w - 1);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* p = pp + (size_t)c * h * w;
    float hx[4][4];   // [low-res row k][output column q]
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float* q = p + (size_t)rr[k] * w;
      const float v0 = __ldg(q + cc[0]), v1 = __ldg(q + cc[1]), v2 = __ldg(q + cc[2]), v3 = __ldg(q + cc[3]);
      hx[k][0] = cx[0][0] * v0 + cx[0][1] * v1 + cx[0][2] * v2;
      hx[k][1] = cx[1][0] * v0 + cx[1][1] * v1 + cx[1][2] * v2;
      hx[k][2] = cx[2][0] * v1 + cx[2][1] * v2 + cx[2][2] * v3;
      hx[k][3] = cx[3][0] * v1 + cx[3][1] * v2 + cx[3][2] * v3;
    }
    float* d = dst + ((size_t)c * H + 2 * i0) * W + 2 * j0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int k0 = r >> 1;   // output rows 0,1 read low-res rows k = 0..2; rows 2,3 read k = 1..3
      float4 o;
      o.x = cy[r][0] * hx[k0][0] + cy[r][1] * hx[k0 + 1][0] + cy[r][2] * hx[k0 + 2][0];
      o.y = cy[r][0] * hx[k0][1] + cy[r][1] * hx[k0 + 1][1] + cy[r][2] * hx[k0 + 2][1];
      o.z = cy[r][0] * hx[k0][2] + cy[r][1] * hx[k0 + 1][2] + cy[r][2] * hx[k0 + 2][2];
      o.w = cy[r][0] * hx[k0][3] + cy[r][1] * hx[k0 + 1][3] + cy[r][2] * hx[k0 + 2][3];
      *reinterpret_cast<float4*>(d + (size_t)r * W) = o;
    }
  }
}

inline int launch_rgb_prev_up_blur(const float* prev, long long prev_bstride, float* rgb, int B, int H, int W, cudaStream_t st) {
  if (B == 0) return SX_OK;
  SX_REQUIRE(H % 2 == 0 && W % 2 == 0 && H >= 4 && W >= 4, "rgb_prev: bad size %dx%d", H, W);
  static const bool quad_only = getenv("SX_RGB_PREV_QUAD") != nullptr;   // A/B: the one-quad-per-thread form
  if (!quad_only && H % 4 == 0 && W % 4 == 0 && H >= 16 && W >= 16 && (reinterpret_cast<uintptr_t>(rgb) & 15) == 0) {
    const int wp = W / 4;                                   // low-res column pairs
    const int threads = wp >= 128 ? 128 : (wp >= 64 ? 64 : 32);
    dim3 grid((wp + threads - 1) / threads, H / 4, B);
    SX_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "rgb_prev: grid too large");
    rgb_prev_up_blur4_kernel<<<grid, threads, 0, st>>>(prev, prev_bstride, rgb, H / 2, W / 2);
    SX_CHECK_LAUNCH();
    return SX_OK;
  }
  const int w = W / 2;
  const int threads = w >= 128 ? 128 : (w >= 64 ? 64 : 32);
  dim3 grid((w + threads - 1) / threads, H / 2, B);
  SX_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "rgb_prev: grid too large");
  rgb_prev_up_blur_kernel<<<grid, threads, 0, st>>>(prev, prev_bstride, rgb, H / 2, W / 2);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

// =================================================================================================
// per-batch tables
// =================================================================================================

// K5: styles[b, j] = bias[j] + sum_k w[b, layer(j), k] * A[j, k]   (one warp per output, shuffle reduction)
__global__ void __launch_bounds__(256) styles_affine_kernel(const float* __restrict__ w, const float* __restrict__ A,
                                                            const float* __restrict__ bias,
                                                            const int* __restrict__ col_layer, float* __restrict__ out,
                                                            int B, int L, int latent, int row) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long t = warp; t < (long long)B * row; t += nwarps) {
    const int b = (int)(t / row), j = (int)(t - (long long)b * row);
    const float* wv = w + ((long long)b * L + __ldg(col_layer + j)) * latent;
    const float* a = A + (long long)j * latent;
    float acc = 0.f;
    for (int k = lane; k < latent; k += 32) acc = fmaf(__ldg(wv + k), __ldg(a + k), acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) out[t] = acc + __ldg(bias + j);
  }
}

// demodulation coefficients of every conv from `first_conv` on, for the whole batch:
//   d[b, doff_c + o] = rsqrt( sum_i (style[b, soff_c + i] + 1)^2 * wsq_c[i][o] + eps )      (ST:654-656)
struct DemodConv {
  const float* wsq;  // [Ci][Co]
  int ci, co, soff, doff;
};
struct DemodParams {
  DemodConv conv[2 * SX_MAX_BLOCKS];
  int first_conv, num_convs;
  const float* styles;
  int style_stride;
  float* dcoef;
  int dcoef_stride;
  float eps;
};
// One CTA = 128 output channels x DEMOD_BPC samples: every wsq element fetched from L2 serves DEMOD_BPC samples (the
// one-sample-per-CTA form re-read the conv's whole wsq matrix per sample -- 2 GB of L2 traffic per 256-sample launch,
// 90 us).  Per sample the accumulation order over i is unchanged, so the coefficients are bit-identical.
constexpr int DEMOD_BPC = 8;
__global__ void __launch_bounds__(128) demod_kernel(DemodParams p, int B) {
  const DemodConv cv = p.conv[p.first_conv + blockIdx.z];
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const int b0 = blockIdx.y * DEMOD_BPC;
  if (blockIdx.x * blockDim.x >= cv.co) return;   // whole CTA out of range for this (narrower) conv
  extern __shared__ float s_s2[];  // [ci][DEMOD_BPC]: (style+1)^2 of the CTA's samples for this conv
  for (int t = threadIdx.x; t < cv.ci * DEMOD_BPC; t += blockDim.x) {
    const int k = t / cv.ci, i = t - k * cv.ci;
    const int b = b0 + k;
    float v = 0.f;
    if (b < B) {
      v = __ldg(p.styles + (long long)b * p.style_stride + cv.soff + i) + 1.f;
      v *= v;
    }
    s_s2[i * DEMOD_BPC + k] = v;
  }
  __syncthreads();
  if (o >= cv.co) return;
  float acc[DEMOD_BPC];
#pragma unroll
  for (int k = 0; k < DEMOD_BPC; ++k) acc[k] = 0.f;
  // 16 independent L2 loads in flight per thread: with the loads issued one per iteration this loop ran at L2 latency
#pragma unroll 16
  for (int i = 0; i < cv.ci; ++i) {
    const float w = __ldg(cv.wsq + (long long)i * cv.co + o);
    const float4 sa = *reinterpret_cast<const float4*>(s_s2 + i * DEMOD_BPC);
    const float4 sb = *reinterpret_cast<const float4*>(s_s2 + i * DEMOD_BPC + 4);
    acc[0] = fmaf(sa.x, w, acc[0]); acc[1] = fmaf(sa.y, w, acc[1]); acc[2] = fmaf(sa.z, w, acc[2]); acc[3] = fmaf(sa.w, w, acc[3]);
    acc[4] = fmaf(sb.x, w, acc[4]); acc[5] = fmaf(sb.y, w, acc[5]); acc[6] = fmaf(sb.z, w, acc[6]); acc[7] = fmaf(sb.w, w, acc[7]);
  }
#pragma unroll
  for (int k = 0; k < DEMOD_BPC; ++k)
    if (b0 + k < B) p.dcoef[(long long)(b0 + k) * p.dcoef_stride + cv.doff + o] = rsqrtf(acc[k] + p.eps);
}

inline int launch_demod(const DemodParams& dp, int B, int max_ci, int max_co, cudaStream_t st) {
  if (B == 0 || dp.num_convs == 0) return SX_OK;
  const size_t smem = (size_t)max_ci * DEMOD_BPC * sizeof(float);
  SX_REQUIRE(smem <= 200 * 1024, "demod: Ci=%d too large for the shared-memory style table", max_ci);
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    SX_CUDA(cudaFuncSetAttribute(demod_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid((max_co + 127) / 128, (B + DEMOD_BPC - 1) / DEMOD_BPC, dp.num_convs);
  SX_REQUIRE(grid.y <= 65535, "demod: batch too large");
  demod_kernel<<<grid, 128, smem, st>>>(dp, B);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

// =================================================================================================
// weight packing (sx_generator_load) -- one-off
// =================================================================================================
// W[Co][Ci][k][k] fp32  ->  wpk[tap][Ci][Co] fp32,  wbf[Co][tap*Ci + ci] bf16 (K-major),  wsq[Ci][Co] = sum_tap W^2
__global__ void pack_weights_kernel(const float* __restrict__ W, float* __restrict__ wpk, __nv_bfloat16* __restrict__ wbf,
                                    float* __restrict__ wsq, int Co, int Ci, int taps) {
  const long long total = (long long)Co * Ci;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i / Ci), c = (int)(i - (long long)o * Ci);
    float sq = 0.f;
    for (int t = 0; t < taps; ++t) {
      const float v = W[i * taps + t];
      sq = fmaf(v, v, sq);
      if (wpk) wpk[((long long)t * Ci + c) * Co + o] = v;
      if (wbf) wbf[(long long)o * taps * Ci + (long long)t * Ci + c] = __float2bfloat16_rn(v);
    }
    if (wsq) wsq[(long long)c * Co + o] = sq;
  }
}

// x0[y][x][o] = bias[o] + sum_{i,dy,dx} W[o][i][dy][dx] * blk[i][y+dy-1][x+dx-1]   (4x4, zero padded; ST:802,806)
__global__ void initial_conv_kernel(const float* __restrict__ blk, const float* __restrict__ W, const float* __restrict__ bias,
                                    float* __restrict__ x0, int C) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 16 * C) return;
  const int o = idx % C, pix = idx / C, y = pix / 4, x = pix % 4;
  float acc = 0.f;
  for (int i = 0; i < C; ++i)
    for (int dy = 0; dy < 3; ++dy)
      for (int dx = 0; dx < 3; ++dx) {
        const int yy = y + dy - 1, xx = x + dx - 1;
        if (yy < 0 || yy > 3 || xx < 0 || xx > 3) continue;
        acc = fmaf(W[(((long long)o * C + i) * 3 + dy) * 3 + dx], blk[(i * 4 + yy) * 4 + xx], acc);
      }
  x0[idx] = acc + bias[o];
}

template <typename T>
__global__ void convert_kernel(const float* __restrict__ in, T* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = from_f<T>(in[i]);
}

// =================================================================================================
// reference-layout (NCHW fp32) ops behind the nn.Module forwards
// =================================================================================================

// x[B,C,H,W] fp32 * (style[b,c]+1)  ->  NHWC T   (32x32 shared-memory transpose, coalesced both ways)
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_modulate_kernel(const float* __restrict__ x, const float* __restrict__ style,
                                                                    T* __restrict__ out, int C, int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, p = p0 + tx;
    float v = 0.f;
    if (c < C && p < HW) {
      v = x[((long long)b * C + c) * HW + p];
      if (style) v *= __ldg(style + (long long)b * C + c) + 1.f;
    }
    tile[j][tx] = v;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int p = p0 + j, c = c0 + tx;
    if (c < C && p < HW) out[((long long)b * HW + p) * C + c] = from_f<T>(tile[tx][j]);
  }
}

__global__ void __launch_bounds__(256) upsample2x_nchw_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                              long long planes, int H, int W) {
  const int OH = 2 * H, OW = 2 * W;
  const long long total = planes * OH * OW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % OW);
    const long long r = i / OW;
    const int oy = (int)(r % OH);
    const long long pl = r / OH;
    out[i] = up2(x + pl * H * W, nullptr, H, W, oy, ox);
  }
}

__global__ void __launch_bounds__(256) blur_nchw_kernel(const float* __restrict__ x, float* __restrict__ out, long long planes,
                                                        int H, int W) {
  const long long total = planes * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W);
    const long long r = i / W;
    const int yy = (int)(r % H);
    const float* p = x + (r / H) * H * W;
    float acc = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int y2 = reflect1(yy + dy, H);
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int x2 = reflect1(xx + dx, W);
        acc += ((dy == 0 ? 2.f : 1.f) * (dx == 0 ? 2.f : 1.f) * (1.f / 16.f)) * __ldg(p + y2 * W + x2);
      }
    }
    out[i] = acc;
  }
}

__global__ void __launch_bounds__(256) noise_lrelu_nchw_kernel(const float* __restrict__ x, const float* __restrict__ inoise,
                                                               const float* __restrict__ nw, const float* __restrict__ nb,
                                                               float* __restrict__ out, int B, int C, int H, int W,
                                                               int noise_batch, int S) {
  const long long total = (long long)B * C * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W);
    long long r = i / W;
    const int yy = (int)(r % H);
    r /= H;
    const int c = (int)(r % C);
    const int b = (int)(r / C);
    const float nz = __ldg(inoise + (long long)(noise_batch == 1 ? 0 : b) * S * S + (long long)xx * S + yy);
    out[i] = lrelu02(x[i] + (nz * __ldg(nw + c) + __ldg(nb + c)));
  }
}

// RGBBlock tail (ST:623-627): out = rgb + prev                       (no upsample)
__global__ void __launch_bounds__(256) rgb_tail_nchw_kernel(const float* __restrict__ rgb, const float* __restrict__ prev,
                                                            float* __restrict__ sum, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    sum[i] = rgb[i] + (prev ? prev[i] : 0.f);
}
//                            out = blur(upsample2x(rgb + prev))         (fused, [planes,h,w] -> [planes,2h,2w])
__global__ void __launch_bounds__(256) rgb_tail_up_blur_nchw_kernel(const float* __restrict__ rgb, const float* __restrict__ prev,
                                                                    float* __restrict__ out, long long planes, int h, int w) {
  const int H = 2 * h, W = 2 * w;
  const long long total = planes * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W);
    const long long r = i / W;
    const int yy = (int)(r % H);
    const long long off = (r / H) * h * w;
    out[i] = up2_blur(rgb + off, prev ? prev + off : nullptr, h, w, yy, xx);
  }
}

// nn.Linear: one warp per output element
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ x, const float* __restrict__ Wt,
                                                     const float* __restrict__ bias, float* __restrict__ out, int B, int in_f,
                                                     int out_f) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long t = warp; t < (long long)B * out_f; t += nwarps) {
    const int b = (int)(t / out_f), j = (int)(t - (long long)b * out_f);
    float acc = 0.f;
    for (int k = lane; k < in_f; k += 32) acc = fmaf(__ldg(x + (long long)b * in_f + k), __ldg(Wt + (long long)j * in_f + k), acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) out[t] = acc + (bias ? __ldg(bias + j) : 0.f);
  }
}

}  // namespace sx
