// preprocess.cuh -- the tensor branch of the classifier wrappers' preprocessing as ONE kernel:
//   torchvision.transforms.functional.resize(images, [OH, OW])   (bilinear, antialias=True: torch's
//   _upsample_bilinear2d_aa)  ->  Normalize(mean, std)  ->  cast  ->  channels_last
// (reference resnet_classifier.py:60-68).  In PyTorch this is five passes over the batch (antialiased resize, sub, div,
// dtype cast, layout change: ~0.65 ms per 128 images at 256px, 9 % of the AttFind step); here it is one read of the
// image and one write of the network input.  The network itself stays PyTorch.
//
// Arithmetic follows ATen's CUDA kernel (aten/src/ATen/native/cuda/UpSampleBilinear2d.cu, upsample_gen2d_aa_out_frame):
//   scale = in / out;  support = scale >= 1 ? scale : 1;  invscale = scale >= 1 ? 1 / scale : 1
//   center = scale * (i + 0.5);  xmin = max(int(center - support + 0.5), 0);  xsize = min(int(center + support + 0.5), in) - xmin
//   w_j = tri((j + xmin - center + 0.5) * invscale), normalised by their sum;   tri(x) = max(0, 1 - |x|)
//   out = sum_y wy[y] * (sum_x wx[x] * src[ymin + y][xmin + x])               (fp32 accumulation in this order)
#pragma once

#include "common.cuh"

namespace sx {

constexpr int AA_MAX_TAPS = 12;  // support <= 5.5 input pixels per side (scale <= 5.5)

template <int TAPS>
__device__ __forceinline__ void aa_weights(int i, float scale, float support, float invscale, int in_size, int& xmin, int& xsize,
                                           float* w) {
  const float center = scale * (i + 0.5f);
  xmin = max((int)(center - support + 0.5f), 0);
  xsize = min((int)(center + support + 0.5f), in_size) - xmin;
  xsize = xsize < TAPS ? xsize : TAPS;
  float total = 0.f;
#pragma unroll
  for (int j = 0; j < TAPS; ++j) {
    float x = (j + xmin - center + 0.5f) * invscale;
    x = fabsf(x);
    const float v = (j < xsize && x < 1.f) ? 1.f - x : 0.f;
    w[j] = v;
    total += v;
  }
  if (total != 0.f) {
#pragma unroll
    for (int j = 0; j < TAPS; ++j) w[j] /= total;
  }
}

struct Norm3 {
  float mean[3], std[3];
  int on;
};

// in [B,3,IH,IW] fp32 NCHW  ->  out [B,OH,OW,3] (T) = a channels_last [B,3,OH,OW] tensor
template <typename T, int TAPS>
__global__ void __launch_bounds__(256) resize_aa_normalize_kernel(const float* __restrict__ in, T* __restrict__ out, int B, int IH, int IW,
                                                                  int OH, int OW, Norm3 nm) {
  const float sy = (float)IH / (float)OH, sx_ = (float)IW / (float)OW;
  const float sup_y = sy >= 1.f ? sy : 1.f, sup_x = sx_ >= 1.f ? sx_ : 1.f;
  const float inv_y = sy >= 1.f ? 1.f / sy : 1.f, inv_x = sx_ >= 1.f ? 1.f / sx_ : 1.f;
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y, b = blockIdx.z;
  if (ox >= OW) return;
  float wx[TAPS], wy[TAPS];
  int xmin, xsize, ymin, ysize;
  aa_weights<TAPS>(ox, sx_, sup_x, inv_x, IW, xmin, xsize, wx);
  aa_weights<TAPS>(oy, sy, sup_y, inv_y, IH, ymin, ysize, wy);
  T* dst = out + (((size_t)b * OH + oy) * OW + ox) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* plane = in + ((size_t)b * 3 + c) * IH * IW;
    float acc = 0.f;
#pragma unroll
    for (int y = 0; y < TAPS; ++y) {
      if (y < ysize) {
        const float* row = plane + (size_t)(ymin + y) * IW + xmin;
        float t = __ldg(row) * wx[0];
#pragma unroll
        for (int x = 1; x < TAPS; ++x)
          if (x < xsize) t += __ldg(row + x) * wx[x];
        acc = y == 0 ? t * wy[0] : acc + t * wy[y];
      }
    }
    if (nm.on) acc = (acc - nm.mean[c]) / nm.std[c];
    dst[c] = from_f<T>(acc);
  }
}

// Space-to-depth variant for the ResNet stem: the 7x7 / stride-2 / pad-3 convolution on 3 channels (which cuDNN runs on
// a legacy mma.sync kernel at ~17 TFLOP/s) equals a 4x4 / stride-1 / pad-0 convolution on the 2x2 space-to-depth image
// with 12 (padded to 16) channels, padded by 2 blocks before and 1 block after (classifiers.py: stem_s2d weights).
//   out[b, Y, X, (dy*2+dx)*3 + c] = resized[b, c, 2(Y-2)+dy, 2(X-2)+dx]   (0 outside the image; channels 12..15 = 0)
// out is the memory of a channels_last [B, 16, OH/2+3, OW/2+3] tensor.  One thread = one (Y, X) block = 32 B of bf16.
template <typename T, int TAPS>
__global__ void __launch_bounds__(128) resize_aa_normalize_s2d_kernel(const float* __restrict__ in, T* __restrict__ out, int B, int IH,
                                                                      int IW, int OH, int OW, Norm3 nm) {
  const float sy = (float)IH / (float)OH, sx_ = (float)IW / (float)OW;
  const float sup_y = sy >= 1.f ? sy : 1.f, sup_x = sx_ >= 1.f ? sx_ : 1.f;
  const float inv_y = sy >= 1.f ? 1.f / sy : 1.f, inv_x = sx_ >= 1.f ? 1.f / sx_ : 1.f;
  const int PH = OH / 2 + 3, PW = OW / 2 + 3;
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int Y = blockIdx.y, b = blockIdx.z;
  if (X >= PW) return;
  float v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = 0.f;
  const int by = Y - 2, bx = X - 2;
  if (by >= 0 && by < OH / 2 && bx >= 0 && bx < OW / 2) {
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      float wy[TAPS];
      int ymin, ysize;
      aa_weights<TAPS>(2 * by + dy, sy, sup_y, inv_y, IH, ymin, ysize, wy);
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float wx[TAPS];
        int xmin, xsize;
        aa_weights<TAPS>(2 * bx + dx, sx_, sup_x, inv_x, IW, xmin, xsize, wx);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float* plane = in + ((size_t)b * 3 + c) * IH * IW;
          float acc = 0.f;
#pragma unroll
          for (int y = 0; y < TAPS; ++y) {
            if (y < ysize) {
              const float* row = plane + (size_t)(ymin + y) * IW + xmin;
              float t = __ldg(row) * wx[0];
#pragma unroll
              for (int x = 1; x < TAPS; ++x)
                if (x < xsize) t += __ldg(row + x) * wx[x];
              acc = y == 0 ? t * wy[0] : acc + t * wy[y];
            }
          }
          if (nm.on) acc = (acc - nm.mean[c]) / nm.std[c];
          v[(dy * 2 + dx) * 3 + c] = to_f(from_f<T>(acc));
        }
      }
    }
  }
  constexpr int V = Elem<T>::kVec;
  using vec_t = typename Elem<T>::vec_t;
  vec_t* dst = reinterpret_cast<vec_t*>(out + (((size_t)b * PH + Y) * PW + X) * 16);
#pragma unroll
  for (int k = 0; k < 16 / V; ++k) {
    vec_t pk;
    pack(v + k * V, pk);
    dst[k] = pk;
  }
}

// Separable form of the kernel above (4-tap case: scale < 1.5, the 256 -> 224 resize of the AttFind step).  The direct kernel
// evaluates the horizontal filter of an input row once per OUTPUT row that reads it (~2.6x redundant at 256 -> 224) and was
// instruction / L1-latency bound (ncu: 0.96 TB/s, 4.7 % of the step).  Here one CTA owns S2D_TBY block rows (2*S2D_TBY output
// rows) of one sample: pass 1 filters every input row the tile needs horizontally into shared memory, t[r][c][ox] -- each
// value computed once, global reads coalesced along the row -- and pass 2 applies the vertical filter out of shared memory
// and writes the space-to-depth blocks.  Same fp32 operations in the same order as the direct kernel: bit-identical
// (test_resize_s2d_separable_equals_direct).
constexpr int S2D_TBY = 8;
constexpr int S2D_THREADS = 512;

template <typename T>
__global__ void __launch_bounds__(S2D_THREADS) resize_aa_normalize_s2d_sep_kernel(const float* __restrict__ in, T* __restrict__ out, int B,
                                                                          int IH, int IW, int OH, int OW, Norm3 nm, int max_rows) {
  constexpr int TAPS = 4;
  extern __shared__ float smem_f[];
  float* s_wx = smem_f;                          // [OW][4]
  int* s_xmin = reinterpret_cast<int*>(s_wx + 4 * OW);   // [OW]
  int* s_xsize = s_xmin + OW;                    // [OW]
  float* s_wy = reinterpret_cast<float*>(s_xsize + OW);  // [2*S2D_TBY][4]
  int* s_ymin = reinterpret_cast<int*>(s_wy + 4 * 2 * S2D_TBY);   // [2*S2D_TBY]
  int* s_ysize = s_ymin + 2 * S2D_TBY;           // [2*S2D_TBY]
  float* s_t = reinterpret_cast<float*>(s_ysize + 2 * S2D_TBY);   // [max_rows][3][OW]

  const float sy = (float)IH / (float)OH, sx_ = (float)IW / (float)OW;
  const float sup_y = sy >= 1.f ? sy : 1.f, sup_x = sx_ >= 1.f ? sx_ : 1.f;
  const float inv_y = sy >= 1.f ? 1.f / sy : 1.f, inv_x = sx_ >= 1.f ? 1.f / sx_ : 1.f;
  const int PH = OH / 2 + 3, PW = OW / 2 + 3;
  const int b = blockIdx.y;
  const int Y0 = blockIdx.x * S2D_TBY;
  // output rows of this tile: block row Y holds image rows 2(Y-2), 2(Y-2)+1
  const int oy_lo = max(2 * (Y0 - 2), 0), oy_hi = min(2 * (Y0 + S2D_TBY - 2), OH);   // [oy_lo, oy_hi)

  for (int ox = threadIdx.x; ox < OW; ox += blockDim.x) {
    float w[TAPS];
    int xmin, xsize;
    aa_weights<TAPS>(ox, sx_, sup_x, inv_x, IW, xmin, xsize, w);
    s_xmin[ox] = xmin;
    s_xsize[ox] = xsize;
#pragma unroll
    for (int j = 0; j < TAPS; ++j) s_wx[4 * ox + j] = w[j];
  }
  for (int k = threadIdx.x; k < oy_hi - oy_lo; k += blockDim.x) {
    float w[TAPS];
    int ymin, ysize;
    aa_weights<TAPS>(oy_lo + k, sy, sup_y, inv_y, IH, ymin, ysize, w);
    s_ymin[k] = ymin;
    s_ysize[k] = ysize;
#pragma unroll
    for (int j = 0; j < TAPS; ++j) s_wy[4 * k + j] = w[j];
  }
  __syncthreads();
  int r0 = 0, r1 = 0;                                        // input rows [r0, r1) feed this tile
  if (oy_hi > oy_lo) {
    r0 = s_ymin[0];
    r1 = s_ymin[oy_hi - oy_lo - 1] + s_ysize[oy_hi - oy_lo - 1];
  }
  const int rows = r1 - r0;   // <= max_rows (host bound)

  // pass 1: t[r][c][ox] = sum_x wx[ox][x] * src[c][r0 + r][xmin + x].  One thread = one output column of FOUR (row, channel)
  // lines, fully unrolled so its 16 global loads are in flight together (with one line per iteration the pass ran at
  // DRAM latency: 3x slower than the direct kernel).
  const int nrc = rows * 3;
  for (int i = threadIdx.x; i < ((nrc + 3) / 4) * OW; i += blockDim.x) {
    const int ox = i % OW;
    const int g4 = (i / OW) * 4;
    const int xmin = s_xmin[ox], xsize = s_xsize[ox];
    const float4 w4 = *reinterpret_cast<const float4*>(s_wx + 4 * ox);
    const float wx[TAPS] = {w4.x, w4.y, w4.z, w4.w};
    float src[4][TAPS];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int rc = min(g4 + q, nrc - 1);
      const int c = rc % 3, r = rc / 3;
      const float* row = in + (((size_t)b * 3 + c) * IH + (r0 + r)) * IW + xmin;
#pragma unroll
      for (int x = 0; x < TAPS; ++x) src[q][x] = x < xsize ? __ldg(row + x) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float t = src[q][0] * wx[0];
#pragma unroll
      for (int x = 1; x < TAPS; ++x)
        if (x < xsize) t += src[q][x] * wx[x];
      if (g4 + q < nrc) s_t[(g4 + q) * OW + ox] = t;
    }
  }
  __syncthreads();

  // pass 2: one thread = one (Y, X) block = 16 channels (4 pixels x 3 + 4 zero), like the direct kernel
  constexpr int V = Elem<T>::kVec;
  using vec_t = typename Elem<T>::vec_t;
  const int tile_rows = min(S2D_TBY, PH - Y0);
  for (int i = threadIdx.x; i < tile_rows * PW; i += blockDim.x) {
    const int X = i % PW, Y = Y0 + i / PW;
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = 0.f;
    const int by = Y - 2, bx = X - 2;
    if (by >= 0 && by < OH / 2 && bx >= 0 && bx < OW / 2) {
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const int k = 2 * by + dy - oy_lo;
        const int ymin = s_ymin[k], ysize = s_ysize[k];
        const float4 w4 = *reinterpret_cast<const float4*>(s_wy + 4 * k);
        const float wy[TAPS] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int ox = 2 * bx + dx;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float* col = s_t + ((ymin - r0) * 3 + c) * OW + ox;
            float acc = 0.f;
#pragma unroll
            for (int y = 0; y < TAPS; ++y) {
              if (y < ysize) {
                const float t = col[y * 3 * OW];
                acc = y == 0 ? t * wy[0] : acc + t * wy[y];
              }
            }
            if (nm.on) acc = (acc - nm.mean[c]) / nm.std[c];
            v[(dy * 2 + dx) * 3 + c] = to_f(from_f<T>(acc));
          }
        }
      }
    }
    vec_t* dst = reinterpret_cast<vec_t*>(out + (((size_t)b * PH + Y) * PW + X) * 16);
#pragma unroll
    for (int k = 0; k < 16 / V; ++k) {
      vec_t pk;
      pack(v + k * V, pk);
      dst[k] = pk;
    }
  }
}

// ---- max_pool2d(kernel 3, stride 2, padding 1) on a channels_last bf16 / fp32 tensor (torchvision ResNet stem pool) -----------
// ATen's max_pool_forward_nhwc ran at ~0.7 TB/s on the [B,64,112,112] stem output (9 % of the AttFind step).  One thread =
// one output pixel x one 16-byte channel group: up to 9 vector loads (the 2.25x re-read is served by L1/L2), a packed max,
// one vector store.  Padding never wins (-inf), NaNs propagate like torch's kernel; the result is bit-identical.
__device__ __forceinline__ uint4 vmax(const uint4& a, const uint4& b, __nv_bfloat16) {
  uint4 r;
  const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* z = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) z[i] = __hmax2_nan(x[i], y[i]);
  return r;
}
__device__ __forceinline__ float4 vmax(const float4& a, const float4& b, float) {
  auto m = [](float p, float q) { return (p != p || q != q) ? __int_as_float(0x7fffffff) : fmaxf(p, q); };
  return make_float4(m(a.x, b.x), m(a.y, b.y), m(a.z, b.z), m(a.w, b.w));
}

template <typename T>
__global__ void __launch_bounds__(256) maxpool3x3s2_nhwc_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int H, int W, int C,
                                                                int OH, int OW) {
  constexpr int V = Elem<T>::kVec;
  using vec_t = typename Elem<T>::vec_t;
  const unsigned groups = (unsigned)(C / V);
  const unsigned total = (unsigned)B * OH * OW * groups;   // < 2^31 (checked by the launcher): 32-bit index arithmetic
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned g = i % groups;
    unsigned t = i / groups;
    const int ox = (int)(t % OW);
    t /= OW;
    const int oy = (int)(t % OH);
    const int b = (int)(t / OH);
    const int y0 = max(2 * oy - 1, 0), y1 = min(2 * oy + 1, H - 1);
    const int x0 = max(2 * ox - 1, 0), x1 = min(2 * ox + 1, W - 1);
    const vec_t* src = reinterpret_cast<const vec_t*>(in) + (unsigned)(b * H * W) * groups + g;
    vec_t m = __ldg(src + (unsigned)(y0 * W + x0) * groups);
    for (int y = y0; y <= y1; ++y)
      for (int x = x0; x <= x1; ++x) m = vmax(m, __ldg(src + (unsigned)(y * W + x) * groups), T());
    reinterpret_cast<vec_t*>(out)[i] = m;
  }
}

template <typename T>
int launch_maxpool3x3s2_nhwc(const T* in, T* out, int B, int H, int W, int C, cudaStream_t st) {
  if (B == 0) return SX_OK;
  SX_REQUIRE(C % Elem<T>::kVec == 0, "maxpool: channels (%d) must be a multiple of %d", C, Elem<T>::kVec);
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  const long long total = (long long)B * OH * OW * (C / Elem<T>::kVec);
  SX_REQUIRE(total < (1ll << 31) && (long long)B * H * W * (C / Elem<T>::kVec) < (1ll << 31), "maxpool: tensor too large for 32-bit indexing");
  maxpool3x3s2_nhwc_kernel<T><<<ew_grid(total, 256, 16), 256, 0, st>>>(in, out, B, H, W, C, OH, OW);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

template <typename T>
int launch_resize_aa_normalize_s2d(const float* in, T* out, int B, int IH, int IW, int OH, int OW, const Norm3& nm, cudaStream_t st) {
  if (B == 0) return SX_OK;
  SX_REQUIRE(OH % 2 == 0 && OW % 2 == 0, "resize(s2d): output size %dx%d must be even", OH, OW);
  const float sy = (float)IH / OH, sxx = (float)IW / OW;
  const float sup = fmaxf(sy >= 1.f ? sy : 1.f, sxx >= 1.f ? sxx : 1.f);
  const int taps = (int)(2.f * sup) + 2;
  SX_REQUIRE(taps <= AA_MAX_TAPS, "resize: scale factor %.2f too large (max %d taps)", sup, AA_MAX_TAPS);
  const int PH = OH / 2 + 3, PW = OW / 2 + 3;
  if (taps <= 4 && B <= 65535 && !getenv("SX_RESIZE_DIRECT")) {
    // separable kernel: rows of horizontally filtered input one tile needs (2*S2D_TBY output rows + the filter support)
    const int max_rows = (int)ceilf(2 * S2D_TBY * sy + 2.f * (sy >= 1.f ? sy : 1.f)) + 3;
    const size_t smem = (size_t)(4 * OW + 2 * OW + 6 * 2 * S2D_TBY) * 4 + (size_t)max_rows * 3 * OW * 4;
    if (smem <= 200 * 1024) {
      auto kern = resize_aa_normalize_s2d_sep_kernel<T>;
      static size_t configured = 0;  // per instantiation
      if (smem > configured) {
        SX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
      }
      dim3 sgrid((PH + S2D_TBY - 1) / S2D_TBY, B);
      kern<<<sgrid, S2D_THREADS, smem, st>>>(in, out, B, IH, IW, OH, OW, nm, max_rows);
      SX_CHECK_LAUNCH();
      return SX_OK;
    }
  }
  dim3 grid((PW + 127) / 128, PH, B);
  SX_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "resize: grid too large");
  if (taps <= 4)
    resize_aa_normalize_s2d_kernel<T, 4><<<grid, 128, 0, st>>>(in, out, B, IH, IW, OH, OW, nm);
  else
    resize_aa_normalize_s2d_kernel<T, AA_MAX_TAPS><<<grid, 128, 0, st>>>(in, out, B, IH, IW, OH, OW, nm);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

template <typename T>
int launch_resize_aa_normalize(const float* in, T* out, int B, int IH, int IW, int OH, int OW, const Norm3& nm, cudaStream_t st) {
  if (B == 0) return SX_OK;
  const float sy = (float)IH / OH, sxx = (float)IW / OW;
  const float sup = fmaxf(sy >= 1.f ? sy : 1.f, sxx >= 1.f ? sxx : 1.f);
  const int taps = (int)(2.f * sup) + 2;
  SX_REQUIRE(taps <= AA_MAX_TAPS, "resize: scale factor %.2f too large (max %d taps)", sup, AA_MAX_TAPS);
  dim3 grid((OW + 255) / 256, OH, B);
  SX_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "resize: grid too large");
  if (taps <= 4)
    resize_aa_normalize_kernel<T, 4><<<grid, 256, 0, st>>>(in, out, B, IH, IW, OH, OW, nm);
  else
    resize_aa_normalize_kernel<T, AA_MAX_TAPS><<<grid, 256, 0, st>>>(in, out, B, IH, IW, OH, OW, nm);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

}  // namespace sx
