// conv_bwd.cuh -- first-order backward of Conv2DMod.forward (reference ST:647-667).
//
// Forward:  m = style + 1;  xm = x * m[b,i];  z = conv(W, xm);  out = z * d[b,o],
//   d[b,o] = rsqrt(sum_i m[b,i]^2 * wsq[i,o] + eps),  wsq[i,o] = sum_taps W[o,i,t]^2        (d == 1 without demod)
// Backward for an upstream gradient g = dL/dout:
//   gz      = g * d[b,o]
//   gxm     = conv(W^T flipped, gz)                         dgrad: the SAME implicit-GEMM kernels, roles of Ci/Co swapped
//   grad_x  = gxm * m[b,i]
//   gm1[b,i]= sum_hw gxm * x                                 (through the modulation of the activations)
//   gd[b,o] = sum_hw g * z = (sum_hw g * out) / d           (through the demodulation scale)
//   q[b,o]  = gd * d^3 = (sum_hw g * out) * d^2
//   gm2[b,i]= -m[b,i] * sum_o q[b,o] * wsq[i,o]             d(d)/d(m_i) = -d^3 m_i wsq[i,o]
//   grad_style = gm1 + gm2
//   grad_W[o,i,t] = sum_{b,y,x} gz[b,y,x,o] * xm[b,y+ty-p,x+tx-p,i]  -  W[o,i,t] * sum_b q[b,o] m[b,i]^2
//                   \_ wgrad: GEMM over K = B*H*W pixels _/            \_ d(d)/d(W) = -d^3 m_i^2 W _/
// The per-sample weights of the reference (ST:650-656) are never formed: W is shared by the batch in every product.
// All kernels are fp32 (the <= 1e-4 parity mode); the two big contractions are FFMA implicit GEMMs.
#pragma once

#include "common.cuh"
#include "conv_simt.cuh"

namespace sx {

// W[Co][Ci][k][k] -> wT[tap'][Co][Ci] with tap' = the spatially flipped tap: the dgrad is a conv of gz (Co channels in)
// with these weights (Ci channels out), packed like the forward's [tap][Cin][Cout].
__global__ void pack_weights_dgrad_kernel(const float* __restrict__ W, float* __restrict__ wT, int Co, int Ci, int taps) {
  const long long total = (long long)Co * Ci;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i / Ci), c = (int)(i - (long long)o * Ci);
    for (int t = 0; t < taps; ++t) wT[((long long)(taps - 1 - t) * Co + o) * Ci + c] = W[i * taps + t];
  }
}

// NCHW fp32 -> NHWC fp32 with an optional per-(b, c) scale (no "+1": the demodulation coefficients are used as they are)
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_scale_kernel(const float* __restrict__ x, const float* __restrict__ scale,
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
      if (scale) v *= __ldg(scale + (long long)b * C + c);
    }
    tile[j][tx] = v;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int p = p0 + j, c = c0 + tx;
    if (c < C && p < HW) out[((long long)b * HW + p) * C + c] = from_f<T>(tile[tx][j]);
  }
}

// bf16 operands of the tensor-core backward ------------------------------------------------------------------------
// dgrad weights, K-major for the tcgen05 conv kernels: wT[i][tap' * Co + o] = W[o][i][tap], tap' = flipped tap
__global__ void pack_weights_dgrad_bf16_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ wT, int Co, int Ci, int taps) {
  const long long total = (long long)Co * Ci;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(e / Ci), c = (int)(e - (long long)o * Ci);
    for (int t = 0; t < taps; ++t) wT[(long long)c * taps * Co + (long long)(taps - 1 - t) * Co + o] = __float2bfloat16_rn(W[e * taps + t]);
  }
}
// NCHW fp32 -> NCHW bf16 with a per-(b, c) plane scale: scale[plane] (+ add), e.g. d[b,o] or style[b,i] + 1
__global__ void __launch_bounds__(256) nchw_scale_bf16_kernel(const float* __restrict__ x, const float* __restrict__ scale, float add,
                                                              __nv_bfloat16* __restrict__ out, int HW, long long total) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const float sc = scale ? __ldg(scale + e / HW) + add : 1.f;
    out[e] = __float2bfloat16_rn(x[e] * sc);
  }
}
// the wgrad_tc B operand: `copies` (1 or 3) NCHW bf16 copies of x * (style + 1), copy s shifted by dx = s - copies/2 along x with
// zero padding:  out[s][b,c,y,x] = xm[b,c,y,x+dx]   (TMA cannot start a box at an odd pixel of the innermost dimension)
__global__ void __launch_bounds__(256) nchw_scale_shift_bf16_kernel(const float* __restrict__ x, const float* __restrict__ style,
                                                                    __nv_bfloat16* __restrict__ out, int HW, int W, int copies,
                                                                    long long total) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const float sc = __ldg(style + e / HW) + 1.f;
    const int xx = (int)(e % W);
    for (int s = 0; s < copies; ++s) {
      const int dx = s - copies / 2;
      const bool in = xx + dx >= 0 && xx + dx < W;
      out[(long long)s * total + e] = __float2bfloat16_rn(in ? x[e + dx] * sc : 0.f);
    }
  }
}

// One CTA per (b, c) plane of two NCHW tensors: dot[b,c] = sum_hw a * b; optionally a *= (style[b,c] + 1) in place
// afterwards (grad_x = gxm * m).  Fixed-shape tree reduction: deterministic.
__global__ void __launch_bounds__(256) plane_dot_kernel(float* __restrict__ a, const float* __restrict__ bsrc, const float* __restrict__ style,
                                                        float* __restrict__ dot, int HW) {
  const long long plane = blockIdx.x;
  float* pa = a + plane * HW;
  const float* pb = bsrc + plane * HW;
  float acc = 0.f;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) acc = fmaf(pa[i], __ldg(pb + i), acc);
  __shared__ float red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) dot[plane] = red[0];
  if (style) {
    const float m = __ldg(style + plane) + 1.f;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) pa[i] *= m;
  }
}

// q[b,o] = gdot[b,o] * d[b,o]^2   (in place on gdot);  msq[b,i] = (style+1)^2
__global__ void demod_grad_prep_kernel(float* __restrict__ gdot, const float* __restrict__ d, long long n_bo, const float* __restrict__ style,
                                       float* __restrict__ msq, long long n_bi) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_bo) {
    const float dv = d[i];
    gdot[i] = gdot[i] * dv * dv;
  }
  if (i < n_bi) {
    const float m = style[i] + 1.f;
    msq[i] = m * m;
  }
}

// grad_style[b,i] = gm1[b,i] - (style[b,i]+1) * sum_o q[b,o] * wsq[i][o]       (one warp per (b, i); q == null: no demod)
__global__ void __launch_bounds__(256) style_grad_kernel(const float* __restrict__ gm1, const float* __restrict__ style,
                                                         const float* __restrict__ q, const float* __restrict__ wsq,
                                                         float* __restrict__ grad_style, int B, int Ci, int Co) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= (long long)B * Ci) return;
  const int b = (int)(warp / Ci), i = (int)(warp - (long long)b * Ci);
  float acc = 0.f;
  if (q) {
    for (int o = lane; o < Co; o += 32) acc = fmaf(__ldg(q + (long long)b * Co + o), __ldg(wsq + (long long)i * Co + o), acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  }
  if (lane == 0) grad_style[warp] = gm1[warp] - (style[warp] + 1.f) * acc;
}

// ---- wgrad: partial[split][tap][o][i] = sum over the split's pixels of gz[pix][o] * xm[pix + tap offset][i] -----------
// GEMM view: M = Co, N = Ci, K = pixels.  64x64 output tile per CTA, K chunks of 16 pixels, 4x4 register micro-tile
// (the forward FFMA kernel's shape with the pixel axis as K); both operands are NHWC, so a K row is a contiguous channel
// vector of one pixel.  grid = (Co tiles * Ci tiles, taps, splits): the pixel range is split so the grid fills the GPU,
// and the partial sums are combined in a fixed order by wgrad_reduce_kernel (deterministic, no atomics).
struct WgradParams {
  const float* gz;   // [B,H,W,Co]
  const float* xm;   // [B,H,W,Ci]
  float* partial;    // [splits][taps][Co][Ci]
  int B, H, W, Ci, Co, KS, splits;
  long long pix_per_split;
};

__global__ void __launch_bounds__(256) wgrad_simt_kernel(WgradParams p) {
  __shared__ __align__(16) float As[16][64 + 4];   // [k][o]
  __shared__ __align__(16) float Bs[16][64 + 4];   // [k][i]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int tiles_i = (p.Ci + 63) / 64;
  const int to = blockIdx.x / tiles_i, ti = blockIdx.x - to * tiles_i;
  const int o0 = to * 64, i0 = ti * 64;
  const int tap = blockIdx.y, split = blockIdx.z;
  const int pad = (p.KS - 1) / 2;
  const int dy = tap / p.KS - pad, dx = tap % p.KS - pad;
  const long long HW = (long long)p.H * p.W, M = (long long)p.B * HW;
  const long long k_begin = (long long)split * p.pix_per_split;
  const long long k_end = k_begin + p.pix_per_split < M ? k_begin + p.pix_per_split : M;

  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  // loaders: thread -> (k row = tid / 16, 4 channels starting at (tid % 16) * 4)
  const int l_k = tid >> 4, l_c = (tid & 15) * 4;
  for (long long k0 = k_begin; k0 < k_end; k0 += 16) {
    const long long pix = k0 + l_k;
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (pix < k_end) {
      const float* ga = p.gz + pix * p.Co + o0 + l_c;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (o0 + l_c + j < p.Co) av[j] = __ldg(ga + j);
      const int b = (int)(pix / HW);
      const int r = (int)(pix - (long long)b * HW);
      const int y = r / p.W + dy, x = r % p.W + dx;
      if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
        const float* xb = p.xm + (((long long)b * p.H + y) * p.W + x) * p.Ci + i0 + l_c;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (i0 + l_c + j < p.Ci) bv[j] = __ldg(xb + j);
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[l_k][l_c + j] = av[j];
      Bs[l_k][l_c + j] = bv[j];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], bb[v], acc[u][v]);
    }
  }
  float* dst = p.partial + ((long long)split * p.KS * p.KS + tap) * p.Co * p.Ci;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int o = o0 + ty * 4 + u;
    if (o >= p.Co) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int i = i0 + tx * 4 + v;
      if (i < p.Ci) dst[(long long)o * p.Ci + i] = acc[u][v];
    }
  }
}

// grad_W[o][i][t] = sum_split partial[split][t][o][i]  -  W[o][i][t] * sum_b q[b,o] * msq[b,i]        (q == null: no demod)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int splits, const float* __restrict__ W,
                                                           const float* __restrict__ q, const float* __restrict__ msq,
                                                           float* __restrict__ grad_w, int B, int Co, int Ci, int taps) {
  const long long total = (long long)Co * Ci;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(e / Ci), i = (int)(e - (long long)o * Ci);
    float a = 0.f;
    if (q)
      for (int b = 0; b < B; ++b) a = fmaf(__ldg(q + (long long)b * Co + o), __ldg(msq + (long long)b * Ci + i), a);
    for (int t = 0; t < taps; ++t) {
      float s = 0.f;
      for (int sp = 0; sp < splits; ++sp) s += __ldg(partial + (((long long)sp * taps + t) * Co + o) * Ci + i);
      grad_w[e * taps + t] = s - W[e * taps + t] * a;
    }
  }
}

inline int wgrad_splits(int B, int Ci, int Co, int H, int W, int k) {
  const long long tiles = (long long)((Co + 63) / 64) * ((Ci + 63) / 64) * k * k;
  const long long M = (long long)B * H * W;
  long long s = (2LL * num_sms() + tiles - 1) / tiles;
  const long long max_s = (M + 255) / 256;   // at least 256 pixels per split
  if (s > max_s) s = max_s;
  if (s > 256) s = 256;
  if (s < 1) s = 1;
  return (int)s;
}

}  // namespace sx
