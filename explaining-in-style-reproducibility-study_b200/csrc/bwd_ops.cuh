// bwd_ops.cuh -- first-order backward of the bandwidth-bound generator ops in the reference's layout (NCHW fp32):
// nn.Linear (to_style*, ST:681,685,608), noise + leaky-ReLU (ST:696-698,705,714), bilinear 2x upsample (ST:679,614) and
// Blur (ST:144-153).  Every kernel GATHERS (one thread per gradient element, fixed summation order): deterministic,
// no atomics.  Training-step slice, SURVEY.md section 8f row 1.
#pragma once

#include "common.cuh"

namespace sx {

// ---- nn.Linear: out[b,n] = sum_k x[b,k] W[n,k] + bias[n] ---------------------------------------------------------
// grad_x[b,k] = sum_n g[b,n] W[n,k]
__global__ void __launch_bounds__(256) linear_bwd_x_kernel(const float* __restrict__ g, const float* __restrict__ W, float* __restrict__ gx,
                                                           int B, int K, int Nf) {
  const long long total = (long long)B * K;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(e / K), k = (int)(e - (long long)b * K);
    float acc = 0.f;
#pragma unroll 8
    for (int n = 0; n < Nf; ++n) acc = fmaf(__ldg(g + (long long)b * Nf + n), __ldg(W + (long long)n * K + k), acc);
    gx[e] = acc;
  }
}
// grad_W[n,k] = sum_b g[b,n] x[b,k];   grad_bias[n] = sum_b g[b,n]
__global__ void __launch_bounds__(256) linear_bwd_w_kernel(const float* __restrict__ g, const float* __restrict__ x, float* __restrict__ gW,
                                                           float* __restrict__ gb, int B, int K, int Nf) {
  const long long total = (long long)Nf * K;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(e / K), k = (int)(e - (long long)n * K);
    float acc = 0.f, accb = 0.f;
    for (int b = 0; b < B; ++b) {
      const float gv = __ldg(g + (long long)b * Nf + n);
      acc = fmaf(gv, __ldg(x + (long long)b * K + k), acc);
      accb += gv;
    }
    gW[e] = acc;
    if (k == 0 && gb) gb[n] = accb;
  }
}

// ---- noise + leaky-ReLU: out = lrelu(x + noise[b|0, xx, yy] * nw[c] + nb[c]) ----------------------------------------
// One CTA per (b, c) plane: gx = g * (out > 0 ? 1 : 0.2)   [sign(out) == sign(pre-activation); torch uses slope at 0]
// and the plane's partial sums  pw[b,c] = sum gx * noise (transposed, quirk Q1),  pb[b,c] = sum gx.
__global__ void __launch_bounds__(256) noise_lrelu_bwd_kernel(const float* __restrict__ out, const float* __restrict__ g,
                                                              const float* __restrict__ inoise, float* __restrict__ gx,
                                                              float* __restrict__ pw, float* __restrict__ pb, int C, int H, int W,
                                                              int noise_batch, int S) {
  const long long plane = blockIdx.x;
  const int b = (int)(plane / C);
  const float* nz = inoise + (long long)(noise_batch == 1 ? 0 : b) * S * S;
  const long long base = plane * H * W;
  float aw = 0.f, ab = 0.f;
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    const int yy = i / W, xx = i - yy * W;
    const float v = g[base + i] * (out[base + i] > 0.f ? 1.f : 0.2f);
    gx[base + i] = v;
    aw = fmaf(v, __ldg(nz + (long long)xx * S + yy), aw);
    ab += v;
  }
  __shared__ float rw[256], rb[256];
  rw[threadIdx.x] = aw;
  rb[threadIdx.x] = ab;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      rw[threadIdx.x] += rw[threadIdx.x + s];
      rb[threadIdx.x] += rb[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    pw[plane] = rw[0];
    pb[plane] = rb[0];
  }
}
// grad_nw[c] = sum_b pw[b,c], grad_nb[c] = sum_b pb[b,c]   (fixed order)
__global__ void noise_param_reduce_kernel(const float* __restrict__ pw, const float* __restrict__ pb, float* __restrict__ gnw,
                                          float* __restrict__ gnb, int B, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float aw = 0.f, ab = 0.f;
  for (int b = 0; b < B; ++b) {
    aw += pw[(long long)b * C + c];
    ab += pb[(long long)b * C + c];
  }
  gnw[c] = aw;
  gnb[c] = ab;
}

// ---- bilinear 2x upsample (align_corners=False), adjoint ------------------------------------------------------------
// Forward per axis (n inputs -> 2n outputs): out[2i] = 0.25 in[max(i-1,0)] + 0.75 in[i],  out[2i+1] = 0.75 in[i] + 0.25 in[min(i+1,n-1)].
// Adjoint weights of input i over the outputs 2i-1 .. 2i+2 (clamped taps fold back onto the border input).
__device__ __forceinline__ void up2_adjoint_taps(int i, int n, int* o, float* w) {
  o[0] = 2 * i - 1; w[0] = i >= 1 ? 0.25f : 0.f;           // out[2(i-1)+1] reads in[i] with 0.25
  o[1] = 2 * i;     w[1] = i == 0 ? 1.0f : 0.75f;          // out[0] = in[0] (clamped tap)
  o[2] = 2 * i + 1; w[2] = i == n - 1 ? 1.0f : 0.75f;      // out[2n-1] = in[n-1]
  o[3] = 2 * i + 2; w[3] = i <= n - 2 ? 0.25f : 0.f;       // out[2(i+1)] reads in[i] with 0.25
}
__global__ void __launch_bounds__(256) upsample2x_bwd_nchw_kernel(const float* __restrict__ g, float* __restrict__ gx, long long planes,
                                                                  int H, int W) {
  const int OW = 2 * W;
  const long long total = planes * H * W;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % W);
    const long long r = e / W;
    const int i = (int)(r % H);
    const float* gp = g + (r / H) * 4 * H * W;
    int oy[4], ox[4];
    float wy[4], wx[4];
    up2_adjoint_taps(i, H, oy, wy);
    up2_adjoint_taps(j, W, ox, wx);
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (wy[a] == 0.f) continue;
      float row = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (wx[b] != 0.f) row = fmaf(wx[b], __ldg(gp + (long long)oy[a] * OW + ox[b]), row);
      acc = fmaf(wy[a], row, acc);
    }
    gx[e] = acc;
  }
}

// ---- Blur ([1,2,1]x[1,2,1]/16, reflect border), adjoint --------------------------------------------------------------
// Forward per axis: out[j] = 0.25 in[refl(j-1)] + 0.5 in[j] + 0.25 in[refl(j+1)], refl(-1) = 1, refl(n) = n-2.
// Adjoint weight of g[a] (a in {i-1, i, i+1}) in grad_in[i]: the plain taps plus the two reflected ones.
__device__ __forceinline__ float blur_adjoint_w(int i, int a, int n) {
  if (a < 0 || a > n - 1) return 0.f;
  float w = a == i ? 0.5f : 0.25f;
  if (i == 1 && a == 0) w += 0.25f;              // out[0] reads in[refl(-1)] = in[1]
  if (i == n - 2 && a == n - 1) w += 0.25f;      // out[n-1] reads in[refl(n)] = in[n-2]
  return w;
}
__global__ void __launch_bounds__(256) blur_bwd_nchw_kernel(const float* __restrict__ g, float* __restrict__ gx, long long planes, int H,
                                                            int W) {
  const long long total = planes * H * W;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % W);
    const long long r = e / W;
    const int i = (int)(r % H);
    const float* gp = g + (r / H) * H * W;
    float acc = 0.f;
#pragma unroll
    for (int da = -1; da <= 1; ++da) {
      const float wy = blur_adjoint_w(i, i + da, H);
      if (wy == 0.f) continue;
      float row = 0.f;
#pragma unroll
      for (int db = -1; db <= 1; ++db) {
        const float wx = blur_adjoint_w(j, j + db, W);
        if (wx != 0.f) row = fmaf(wx, __ldg(gp + (long long)(i + da) * W + (j + db)), row);
      }
      acc = fmaf(wy, row, acc);
    }
    gx[e] = acc;
  }
}

}  // namespace sx
