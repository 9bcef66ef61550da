// conv_simt.cuh -- fp32 implicit-GEMM convolution on the CUDA cores (FFMA), NHWC activations.
//
// This is the SX_PREC_FP32 path: the <=1e-4 parity mode of the modulated convolution (tcgen05 has no
// fp32 MMA), the kernel behind the generic Conv2DMod op for shapes the tensor-core kernel does not
// take (k=1, tiny channel counts), and the cross-check of the tcgen05 kernel (sx_tc_selftest).
//
// GEMM view (reference Conv2DMod.forward ST:647-667 with the modulation moved to the activations):
//   M = B*H*W pixels, N = Co, K = k*k*Ci;  A[m, (tap,ci)] = xmod[b, y+dy, x+dx, ci] (zero padded),
//   B[(tap,ci), o] = W[o, ci, tap]  (packed [tap][Ci][Co], shared by the whole batch).
// 256 threads, register micro-tiles, double-buffered shared memory (tile configurations below).
#pragma once

#include "common.cuh"

namespace sx {

constexpr int SIMT_BK = 16, SIMT_THREADS = 256;

struct ConvSimtParams {
  const float* x;       // [Bx, H, W, Ci] NHWC, already modulated by (style+1)
  long long x_bstride;  // elements between samples (0: one sample broadcast to the whole batch)
  const float* wpk;     // [taps][Ci][Co]
  int B, Ci, Co, H, W, KS;
  ConvEpilogue ep;
};

// packed fp32x2 FMA (sm_100 FFMA2: two independent IEEE fmas per instruction -- same results as two fmaf)
__device__ __forceinline__ uint64_t simt_pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void simt_upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t simt_fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// Tile BM pixels x BN channels x 16, 256 threads, TM x TN register micro-tile.  Every output is the same sequential fp32
// FMA chain over (tap, channel) whatever the tiling, so all configurations give bit-identical results:
//   <128, 64, 8, 4>  large layers: 32 FMAs per 12 shared-memory floats (the 64 x 64 / 4 x 4 form: 16 per 8)
//   <128, 32, 4, 4>  Co <= 32 (the 256 px layers of the generator; a 64-wide tile wasted half of its columns there)
//   < 64, 64, 4, 4>  small launches (more CTAs)
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(SIMT_THREADS) conv_simt_kernel(ConvSimtParams p) {
  constexpr int TX = BN / TN, TY = BM / TM;
  static_assert(TX * TY == SIMT_THREADS && TN == 4 && TM % 4 == 0 && BM % 64 == 0, "tile / thread layout");
  constexpr int A_LD = BM / 64;                 // float4 A loads per thread and K chunk
  constexpr int B_COLS = BN / 4;                // float4 columns of the B tile
  __shared__ __align__(16) float As[2][SIMT_BK][BM + 4];
  __shared__ __align__(16) float Bs[2][SIMT_BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int HW = p.H * p.W;
  const long long M = (long long)p.B * HW;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int pad = (p.KS - 1) / 2;
  const int taps = p.KS * p.KS;
  const int kchunks = (p.Ci + SIMT_BK - 1) / SIMT_BK;
  const int iters = taps * kchunks;
  const bool ci_vec = (p.Ci & 3) == 0;
  const bool co_vec = (p.Co & 3) == 0;

  // A loader: pixels a_pix + 64 l of the tile, 4 channels starting at a_cg of the K chunk
  const int a_pix = tid >> 2, a_cg = (tid & 3) * 4;
  bool a_ok[A_LD];
  int a_y[A_LD], a_x[A_LD];
  const float* a_base[A_LD];
#pragma unroll
  for (int l = 0; l < A_LD; ++l) {
    const long long a_m = m0 + a_pix + 64 * l;
    a_ok[l] = a_m < M;
    int a_b = 0;
    a_y[l] = a_x[l] = 0;
    if (a_ok[l]) {
      a_b = (int)(a_m / HW);
      const int r = (int)(a_m - (long long)a_b * HW);
      a_y[l] = r / p.W;
      a_x[l] = r - a_y[l] * p.W;
    }
    a_base[l] = p.x + (long long)a_b * p.x_bstride;
  }
  // B loader: K row b_kr of the chunk, 4 output channels starting at b_ng (threads beyond the tile idle)
  const int b_kr = tid / B_COLS, b_ng = (tid % B_COLS) * 4;
  const bool b_active = b_kr < SIMT_BK;

  float a_reg[A_LD][4], b_reg[4];
  auto load_tiles = [&](int it) {
    const int tap = it / kchunks;
    const int c0 = (it - tap * kchunks) * SIMT_BK;
    const int dy = tap / p.KS - pad, dx = tap % p.KS - pad;
    // ---- A
    const int c = c0 + a_cg;
#pragma unroll
    for (int l = 0; l < A_LD; ++l) {
      const int sy = a_y[l] + dy, sx_ = a_x[l] + dx;
      const bool in = a_ok[l] && sy >= 0 && sy < p.H && sx_ >= 0 && sx_ < p.W;
#pragma unroll
      for (int j = 0; j < 4; ++j) a_reg[l][j] = 0.f;
      if (in) {
        const float* src = a_base[l] + ((long long)sy * p.W + sx_) * p.Ci + c;
        if (ci_vec && c + 3 < p.Ci) {
          float4 v = __ldg(reinterpret_cast<const float4*>(src));
          a_reg[l][0] = v.x; a_reg[l][1] = v.y; a_reg[l][2] = v.z; a_reg[l][3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c + j < p.Ci) a_reg[l][j] = __ldg(src + j);
        }
      }
    }
    // ---- B
    const int kc = c0 + b_kr;
    const int n = n0 + b_ng;
#pragma unroll
    for (int j = 0; j < 4; ++j) b_reg[j] = 0.f;
    if (b_active && kc < p.Ci) {
      const float* src = p.wpk + ((long long)tap * p.Ci + kc) * p.Co + n;
      if (co_vec && n + 3 < p.Co) {
        float4 v = __ldg(reinterpret_cast<const float4*>(src));
        b_reg[0] = v.x; b_reg[1] = v.y; b_reg[2] = v.z; b_reg[3] = v.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < p.Co) b_reg[j] = __ldg(src + j);
      }
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int l = 0; l < A_LD; ++l)
#pragma unroll
      for (int j = 0; j < 4; ++j) As[buf][a_cg + j][a_pix + 64 * l] = a_reg[l][j];
    if (b_active) *reinterpret_cast<float4*>(&Bs[buf][b_kr][b_ng]) = make_float4(b_reg[0], b_reg[1], b_reg[2], b_reg[3]);
  };

  // accumulators packed two rows per register pair: acc2[i2][j] = (row 2 i2, row 2 i2 + 1) of column j.  The row pairs come
  // straight out of the float4 shared-memory loads; the TN column values are broadcast into both halves once per k.
  // TM * TN / 2 FFMA2 + TN moves per k instead of TM * TN FFMA; per output element the same sequential fma chain.
  uint64_t acc2[TM / 2][TN];
#pragma unroll
  for (int i = 0; i < TM / 2; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc2[i][j] = 0ull;

  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
    const int buf = it & 1;
    if (it + 1 < iters) load_tiles(it + 1);
#pragma unroll
    for (int k = 0; k < SIMT_BK; ++k) {
      uint64_t pa[TM / 2];
#pragma unroll
      for (int i4 = 0; i4 < TM / 4; ++i4) {
        const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + 4 * i4]);
        pa[2 * i4] = simt_pk2(a.x, a.y);
        pa[2 * i4 + 1] = simt_pk2(a.z, a.w);
      }
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN]);
      const uint64_t bb[4] = {simt_pk2(b.x, b.x), simt_pk2(b.y, b.y), simt_pk2(b.z, b.z), simt_pk2(b.w, b.w)};
#pragma unroll
      for (int i = 0; i < TM / 2; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc2[i][j] = simt_fma2(pa[i], bb[j], acc2[i][j]);
    }
    if (it + 1 < iters) store_tiles(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM / 2; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) simt_upk2(acc2[i][j], acc[2 * i][j], acc[2 * i + 1][j]);
  const ConvEpilogue& ep = p.ep;
  const int S = ep.noise_size;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long m = m0 + ty * TM + i;
    if (m >= M) continue;
    const int b = (int)(m / HW);
    const int r = (int)(m - (long long)b * HW);
    const int y = r / p.W, x = r - y * p.W;
    float nz = 0.f;
    if (ep.noise) nz = __ldg(ep.noise + (long long)(ep.noise_batch == 1 ? 0 : b) * S * S + (long long)x * S + y);
    float v[4], vr[4];
    const int o0 = n0 + tx * TN;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int o = o0 + j;
      float t = acc[i][j];
      if (o < p.Co) {
        if (ep.dcoef) t *= __ldg(ep.dcoef + (long long)b * ep.dcoef_stride + o);
        if (ep.noise) t += nz * __ldg(ep.noise_w + o) + __ldg(ep.noise_b + o);
        if (ep.act) t = lrelu02(t);
      }
      vr[j] = t;
      if (o < p.Co && ep.next_style) t *= __ldg(ep.next_style + (long long)b * ep.next_style_stride + o) + 1.f;
      v[j] = t;
    }
    if (ep.out_nchw_f32) {
      float* out = reinterpret_cast<float*>(ep.out);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (o0 + j < p.Co) out[(((long long)b * p.Co + o0 + j) * p.H + y) * p.W + x] = v[j];
    } else {
      float* out = reinterpret_cast<float*>(ep.out) + m * p.Co + o0;
      if (co_vec && o0 + 3 < p.Co) {
        *reinterpret_cast<float4*>(out) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (o0 + j < p.Co) out[j] = v[j];
      }
    }
    if (ep.out_raw) {
      float* out = reinterpret_cast<float*>(ep.out_raw) + m * p.Co + o0;
      if (co_vec && o0 + 3 < p.Co) {
        *reinterpret_cast<float4*>(out) = make_float4(vr[0], vr[1], vr[2], vr[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (o0 + j < p.Co) out[j] = vr[j];
      }
    }
  }
}

inline int launch_conv_simt(const ConvSimtParams& p, cudaStream_t stream) {
  SX_REQUIRE(p.KS == 1 || p.KS == 3, "conv_simt: kernel size %d not supported (1 or 3)", p.KS);
  const long long M = (long long)p.B * p.H * p.W;
  if (M == 0 || p.Co == 0) return SX_OK;
  static const bool old_tile = getenv("SX_SIMT_64") != nullptr;   // A/B: the 64 x 64 / 4 x 4 form everywhere
  const long long ctas_big = ((M + 127) / 128) * ((p.Co + 63) / 64);
  if (!old_tile && p.Co <= 32 && M >= 128 * 2 * num_sms()) {
    dim3 grid((unsigned)((M + 127) / 128), (unsigned)((p.Co + 31) / 32));
    conv_simt_kernel<128, 32, 4, 4><<<grid, SIMT_THREADS, 0, stream>>>(p);
  } else if (!old_tile && ctas_big >= 2 * num_sms()) {
    dim3 grid((unsigned)((M + 127) / 128), (unsigned)((p.Co + 63) / 64));
    conv_simt_kernel<128, 64, 8, 4><<<grid, SIMT_THREADS, 0, stream>>>(p);
  } else {
    dim3 grid((unsigned)((M + 63) / 64), (unsigned)((p.Co + 63) / 64));
    conv_simt_kernel<64, 64, 4, 4><<<grid, SIMT_THREADS, 0, stream>>>(p);
  }
  SX_CHECK_LAUNCH();
  return SX_OK;
}

}  // namespace sx
