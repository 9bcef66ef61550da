// stem_tc.cuh -- the ResNet stem of the AttFind classifier (resnet_classifier.py:56-71 -> torchvision conv1 + bn1 + relu
// [+ maxpool]) as one persistent tcgen05 kernel.  The 7x7 / stride-2 / pad-3 convolution on 3 channels is expressed on the
// 2x2 space-to-depth image (classifiers.py enable_s2d_stem: 4x4 taps, stride 1, no padding, 12 -> 16 channels), which is
// what sx_resize_aa_normalize_s2d writes.  cuDNN ran that shape at ~0.45 ms per 256 images (K = 16 per tap is too thin for
// its tiles) and the separate max-pool re-read the 411 MB stem output: 10 % of the AttFind step.
//
// A CTA super-tile is 16 x 16 output pixels = two M128 tiles of 16 rows x 8 pixels.  Per tile ONE TMA box of 19 x 11 input
// pixels x 16 channels (32-byte rows, SWIZZLE_32B; out of bounds = zero) is loaded and the 16 taps read it in place,
// exactly like conv_tc_halo.cuh: A operand of tap (ky, kx) = start (ky * 11 + kx) * 32 B, 16 groups of 8 rows, 11 rows
// apart; K = 16 is one tcgen05.mma per tap.  The 16 x [64 x 16] weight tiles (32 KB) stay resident in shared memory.
// The two tiles of a super-tile accumulate into the two halves of one 128-column TMEM slot.
//
// POOL: the 3x3 / stride-2 / pad-1 max-pool is fused.  Super-tiles then advance by 14 pixels and start at -1, so each one
// holds every conv value its 7 x 7 pool outputs need (rows/cols 2i-1 .. 2i+1): 1.31x the MMAs, but the 411 MB
// intermediate is never written or read.  The epilogue set rounds relu(acc + bias) to bf16 (what cuDNN stores; max
// commutes with the monotone rounding), stages the 16 x 16 x 64 block in shared memory (16-byte chunks XOR-swizzled by
// pixel) and 128 threads reduce the windows.  Conv positions outside the image take the value 0: every window holds at
// least one real value and those are >= 0 after the ReLU, so this equals the pool's -inf padding.
#pragma once

#include "conv_tc.cuh"

namespace sx {
namespace tc {

constexpr int STEM_CI = 16, STEM_CO = 64, STEM_TAPS = 16;
constexpr int STEM_BW = 8, STEM_BH = 16;
constexpr int STEM_HW = STEM_BW + 3, STEM_HH = STEM_BH + 3;           // 11 x 19 input pixels per tile
constexpr int STEM_ROW_BYTES = STEM_CI * 2;                            // 32
constexpr int STEM_BOX_BYTES = STEM_HW * STEM_HH * STEM_ROW_BYTES;     // 6688
constexpr int STEM_STAGE_BYTES = 7168;                                 // box rounded up to 1024 B (swizzle phase)
constexpr int STEM_TAP_BYTES = STEM_CO * STEM_ROW_BYTES;               // 2048
constexpr int STEM_W_BYTES = STEM_TAPS * STEM_TAP_BYTES;               // 32768
constexpr int STEM_SLOT_COLS = 2 * STEM_CO;                            // two tiles per accumulator slot
constexpr int STEM_STAGING_BYTES = 16 * 16 * STEM_CO * 2;              // 32768 per epilogue set (POOL)

struct StemParams {
  int B, Hout, Wout;            // conv output
  int Ho, Wo;                   // what is stored: the conv output, or (POOL) the pooled map
  int tiles_x, tiles_y, num_tiles;
  int debug;                    // SX_STEM_DEBUG bitmask (bottleneck experiments; results are garbage when set):
                                //   1 skip epilogue math / stores, 2 skip MMA issue, 4 skip the activation TMA loads
  const float* bias;
  __nv_bfloat16* out;           // NHWC [B, Ho, Wo, 64]
};

template <int STAGES, int SETS, bool POOL>
struct StemCfg {
  static_assert(STAGES % 2 == 0 && SETS >= 1 && SETS * STEM_SLOT_COLS <= 512, "ring / TMEM shape");
  static constexpr int kThreads = 128 + 128 * SETS;
  static constexpr int kTmemCols = SETS * STEM_SLOT_COLS <= 128 ? 128 : SETS * STEM_SLOT_COLS <= 256 ? 256 : 512;
  static constexpr int kStaging = POOL ? SETS * STEM_STAGING_BYTES : 0;
  static constexpr int kStep = POOL ? 14 : 16, kOrigin = POOL ? -1 : 0;
  static size_t smem_bytes() { return 1024 + STEM_W_BYTES + (size_t)STAGES * STEM_STAGE_BYTES + kStaging + 256 + STEM_CO * sizeof(float); }
};

// K-major SWIZZLE_32B descriptor: 32-byte rows, 8-row groups sbo bytes apart
__device__ __forceinline__ uint64_t make_smem_desc32(uint32_t smem_addr, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (6ull << 61);
}
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("max.NaN.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
// packed fp32x2 -> bf16x2 (round to nearest even) with negative values clamped to +0
__device__ __forceinline__ uint32_t relu_bf16x2_rn(uint64_t v2) {
  float lo, hi;
  upk2(v2, lo, hi);
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

template <int STAGES, int SETS, bool POOL>
__global__ void __launch_bounds__(StemCfg<STAGES, SETS, POOL>::kThreads)
stem_s2d_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, const StemParams p) {
  using Cfg = StemCfg<STAGES, SETS, POOL>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem_w + STEM_W_BYTES;
  uint8_t* smem_stage = smem_a + STAGES * STEM_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_stage + Cfg::kStaging);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + SETS;
  uint64_t* w_bar = tmem_empty_bar + SETS;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_bar + 1);
  float* s_bias = reinterpret_cast<float*>(smem_stage + Cfg::kStaging + 256);
  static_assert((2 * STAGES + 2 * SETS + 1) * 8 + 4 <= 256, "barrier block overflow");

  const int warp_id = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_per_b = p.tiles_x * p.tiles_y;

  if (warp_id == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < SETS; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 128);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  } else if (warp_id == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
  } else if (warp_id >= 2 && warp_id < 4) {
    s_bias[threadIdx.x - 64] = p.bias[threadIdx.x - 64];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_id == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(w_bar, STEM_W_BYTES);
#pragma unroll
      for (int j = 0; j < 4; ++j) tma_load_2d(smem_w + j * (STEM_W_BYTES / 4), &tmap_w, w_bar, 0, j * (STEM_TAPS * STEM_CO / 4));
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int st = blockIdx.x; st < p.num_tiles; st += gridDim.x) {
      const int b = st / tiles_per_b, r = st - b * tiles_per_b;
      const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const int y0 = ty * Cfg::kStep + Cfg::kOrigin, x0 = tx * Cfg::kStep + Cfg::kOrigin;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 0);
        if (elect_one()) {
          if (p.debug & 4) {
            mbar_arrive(&full_bar[stage]);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], STEM_BOX_BYTES);
            tma_load_4d(smem_a + stage * STEM_STAGE_BYTES, &tmap_a, &full_bar[stage], 0, x0 + STEM_BW * half, y0, b);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp_id == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(STEM_CO);
    mbar_wait(w_bar, 0, 5);
    tc_fence_after();
    const uint32_t w0 = smem_u32(smem_w);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int st = blockIdx.x; st < p.num_tiles; st += gridDim.x, ++it) {
      const int slot = it % SETS;
      mbar_wait(&tmem_empty_bar[slot], (uint32_t)(((it / SETS) & 1) ^ 1), 3);
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t tmem_d = tmem_base + (uint32_t)(slot * STEM_SLOT_COLS + half * STEM_CO);
        mbar_wait(&full_bar[stage], phase, 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(smem_a + stage * STEM_STAGE_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int tap = 0; tap < ((p.debug & 2) ? 1 : STEM_TAPS); ++tap) {
            const uint64_t da = make_smem_desc32(a0 + (uint32_t)(((tap >> 2) * STEM_HW + (tap & 3)) * STEM_ROW_BYTES), STEM_HW * STEM_ROW_BYTES);
            const uint64_t db = make_smem_desc32(w0 + (uint32_t)(tap * STEM_TAP_BYTES), 8 * STEM_ROW_BYTES);
            umma_bf16(tmem_d, da, db, idesc, tap != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (half == 1) umma_commit(&tmem_full_bar[slot]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp_id >= 4) {
    // ===================== epilogue sets =====================
    const int set = (warp_id - 4) >> 2;
    const int q = warp_id & 3;
    const int et = threadIdx.x - 128 - set * 128;   // 0..127 inside the set
    const int r = q * 32 + lane;                    // TMEM lane = tile pixel
    const int yy = r >> 3, xx = r & 7;
    uint8_t* staging = smem_stage + set * STEM_STAGING_BYTES;
    for (int it = set;; it += SETS) {
      const int st = blockIdx.x + it * (int)gridDim.x;
      if (st >= p.num_tiles) break;
      const int b = st / tiles_per_b, rr = st - b * tiles_per_b;
      const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
      const int y = ty * Cfg::kStep + Cfg::kOrigin + yy;
      mbar_wait(&tmem_full_bar[set], (uint32_t)((it / SETS) & 1), 2);
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * STEM_SLOT_COLS + half * STEM_CO);
        uint32_t v[64];
        tmem_ld32(taddr, v);
        tmem_ld32(taddr + 32, v + 32);
        tmem_ld_wait();
        if (half == 1) {
          tc_fence_before();
          mbar_arrive(&tmem_empty_bar[set]);   // both halves are in registers: the MMA warp may refill the slot
        }
        if (p.debug & 1) continue;
        const int xl = STEM_BW * half + xx;
        const int x = tx * Cfg::kStep + Cfg::kOrigin + xl;
        const bool valid = y >= 0 && y < p.Hout && x >= 0 && x < p.Wout;
        // relu(acc + bias) -> bf16: one packed FFMA2 (x * 1 + b is the exact sum) and one converting instruction with the
        // ReLU folded in per channel pair
        uint32_t w[32];
        const uint64_t one2 = pk2(1.f, 1.f);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 bb = *reinterpret_cast<const float4*>(s_bias + 4 * j);
          w[2 * j] = relu_bf16x2_rn(fma2(pk2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), one2, pk2(bb.x, bb.y)));
          w[2 * j + 1] = relu_bf16x2_rn(fma2(pk2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), one2, pk2(bb.z, bb.w)));
        }
        if (POOL) {
          const int pl = yy * 16 + xl;
          uint8_t* dst = staging + pl * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 o = valid ? make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]) : make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(dst + ((c ^ (pl & 7)) << 4)) = o;
          }
        } else if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + (((size_t)b * p.Hout + y) * p.Wout + x) * STEM_CO);
#pragma unroll
          for (int c = 0; c < 8; ++c) dst[c] = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        }
      }
      if (POOL && !(p.debug & 1)) {
        named_bar_sync(1 + set, 128);
        for (int wi = et; wi < 49 * 8; wi += 128) {
          const int pp = wi >> 3, c = wi & 7;
          const int pi = pp / 7, pj = pp - pi * 7;
          const int gi = ty * 7 + pi, gj = tx * 7 + pj;
          if (gi < p.Ho && gj < p.Wo) {
            uint4 m = make_uint4(0u, 0u, 0u, 0u);   // conv values are >= 0 (bf16 +0 = the identity of this max)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                const int pl = (2 * pi + dy) * 16 + 2 * pj + dx;
                const uint4 t = *reinterpret_cast<const uint4*>(staging + pl * 128 + ((c ^ (pl & 7)) << 4));
                m.x = max_bf16x2(m.x, t.x);
                m.y = max_bf16x2(m.y, t.y);
                m.z = max_bf16x2(m.z, t.z);
                m.w = max_bf16x2(m.w, t.w);
              }
            }
            *reinterpret_cast<uint4*>(p.out + (((size_t)b * p.Ho + gi) * p.Wo + gj) * STEM_CO + c * 8) = m;
          }
        }
        named_bar_sync(1 + set, 128);   // the staging block is free for the set's next super-tile
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_id == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---- host side -------------------------------------------------------------------------------------
template <int STAGES, int SETS, bool POOL>
int launch_stem_cfg(const __nv_bfloat16* x, const __nv_bfloat16* w_taps, const float* bias, __nv_bfloat16* out, int B, int Hin, int Win,
                    cudaStream_t stream) {
  using Cfg = StemCfg<STAGES, SETS, POOL>;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(SX_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  CUtensorMap ta, tw;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)STEM_CI, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)STEM_ROW_BYTES, (cuuint64_t)Win * STEM_ROW_BYTES, (cuuint64_t)Hin * Win * STEM_ROW_BYTES};
    cuuint32_t box[4] = {(cuuint32_t)STEM_CI, (cuuint32_t)STEM_HW, (cuuint32_t)STEM_HH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(x), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SX_ECUDA, "cuTensorMapEncodeTiled(stem input) failed: %d (B=%d H=%d W=%d)", (int)r, B, Hin, Win);
  }
  {
    cuuint64_t gdim[2] = {(cuuint64_t)STEM_CI, (cuuint64_t)STEM_TAPS * STEM_CO};
    cuuint64_t gstr[1] = {(cuuint64_t)STEM_ROW_BYTES};
    cuuint32_t box[2] = {(cuuint32_t)STEM_CI, (cuuint32_t)(STEM_TAPS * STEM_CO / 4)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(w_taps), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SX_ECUDA, "cuTensorMapEncodeTiled(stem weights) failed: %d", (int)r);
  }
  StemParams p;
  p.B = B;
  p.Hout = Hin - 3;
  p.Wout = Win - 3;
  p.Ho = POOL ? (p.Hout - 1) / 2 + 1 : p.Hout;
  p.Wo = POOL ? (p.Wout - 1) / 2 + 1 : p.Wout;
  const int per = POOL ? 7 : 16;
  p.tiles_y = (p.Ho + per - 1) / per;
  p.tiles_x = (p.Wo + per - 1) / per;
  const long long total = (long long)B * p.tiles_y * p.tiles_x;
  if (total > 0x7fffffffLL) return fail(SX_EINVAL, "stem: too many tiles");
  p.num_tiles = (int)total;
  p.bias = bias;
  p.out = out;
  static const int debug = [] { const char* e = getenv("SX_STEM_DEBUG"); return e ? atoi(e) : 0; }();
  p.debug = debug;
  auto kern = stem_s2d_tc_kernel<STAGES, SETS, POOL>;
  const size_t smem = Cfg::smem_bytes();
  static bool configured = false;  // per instantiation
  if (!configured) {
    SX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const unsigned grid = (unsigned)(total < (long long)num_sms() ? total : (long long)num_sms());
  kern<<<grid, Cfg::kThreads, smem, stream>>>(ta, tw, p);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

// x: NHWC bf16 [B, Hin, Win, 16] (the space-to-depth network input); w_taps: bf16 [4][4][64][16] (ky, kx, co, ci);
// bias: fp32 [64]; out: NHWC bf16 [B, Hin-3, Win-3, 64], or with fuse_pool the 3x3/2/1 max-pooled map of it.
inline int launch_stem_s2d(const __nv_bfloat16* x, const __nv_bfloat16* w_taps, const float* bias, __nv_bfloat16* out, int B, int Hin,
                           int Win, int fuse_pool, cudaStream_t stream) {
  static const int variant = [] { const char* e = getenv("SX_STEM_VARIANT"); return e ? atoi(e) : 0; }();
  if (fuse_pool) {
    // measured alone at B = 256 (profiles/exp_stem.py, r03): 2 epilogue sets 0.187 ms, 3 sets 0.162 ms; the MMAs alone
    // (SX_STEM_DEBUG=5) take 0.143 ms -- K = 16 per instruction is paced by the 6 KB of operands it reads from shared memory
    if (variant == 1) return launch_stem_cfg<8, 2, true>(x, w_taps, bias, out, B, Hin, Win, stream);
    if (variant == 2) return launch_stem_cfg<6, 4, true>(x, w_taps, bias, out, B, Hin, Win, stream);
    return launch_stem_cfg<8, 3, true>(x, w_taps, bias, out, B, Hin, Win, stream);
  }
  return launch_stem_cfg<8, 3, false>(x, w_taps, bias, out, B, Hin, Win, stream);
}

}  // namespace tc
}  // namespace sx
