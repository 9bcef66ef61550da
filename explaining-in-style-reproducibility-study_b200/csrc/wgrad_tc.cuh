// wgrad_tc.cuh -- Conv2DMod weight gradient on the 5th-gen tensor cores (sm_100a), bf16 operands, fp32 accumulation.
//
//   grad_W[o, i, tap] = sum_{b, y, x} gz[b, o, y, x] * xm[b, i, y + dy, x + dx]            (conv_bwd.cuh)
//
// GEMM view: M = Co, N = Ci, K = pixels.  The trick that keeps this an ordinary K-major tcgen05 GEMM: both operands are
// kept in the reference's NCHW layout (bf16 copies, already scaled by d / (style+1)), so the K axis -- 64 consecutive
// pixels of one (sample, channel) plane, a BW x BH box with BW * BH = 64 -- is contiguous per channel.  One 4-D TMA box
// {BW, BH, channels, 1} per operand lands in shared memory as [channel][64 pixels] = 128-byte rows under SWIZZLE_128B, exactly
// the K-major tile layout conv_tc.cuh feeds to tcgen05.mma.  The tap's row offset dy is a coordinate offset of the xm box (TMA's
// out-of-bounds zero fill is the convolution padding, and pads Co / Ci up to the tile).  The column offset dx cannot be one:
// x is the innermost (contiguous) dimension and TMA needs a 16-byte aligned start there (an odd pixel offset raises an
// illegal-instruction fault), so the prep kernel writes xm three times, shifted by dx = -1, 0, +1 with zero padding, and
// the kernel picks the copy (one tensor map each).  No MN-major descriptors, no transposes.
// Work item = (Co tile of 128, Ci tile of BLOCK_N, tap, pixel split): a CTA accumulates its split's pixel chunks in TMEM and
// writes an fp32 partial [split][tap][Co][Ci]; wgrad_reduce_kernel (conv_bwd.cuh) combines the splits in a fixed order.
// Warp roles as in conv_tc.cuh: 0 = TMA producer, 1 = TMEM alloc + elected-thread MMA issue, 2..5 = epilogue.
#pragma once

#include "common.cuh"
#include "conv_tc.cuh"

namespace sx {
namespace tc {

struct WgradTcParams {
  int B, H, W, Ci, Co, KS;
  int BW, BH;                 // pixel box of one K chunk: BW * BH == 64
  int tiles_o, tiles_i, splits;
  int chunks_total;           // B * H * W / 64
  int chunks_per_split;
  int rows_a, rows_b;         // channel rows of the two TMA boxes: min(tile, channels) -- a box never exceeds the tensor; the tile
                              // rows it leaves untouched hold stale shared memory and only feed output rows / columns that are
                              // never stored (a GEMM row depends on its own operand row only)
  float* partial;             // [splits][taps][Co][Ci]
};

template <int BLOCK_N, int STAGES>
struct WgradCfg {
  static constexpr int kABytes = BLOCK_M * 64 * 2;
  static constexpr int kBBytes = BLOCK_N * 64 * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;
  static constexpr size_t smem_bytes = 1024 + (size_t)STAGES * kStageBytes + 256;
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_g,
                                                               const __grid_constant__ CUtensorMap tmap_x0,
                                                               const __grid_constant__ CUtensorMap tmap_x1,
                                                               const __grid_constant__ CUtensorMap tmap_x2, const WgradTcParams p) {
  using Cfg = WgradCfg<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp_id = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work item
  const int taps = p.KS * p.KS;
  int w = blockIdx.x;
  const int split = w % p.splits; w /= p.splits;
  const int tap = w % taps; w /= taps;
  const int ti = w % p.tiles_i;
  const int to = w / p.tiles_i;
  const int o0 = to * BLOCK_M, i0 = ti * BLOCK_N;
  const int pad = (p.KS - 1) / 2;
  const int dy = tap / p.KS - pad, dx = tap % p.KS - pad;
  // xm copy shifted by dx (k = 1: only the unshifted copy exists and it is tmap_x0)
  const CUtensorMap* tmap_x = p.KS == 1 ? &tmap_x0 : (dx < 0 ? &tmap_x0 : (dx == 0 ? &tmap_x1 : &tmap_x2));
  const int q0 = split * p.chunks_per_split;
  const int q1 = min(q0 + p.chunks_per_split, p.chunks_total);
  const int num_kb = q1 - q0;                       // >= 1 by construction of the splits
  const int cpr = p.W / p.BW;                       // chunks per image row band
  const int cpi = (p.H / p.BH) * cpr;               // chunks per image

  if (warp_id == 0 && lane == 0) {
    prefetch_tmap(&tmap_g);
    prefetch_tmap(tmap_x);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  } else if (warp_id == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_id == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      const int q = q0 + kb;
      const int b = q / cpi;
      const int r = q - b * cpi;
      const int y0 = (r / cpr) * p.BH, x0 = (r % cpr) * p.BW;
      mbar_wait(&empty_bar[stage], phase ^ 1, 10);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)((p.rows_a + p.rows_b) * 128));
        tma_load_4d(smem_a + stage * Cfg::kABytes, &tmap_g, &full_bar[stage], x0, y0, o0, b);
        tma_load_4d(smem_b + stage * Cfg::kBBytes, tmap_x, &full_bar[stage], x0, y0 + dy, i0, b);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp_id == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full_bar[stage], phase, 11);
      tc_fence_after();
      const uint64_t da = make_smem_desc<64>(smem_u32(smem_a + stage * Cfg::kABytes));
      const uint64_t db = make_smem_desc<64>(smem_u32(smem_b + stage * Cfg::kBBytes));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 64 / UMMA_K; ++k)
          umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(tmem_full_bar);
    __syncwarp();
  } else {
    // ===================== epilogue: TMEM -> fp32 partial[split][tap][o][i] =====================
    const int q = warp_id & 3;
    const int o = o0 + q * 32 + lane;
    float* dst = p.partial + (((size_t)split * taps + tap) * p.Co + o) * p.Ci + i0;
    mbar_wait(tmem_full_bar, 0, 12);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      if (o < p.Co) {
        if (i0 + c0 + 32 <= p.Ci && (p.Ci & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                    __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (i0 + c0 + j < p.Ci) dst[c0 + j] = __uint_as_float(v[j]);
        }
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

inline bool wgrad_tc_supported(int Ci, int Co, int H, int W, int KS) {
  auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  // W >= 64: the K chunk is then one 64-pixel (128-byte) run of a row.  Narrower maps would need a BW x BH box with a
  // 2*BW-byte inner extent under SWIZZLE_128B, which faults on hardware (measured: illegal memory access at W = 32) --
  // those layers keep the FFMA wgrad.
  return (KS == 1 || KS == 3) && H == W && pow2(W) && W >= 64 && Ci >= 1 && Co >= 1;
}

inline int wgrad_tc_block_n(int Ci) { return Ci > 128 ? 256 : (Ci > 64 ? 128 : (Ci > 32 ? 64 : 32)); }

// number of pixel splits: fill about two waves of CTAs, keep at least 8 K chunks (512 pixels) per split
inline int wgrad_tc_splits(int B, int Ci, int Co, int H, int W, int KS) {
  const int bn = wgrad_tc_block_n(Ci);
  const long long tiles = (long long)((Co + BLOCK_M - 1) / BLOCK_M) * ((Ci + bn - 1) / bn) * KS * KS;
  const long long chunks = (long long)B * H * W / 64;
  long long s = (2LL * num_sms() + tiles - 1) / tiles;
  const long long max_s = chunks / 8 > 0 ? chunks / 8 : 1;
  if (s > max_s) s = max_s;
  if (s > 1024) s = 1024;
  if (s < 1) s = 1;
  // no empty split: recompute from the rounded-up chunk count per split
  const long long cps = (chunks + s - 1) / s;
  return (int)((chunks + cps - 1) / cps);
}

template <int BLOCK_N, int STAGES>
int launch_wgrad_tc_cfg(const CUtensorMap& tg, const CUtensorMap* tx, const WgradTcParams& p, cudaStream_t stream) {
  using Cfg = WgradCfg<BLOCK_N, STAGES>;
  auto kern = wgrad_tc_kernel<BLOCK_N, STAGES>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    SX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes));
    configured = true;
  }
  const long long grid = (long long)p.tiles_o * p.tiles_i * p.KS * p.KS * p.splits;
  SX_REQUIRE(grid <= 0x7fffffffLL, "wgrad_tc: grid too large");
  kern<<<(unsigned)grid, NUM_THREADS, Cfg::smem_bytes, stream>>>(tg, tx[0], tx[1], tx[2], p);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

// gz: [B,Co,H,W] bf16 (grad_out * d); xm: [KS][B,Ci,H,W] bf16, copy s = x * (style+1) shifted by dx = s - (KS-1)/2 along x with
// zero padding (nchw_scale_shift_bf16_kernel); partial: [splits][KS*KS][Co][Ci] fp32
inline int launch_wgrad_tc(const __nv_bfloat16* gz, const __nv_bfloat16* xm, float* partial, int B, int Ci, int Co, int H, int W,
                           int KS, int splits, cudaStream_t stream) {
  if (!wgrad_tc_supported(Ci, Co, H, W, KS))
    return fail(SX_EUNSUPPORTED, "wgrad_tc: unsupported shape Ci=%d Co=%d H=%d W=%d k=%d (need square power-of-two H=W>=64)", Ci, Co, H, W, KS);
  if (B == 0) return SX_OK;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(SX_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  WgradTcParams p;
  p.B = B; p.H = H; p.W = W; p.Ci = Ci; p.Co = Co; p.KS = KS;
  p.BW = W < 64 ? W : 64;
  p.BH = 64 / p.BW;
  const int bn = wgrad_tc_block_n(Ci);
  p.tiles_o = (Co + BLOCK_M - 1) / BLOCK_M;
  p.tiles_i = (Ci + bn - 1) / bn;
  p.chunks_total = (int)((long long)B * H * W / 64);
  p.splits = splits;
  p.chunks_per_split = (p.chunks_total + splits - 1) / splits;
  p.partial = partial;
  p.rows_a = Co < BLOCK_M ? Co : BLOCK_M;
  p.rows_b = Ci < bn ? Ci : bn;
  SX_REQUIRE((long long)(splits - 1) * p.chunks_per_split < p.chunks_total, "wgrad_tc: empty pixel split");
  CUtensorMap tg, tx[3];
  auto make = [&](CUtensorMap* tm, const __nv_bfloat16* ptr, int C, int rows) -> int {
    cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)W * 2, (cuuint64_t)H * W * 2, (cuuint64_t)C * H * W * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.BW, (cuuint32_t)p.BH, (cuuint32_t)rows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(ptr), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SX_ECUDA, "cuTensorMapEncodeTiled(wgrad) failed: %d (B=%d C=%d H=%d W=%d rows=%d)", (int)r, B, C, H, W, rows);
    return SX_OK;
  };
  SX_TRY(make(&tg, gz, Co, p.rows_a));
  for (int s = 0; s < 3; ++s) SX_TRY(make(&tx[s], xm + (size_t)(s < KS ? s : 0) * B * Ci * H * W, Ci, p.rows_b));
  if (bn == 256) return launch_wgrad_tc_cfg<256, 4>(tg, tx, p, stream);
  if (bn == 128) return launch_wgrad_tc_cfg<128, 6>(tg, tx, p, stream);
  if (bn == 64) return launch_wgrad_tc_cfg<64, 6>(tg, tx, p, stream);
  return launch_wgrad_tc_cfg<32, 6>(tg, tx, p, stream);
}

}  // namespace tc
}  // namespace sx
