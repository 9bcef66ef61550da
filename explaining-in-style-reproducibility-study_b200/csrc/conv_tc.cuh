// conv_tc.cuh -- bf16 implicit-GEMM 3x3 convolution on the 5th-gen tensor cores (sm_100a):
// TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory -> tcgen05.mma (kind::f16, fp32 accumulators in
// TMEM) -> tcgen05.ld epilogue (demodulate, noise, leaky-ReLU, next-layer modulation) -> NHWC bf16 stores.
//
// GEMM view (Conv2DMod.forward ST:647-667 with the modulation moved onto the activations):
//   D[m, o] = sum_{tap, ci} A[m, (tap, ci)] * Wk[o, (tap, ci)]
//   m  = one output pixel; a CTA tile is 128 pixels = a (BB x BH x BW) box of the NHWC activation tensor
//   A  = the same box shifted by the tap offset (dy, dx); TMA zero-fills out-of-bounds pixels, which IS the
//        conv's zero padding -- no im2col buffer ever exists
//   Wk = weights packed [Co][tap*Ci + ci] (K-major), shared by the whole batch (style lives in A / epilogue)
// One K block = BLOCK_K channels of one tap.  Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator +
// single-thread MMA issuer, warps 2..5 = epilogue (one TMEM lane quarter each).
// PERSISTENT: one CTA per SM walks the (M tile, N tile) list round-robin; the accumulator is double-buffered in TMEM
// (2 x BLOCK_N columns), so the epilogue of tile i drains slot i&1 while the MMAs of tile i+1 fill the other slot and
// the TMA ring keeps streaming across tile boundaries (bench: the one-tile-per-CTA form left the 256->256@32 layer at
// 66 % of the tensor peak because load -> MMA -> epilogue ran back to back inside each CTA).
#pragma once

#include "common.cuh"

namespace sx {
namespace tc {

constexpr int BLOCK_M = 128;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr long long kWatchdogCycles = 4000000000LL;  // ~2 s: a dead barrier traps instead of hanging the box

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// try_wait with a suspend-time hint: the thread sleeps IN HARDWARE until the phase completes or ~hint_ns elapse.
__device__ __forceinline__ uint32_t mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok;
}
// ncu (profiles/README.md): with a bare try_wait + clock64 watchdog in the loop, the ~25 waiting warps of the persistent
// kernels polled ~800 times per tile and burnt 50-70 % of the SM's issue slots (ISETP / BRA / IADD3 / CS2R / YIELD were
// the five most executed opcodes), starving the MMA-issuing and epilogue warps.  Now: one plain try_wait (the common,
// already-complete case), then hardware-suspended waits of up to 20 us each; the watchdog counts those (~2 s in total)
// (~2 s of wall clock, sampled every 64 wake-ups) instead of reading the clock per poll.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int which) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {
    if ((++spins & 63u) == 0) {   // the clock is read once per 64 wake-ups, not per poll
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kWatchdogCycles) {
        printf("stylex_b200 conv_tc: barrier %d timed out (block %d,%d thread %d)\n", which, blockIdx.x, blockIdx.y, threadIdx.x);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// One elected lane of a fully converged warp.  The TMA / MMA warps run their loops warp-uniformly and predicate only
// the issuing instruction with this: inside an `if (lane == 0)` region the compiler cannot prove that the descriptor
// operands of UTCHMMA / UBLKCP are warp-uniform and wraps every one of them in a waterfall loop
// (ELECT ... BRA.U.ANY, ~47 SASS instructions per filter tap: the single issuing thread then paces the narrow layers).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// K-major, swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of BLOCK_K bf16
// (128 B -> SWIZZLE_128B, 64 B -> SWIZZLE_64B), 8-row groups SBO bytes apart, version 1 (Blackwell).
template <int BLOCK_K>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  constexpr uint64_t row_bytes = BLOCK_K * 2;
  constexpr uint64_t sbo = 8 * row_bytes;                // bytes between 8-row core-matrix groups
  constexpr uint64_t layout = row_bytes == 128 ? 2 : 4;  // SWIZZLE_128B = 2, SWIZZLE_64B = 4
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((sbo >> 4) << 32) | (1ull << 46) | (layout << 61);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major, M=128, N.
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

// ---- epilogue math on 32 accumulator columns of one output pixel (shared by both tcgen05 kernels) ----------------
// All per-column operands come from shared-memory tables read as float4 (7 LDS.128 per 4 columns instead of 28 LDS.32:
// the epilogue of the narrow layers is issue-bound, and these loads were most of its instructions).
__device__ __forceinline__ void epi_chunk32(const uint32_t* __restrict__ v, const float* __restrict__ dd, const float* __restrict__ nw,
                                            const float* __restrict__ nb, const float* __restrict__ mm, float nz, int act,
                                            bool round_bf16, float* __restrict__ f, float* __restrict__ fr) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 d4 = *reinterpret_cast<const float4*>(dd + j);
    const float4 w4 = *reinterpret_cast<const float4*>(nw + j);
    const float4 b4 = *reinterpret_cast<const float4*>(nb + j);
    const float4 m4 = *reinterpret_cast<const float4*>(mm + j);
    const float dv[4] = {d4.x, d4.y, d4.z, d4.w}, wv[4] = {w4.x, w4.y, w4.z, w4.w};
    const float bv[4] = {b4.x, b4.y, b4.z, b4.w}, mv[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float t = __uint_as_float(v[j + k]) * dv[k];
      t += nz * wv[k] + bv[k];
      if (act) t = lrelu02(t);
      // round the raw activation to the storage type BEFORE the next layer's modulation: the AttFind suffix path
      // re-modulates the cached (bf16) raw tensor, and both paths must produce bit-identical conv inputs
      if (round_bf16) t = __bfloat162float(__float2bfloat16_rn(t));
      fr[j + k] = t;
      f[j + k] = t * mv[k];
    }
  }
}
// fused ToRGB partial sums over the same 32 columns: acc[c] += sum_j fr[j] * rw[c * stride + j]
__device__ __forceinline__ void rgb_chunk32(const float* __restrict__ fr, const float* __restrict__ rw, int stride, float* __restrict__ acc) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float a = acc[c];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 w4 = *reinterpret_cast<const float4*>(rw + c * stride + j);
      a = fmaf(fr[j], w4.x, a);
      a = fmaf(fr[j + 1], w4.y, a);
      a = fmaf(fr[j + 2], w4.z, a);
      a = fmaf(fr[j + 3], w4.w, a);
    }
    acc[c] = a;
  }
}

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 issue two fp32 lanes per instruction) ----------------------------
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint32_t bf16x2_rn(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// The generator's epilogue (leaky-ReLU on, bf16 NHWC output) on 32 accumulator columns, two columns per instruction:
//   t = acc * d + (nz * nw + nb);  t = max(t, 0.2 t);  raw = bf16(t);  out = bf16(raw * m);  rgb_c += raw * rw_c
// Same values as epi_chunk32 (the fused multiply-adds are the ones the compiler formed there; max(t, 0.2t) == lrelu);
// the ToRGB partial sums run in two lanes (even / odd columns) that the caller adds at the end.
// om / orw: 16 packed bf16x2 words each (modulated output, raw output).
template <bool RGB>
__device__ __forceinline__ void epi_fast32(const uint32_t* __restrict__ v, const float* __restrict__ dd, const float* __restrict__ nw,
                                           const float* __restrict__ nb, const float* __restrict__ mm, float nz,
                                           uint32_t* __restrict__ om, uint32_t* __restrict__ orw, const float* __restrict__ rw,
                                           int rw_stride, uint64_t* __restrict__ racc) {
  const uint64_t nz2 = pk2(nz, nz), c02 = pk2(0.2f, 0.2f);
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 d4 = *reinterpret_cast<const float4*>(dd + j);
    const float4 w4 = *reinterpret_cast<const float4*>(nw + j);
    const float4 b4 = *reinterpret_cast<const float4*>(nb + j);
    const float4 m4 = *reinterpret_cast<const float4*>(mm + j);
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
    if (RGB) {
      r0 = *reinterpret_cast<const float4*>(rw + j);
      r1 = *reinterpret_cast<const float4*>(rw + rw_stride + j);
      r2 = *reinterpret_cast<const float4*>(rw + 2 * rw_stride + j);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = j + 2 * h;
      const uint64_t d2 = h ? pk2(d4.z, d4.w) : pk2(d4.x, d4.y);
      const uint64_t w2 = h ? pk2(w4.z, w4.w) : pk2(w4.x, w4.y);
      const uint64_t b2 = h ? pk2(b4.z, b4.w) : pk2(b4.x, b4.y);
      const uint64_t m2 = h ? pk2(m4.z, m4.w) : pk2(m4.x, m4.y);
      const uint64_t t2 = fma2(pk2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])), d2, fma2(nz2, w2, b2));
      const uint64_t l2 = mul2(t2, c02);
      float t0, t1, l0, l1;
      upk2(t2, t0, t1);
      upk2(l2, l0, l1);
      const uint32_t raw = bf16x2_rn(fmaxf(t0, l0), fmaxf(t1, l1));
      orw[c >> 1] = raw;
      const uint64_t fr2 = pk2(__uint_as_float(raw << 16), __uint_as_float(raw & 0xffff0000u));
      float f0, f1;
      upk2(mul2(fr2, m2), f0, f1);
      om[c >> 1] = bf16x2_rn(f0, f1);
      if (RGB) {
        racc[0] = fma2(fr2, h ? pk2(r0.z, r0.w) : pk2(r0.x, r0.y), racc[0]);
        racc[1] = fma2(fr2, h ? pk2(r1.z, r1.w) : pk2(r1.x, r1.y), racc[1]);
        racc[2] = fma2(fr2, h ? pk2(r2.z, r2.w) : pk2(r2.x, r2.y), racc[2]);
      }
    }
  }
}

struct ConvTcParams {
  int B, H, W, Ci, Co, KS;
  int BW, BH, BB;            // pixel box of one M tile: BB*BH*BW == 128
  int tiles_x, tiles_y;      // W/BW, H/BH
  int tiles_m;               // tiles_b * tiles_y * tiles_x M tiles; the persistent grid walks tiles_m * (Co / BLOCK_N) tiles
  ConvEpilogue ep;
};

template <int BLOCK_N, int BLOCK_K, int STAGES>
struct TcConfig {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSlotCols = BLOCK_N < 32 ? 32 : BLOCK_N;  // one accumulator slot
  static constexpr int kTmemCols = 2 * kSlotCols;                // two slots; power of two, 64..512
  // dynamic smem: [1024 align slack][stages][barriers 256B][tables]
  static size_t smem_bytes(int BB) {
    return 1024 + (size_t)STAGES * kStageBytes + 256 + (size_t)(2 + 2 * BB + 3) * BLOCK_N * sizeof(float);
  }
};

template <int BLOCK_N, int BLOCK_K, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS) conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                              const __grid_constant__ CUtensorMap tmap_b,
                                                              const ConvTcParams p) {
  using Cfg = TcConfig<BLOCK_N, BLOCK_K, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  // offset arithmetic on the __shared__ array (not an integer round trip) keeps the shared address space known to the
  // compiler: the table / source-box accesses below compile to LDS/STS instead of generic LD/ST (ncu: long_scoreboard)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;     // [2] accumulator slot complete (MMA -> epilogue)
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2] accumulator slot drained (epilogue -> MMA), 128 arrivals
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* s_nw = reinterpret_cast<float*>(smem + STAGES * Cfg::kStageBytes + 256);
  float* s_nb = s_nw + BLOCK_N;
  float* s_d = s_nb + BLOCK_N;            // [BB][BLOCK_N] demod coefficients
  float* s_m = s_d + p.BB * BLOCK_N;      // [BB][BLOCK_N] next-layer (style+1)
  float* s_rgbw = s_m + p.BB * BLOCK_N;   // [3][BLOCK_N] fused-ToRGB weights of the tile's sample (BB == 1 only)

  const int warp_id = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile list: t -> (M tile, N tile) with the N tiles of one M tile adjacent (CTAs of one wave share the A boxes in L2)
  const int tiles_per_b = p.tiles_x * p.tiles_y;
  const int tiles_n = p.Co / BLOCK_N;
  const int total_tiles = p.tiles_m * tiles_n;
  const int kc_per_tap = p.Ci / BLOCK_K;
  const int num_kb = p.KS * p.KS * kc_per_tap;
  const int pad = (p.KS - 1) / 2;

  if (warp_id == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 128);
    }
    fence_barrier_init();
  } else if (warp_id == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_id == 0) {
    // ===================== TMA producer: the ring runs straight through tile boundaries =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int tm = tile / tiles_n, n0 = (tile - tm * tiles_n) * BLOCK_N;
      const int tb = tm / tiles_per_b;
      const int tr = tm - tb * tiles_per_b;
      const int ty = tr / p.tiles_x, tx = tr - ty * p.tiles_x;
      const int x0 = tx * p.BW, y0 = ty * p.BH, b0 = tb * p.BB;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int tap = kb / kc_per_tap;
        const int c0 = (kb - tap * kc_per_tap) * BLOCK_K;
        const int dy = tap / p.KS - pad, dx = tap % p.KS - pad;
        mbar_wait(&empty_bar[stage], phase ^ 1, 0);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_4d(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], c0, x0 + dx, y0 + dy, b0);
          tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], tap * p.Ci + c0, n0);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp_id == 1) {
    // ===================== MMA issuer (one elected thread) =====================
    constexpr uint32_t idesc = make_idesc(BLOCK_N);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int slot = it & 1;
      // the slot's previous tenant (tile it-2) must have been drained; the first use of each slot passes at once
      mbar_wait(&tmem_empty_bar[slot], (uint32_t)(((it >> 1) & 1) ^ 1), 3);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(slot * Cfg::kSlotCols);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase, 1);
        tc_fence_after();
        const uint64_t da = make_smem_desc<BLOCK_K>(smem_u32(smem_a + stage * Cfg::kABytes));
        const uint64_t db = make_smem_desc<BLOCK_K>(smem_u32(smem_b + stage * Cfg::kBBytes));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance the start address by k * 16 bf16 = 32 bytes inside the swizzled row (>>4 -> +2)
            umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs have read it
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tmem_full_bar[slot]);  // accumulator of this tile complete
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const ConvEpilogue& ep = p.ep;
    const int et = threadIdx.x - 64;  // 0..127
    const bool fuse_rgb = ep.rgb_style != nullptr;   // host guarantees BB == 1 and a single N tile
    const int q = warp_id & 3;          // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;        // tile row = TMEM lane
    const int xx = r % p.BW;
    const int yy = (r / p.BW) % p.BH;
    const int bb = r / (p.BW * p.BH);
    const float* dd = s_d + bb * BLOCK_N;
    const float* mm = s_m + bb * BLOCK_N;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int slot = it & 1;
      const int tm = tile / tiles_n, n0 = (tile - tm * tiles_n) * BLOCK_N;
      const int tb = tm / tiles_per_b;
      const int tr = tm - tb * tiles_per_b;
      const int ty = tr / p.tiles_x, tx = tr - ty * p.tiles_x;
      const int x0 = tx * p.BW, y0 = ty * p.BH, b0 = tb * p.BB;

      // per-tile operand tables (the previous tile's readers are done: they passed the barrier at the end of the loop body)
      for (int i = et; i < BLOCK_N; i += 128) {
        s_nw[i] = ep.noise ? __ldg(ep.noise_w + n0 + i) : 0.f;
        s_nb[i] = ep.noise ? __ldg(ep.noise_b + n0 + i) : 0.f;
      }
      for (int i = et; i < p.BB * BLOCK_N; i += 128) {
        const int tbb = i / BLOCK_N, o = i - tbb * BLOCK_N;
        const int b = b0 + tbb;
        const bool ok = b < p.B;
        s_d[i] = (ok && ep.dcoef) ? __ldg(ep.dcoef + (long long)b * ep.dcoef_stride + n0 + o) : 1.f;
        s_m[i] = (ok && ep.next_style) ? __ldg(ep.next_style + (long long)b * ep.next_style_stride + n0 + o) + 1.f : 1.f;
      }
      if (fuse_rgb) {
        for (int i = et; i < 3 * BLOCK_N; i += 128) {
          const int o = i % BLOCK_N;
          s_rgbw[i] = (__ldg(ep.rgb_style + (long long)b0 * ep.rgb_style_stride + o) + 1.f) * __ldg(ep.rgb_w + (i / BLOCK_N) * p.Co + o);
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only: tables visible

      const int b = b0 + bb, y = y0 + yy, x = x0 + xx;
      const bool valid = b < p.B;
      float nz = 0.f;
      if (valid && ep.noise) {
        const int S = ep.noise_size;
        nz = __ldg(ep.noise + (long long)(ep.noise_batch == 1 ? 0 : b) * S * S + (long long)x * S + y);
      }
      const long long pix = ((long long)b * p.H + y) * p.W + x;
      float rgb_acc[3] = {0.f, 0.f, 0.f};
      float* rgb_dst = nullptr;
      if (fuse_rgb && valid) {
        rgb_dst = ep.rgb_out + ((long long)b * 3 * p.H + y) * p.W + x;
        if (ep.rgb_accumulate) {
#pragma unroll
          for (int c = 0; c < 3; ++c) rgb_acc[c] = __ldg(rgb_dst + (long long)c * p.H * p.W);
        }
      }

      mbar_wait(&tmem_full_bar[slot], (uint32_t)((it >> 1) & 1), 2);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * Cfg::kSlotCols);
      // generator layers (leaky-ReLU, bf16 NHWC output): packed fp32x2 arithmetic, two columns per instruction
      const bool fast = ep.act && !ep.out_nchw_f32;
      uint64_t racc[3] = {0ull, 0ull, 0ull};
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_d + (uint32_t)c0, v);
        tmem_ld_wait();
        if (fast) {
          uint32_t om[16], orw[16];
          if (fuse_rgb) epi_fast32<true>(v, dd + c0, s_nw + c0, s_nb + c0, mm + c0, nz, om, orw, s_rgbw + c0, BLOCK_N, racc);
          else epi_fast32<false>(v, dd + c0, s_nw + c0, s_nb + c0, mm + c0, nz, om, orw, nullptr, 0, racc);
          if (valid && ep.out) {
            uint4* out = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + pix * p.Co + n0 + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) out[j] = make_uint4(om[4 * j], om[4 * j + 1], om[4 * j + 2], om[4 * j + 3]);
          }
          if (valid && ep.out_raw) {
            uint4* out = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out_raw) + pix * p.Co + n0 + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) out[j] = make_uint4(orw[4 * j], orw[4 * j + 1], orw[4 * j + 2], orw[4 * j + 3]);
          }
          continue;
        }
        float f[32], fr[32];
        epi_chunk32(v, dd + c0, s_nw + c0, s_nb + c0, mm + c0, nz, ep.act, !ep.out_nchw_f32, f, fr);
        if (fuse_rgb) rgb_chunk32(fr, s_rgbw + c0, BLOCK_N, rgb_acc);
        if (valid && ep.out) {
          if (ep.out_nchw_f32) {
            float* out = reinterpret_cast<float*>(ep.out);
#pragma unroll
            for (int j = 0; j < 32; ++j) out[(((long long)b * p.Co + n0 + c0 + j) * p.H + y) * p.W + x] = f[j];
          } else {
            __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(ep.out) + pix * p.Co + n0 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 pk;
              pack(f + j, pk);
              *reinterpret_cast<uint4*>(out + j) = pk;
            }
          }
        }
        if (valid && ep.out_raw) {
          __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(ep.out_raw) + pix * p.Co + n0 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 pk;
            pack(fr + j, pk);
            *reinterpret_cast<uint4*>(out + j) = pk;
          }
        }
      }
      // all tcgen05.ld of this slot have completed (wait::ld above): hand the slot back to the MMA warp
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[slot]);
      if (rgb_dst) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float lo, hi;
          upk2(racc[c], lo, hi);   // zero unless the packed epilogue ran (even / odd column partial sums)
          rgb_dst[(long long)c * p.H * p.W] = rgb_acc[c] + (lo + hi);
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // every epilogue thread is done with this tile's tables
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
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

inline bool tc_shape_supported(int Ci, int Co, int H, int W, int KS) {
  auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  if (KS != 3) return false;
  if (Ci % 32 != 0) return false;
  if (!(Co == 32 || Co == 64 || Co == 128 || Co % 256 == 0)) return false;
  if (H != W || !pow2(W) || W < 4) return false;
  return true;
}

template <int BLOCK_N, int BLOCK_K, int STAGES>
int launch_conv_tc_cfg(const __nv_bfloat16* x, const __nv_bfloat16* wk, ConvTcParams p, cudaStream_t stream) {
  using Cfg = TcConfig<BLOCK_N, BLOCK_K, STAGES>;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(SX_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  const CUtensorMapSwizzle swz = BLOCK_K == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMap ta, tb;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)p.Ci, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    cuuint64_t gstr[3] = {(cuuint64_t)p.Ci * 2, (cuuint64_t)p.W * p.Ci * 2, (cuuint64_t)p.H * p.W * p.Ci * 2};
    cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)p.BW, (cuuint32_t)p.BH, (cuuint32_t)p.BB};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(x), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SX_ECUDA, "cuTensorMapEncodeTiled(A) failed: %d (B=%d H=%d W=%d Ci=%d)", (int)r, p.B, p.H, p.W, p.Ci);
  }
  {
    const cuuint64_t ktot = (cuuint64_t)p.KS * p.KS * p.Ci;
    cuuint64_t gdim[2] = {ktot, (cuuint64_t)p.Co};
    cuuint64_t gstr[1] = {ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)BLOCK_N};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(wk), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SX_ECUDA, "cuTensorMapEncodeTiled(B) failed: %d (Co=%d K=%llu)", (int)r, p.Co, (unsigned long long)ktot);
  }
  auto kern = conv_tc_kernel<BLOCK_N, BLOCK_K, STAGES>;
  const size_t smem = Cfg::smem_bytes(p.BB);
  static size_t configured = 0;  // per instantiation
  if (smem > configured) {
    SX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int tiles_b = (p.B + p.BB - 1) / p.BB;
  p.tiles_m = tiles_b * p.tiles_y * p.tiles_x;
  const long long total = (long long)p.tiles_m * (p.Co / BLOCK_N);
  // persistent: one CTA per SM (each CTA owns 2 x BLOCK_N TMEM columns; co-resident CTAs of the narrow instantiations
  // would still fit in the 512 columns, but one per SM keeps the allocation unconditional)
  const unsigned grid = (unsigned)(total < (long long)num_sms() ? total : (long long)num_sms());
  kern<<<grid, NUM_THREADS, smem, stream>>>(ta, tb, p);
  SX_CHECK_LAUNCH();
  return SX_OK;
}

// x: NHWC bf16 [B,H,W,Ci] already modulated; wk: [Co][KS*KS*Ci] bf16.
inline int launch_conv_tc(const __nv_bfloat16* x, const __nv_bfloat16* wk, int B, int Ci, int Co, int H, int W, int KS,
                          const ConvEpilogue& ep, cudaStream_t stream) {
  if (!tc_shape_supported(Ci, Co, H, W, KS))
    return fail(SX_EUNSUPPORTED, "conv_tc: unsupported shape Ci=%d Co=%d H=%d W=%d k=%d (need k=3, Ci%%32==0, Co in {32,64,128,256n}, square pow2)",
                Ci, Co, H, W, KS);
  if (B == 0) return SX_OK;
  ConvTcParams p;
  p.B = B; p.H = H; p.W = W; p.Ci = Ci; p.Co = Co; p.KS = KS;
  p.BW = W < BLOCK_M ? W : BLOCK_M;
  p.BH = (BLOCK_M / p.BW) < H ? (BLOCK_M / p.BW) : H;
  p.BB = BLOCK_M / (p.BW * p.BH);
  p.tiles_x = W / p.BW;
  p.tiles_y = H / p.BH;
  p.ep = ep;
  const int bn = Co % 256 == 0 ? 256 : Co;
  if (ep.rgb_style && (p.BB != 1 || bn != Co))
    return fail(SX_EINVAL, "conv_tc: fused ToRGB needs one sample per M tile and a single N tile (H=%d Co=%d)", H, Co);
  if (!ep.out && !ep.rgb_style) return fail(SX_EINVAL, "conv_tc: no output requested");
  const int bk = Ci % 64 == 0 ? 64 : 32;
#define SX_TC_CASE(N, K, S) \
  if (bn == N && bk == K) return launch_conv_tc_cfg<N, K, S>(x, wk, p, stream);
  SX_TC_CASE(256, 64, 4)
  SX_TC_CASE(128, 64, 6)
  SX_TC_CASE(64, 64, 4)
  SX_TC_CASE(32, 64, 5)
  SX_TC_CASE(256, 32, 6)
  SX_TC_CASE(128, 32, 6)
  SX_TC_CASE(64, 32, 6)
  SX_TC_CASE(32, 32, 6)
#undef SX_TC_CASE
  return fail(SX_EUNSUPPORTED, "conv_tc: no kernel for BLOCK_N=%d BLOCK_K=%d", bn, bk);
}

}  // namespace tc
}  // namespace sx
