// conv_tc_halo.cuh -- persistent, halo-reusing variant of the tcgen05 3x3 convolution for the high-resolution,
// narrow layers (Co <= 128, H >= 16) where conv_tc_kernel is bound by L2->shared-memory traffic: it re-loads the
// 128-pixel activation tile once per tap (9x).  Here a CTA tile is 16 rows x 8 pixels; ONE TMA box of 18 x 10
// pixels (the tile plus its 1-pixel halo; out-of-bounds = zero = the conv padding) is loaded per 64-channel chunk
// and all 9 taps read it in place:
//
//   smem row of halo pixel (hy, hx)        = hy * 10 + hx                     (rows of BLOCK_K bf16, TMA-swizzled)
//   A operand of tap (ky, kx)              = rows (yy + ky) * 10 + (xx + kx),  yy < 16, xx < 8
//                                          = 16 groups of 8 consecutive rows, groups 10 rows apart
//   => UMMA smem descriptor: start = base + (ky*10 + kx) * row_bytes,  SBO = 10 * row_bytes
//
// The start address is then only row-aligned (not 1024 B aligned).  That is legal because the 128B/64B swizzle of
// both TMA and tcgen05.mma is a function of the shared-memory ADDRESS bits (XOR of bits [4,7) with bits [7,10)) --
// the same property the K-advance of +32 B inside a swizzled row relies on (conv_tc.cuh, verified on hardware).
//
// Persistent: grid = resident CTAs; each CTA walks tiles t = blockIdx.x + i * gridDim.x.  Two TMEM accumulators
// (2 x BLOCK_N columns) let the epilogue of tile i overlap the MMAs of tile i+1; the TMA producer runs ahead
// across tile boundaries.  For the last block(s) the whole 9-tap weight set of the layer (<= 72 KB) stays
// RESIDENT in shared memory for the CTA's lifetime; otherwise weights stream through their own ring.
#pragma once

#include <algorithm>
#include <cstring>

#include "conv_tc.cuh"

namespace sx {
namespace tc {

constexpr int HALO_BW = 8, HALO_BH = 16, HALO_W = HALO_BW + 2, HALO_H = HALO_BH + 2, HALO_ROWS = HALO_W * HALO_H;

// like make_smem_desc but with an explicit stride between 8-row groups
template <int BLOCK_K>
__device__ __forceinline__ uint64_t make_smem_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  constexpr uint64_t layout = BLOCK_K * 2 == 128 ? 2 : 4;  // SWIZZLE_128B = 2, SWIZZLE_64B = 4
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (layout << 61);
}

// fused-upsample variant (UPS): the conv input is the LOW-RESOLUTION tensor; a 10 x 6 pixel source box per channel chunk
// is TMA-loaded (un-swizzled) into its own ring and the producer warps write the bilinearly upsampled 18 x 10 halo box
// into the swizzled A stage (see the producer branch of the kernel).
constexpr int UPS_SRC_W = HALO_W / 2 + 1, UPS_SRC_H = HALO_H / 2 + 1;   // 6 x 10 source pixels
constexpr int UPS_S_STAGES = 2;   // default depth of the low-resolution source-box ring (HaloCfg::kSrcStages)

constexpr int halo_pow2_cols(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

// Warp roles: 0 = TMA producer, 1-2 = MMA issuers (1 also allocates TMEM), 3 idle, then SETS epilogue sets of four warps
// (one TMEM lane quarter each), then (UPS) UPS_WARPS upsample producers.  The layers this kernel serves are EPILOGUE-bound, not
// MMA-bound (Co <= 128: a 128 x 32 tile is 288 tensor cycles of MMA but ~1500 warp-instructions of epilogue), so the
// accumulator ring has SETS slots and every slot has its own epilogue set: set s drains tiles s, s + SETS, ... while the
// other sets drain theirs -- 4 * SETS warps hide the TMEM-load / shared-table / global-store latencies of each other.
// Every slot holds TPS consecutive tiles (a "super-tile": same sample, same tile row, x advancing by 8): an epilogue
// thread drains pixel r of all TPS tiles in one pass, so each per-channel table value it fetches from shared memory
// (a broadcast LDS: 128 threads read the same word) serves TPS pixels.  ncu: with TPS = 1 those table loads were 55-60 %
// of the shared-memory wavefronts of the narrow layers, whose L1/shared data pipe ran at 80-85 % -- the actual bound.
//
// PAR (column-parity form, for the Co <= 64 layers whose tcgen05.mma is paced by the shared-memory read of its A operand:
// 4 KB per M128 x K16 instruction whatever N is): a super-tile is 16 rows x 16 pixels and TMEM lane r = (row yy, coarse
// column xc) owns the TWO pixels x0 + 2 xc + delta.  out[2 xc + delta] = sum_kx W[kx] in[2 xc + delta + kx - 1]: the input
// column 2 xc + u - 1 (u = 0..3) feeds (delta, kx) = (0, u) and (1, u - 1), so ONE A operand per (ky, u) -- 4 instead of
// 6 per filter row -- serves both pixels: u = 1, 2 as a single N = 2 Co instruction whose B operand is two ADJACENT
// resident weight tiles (the tiles of a filter row are stored kx = 2, 1, 0), u = 0 / 3 as N = Co instructions into the
// delta = 0 / 1 half of the accumulator.  The stride-2 column sampling comes from TMA (elementStrides = {1, 2, 1, 1}:
// two 18 x 9 boxes per stage, odd columns x0 - 1 + 2 m and even columns x0 + 2 m) or, with the fused upsample, from the
// producer warps, which already compute even and odd output columns separately.
template <int BLOCK_N, int BLOCK_K, int A_STAGES, int B_STAGES, bool RESIDENT_B, bool UPS, int SETS, int UPS_WARPS, int TPS, int NMMA,
          bool PAR = false>
struct HaloCfg {
  static_assert(!PAR || (TPS == 2 && RESIDENT_B), "the parity form owns two pixels per lane and keeps the weights resident");
  static_assert(SETS >= 2 && SETS <= 4, "2..4 accumulator slots / epilogue sets");
  static_assert(TPS == 1 || TPS == 2, "tiles per accumulator slot (W >= 16 guarantees two tiles per tile row)");
  static_assert(NMMA == 1 || NMMA == 2, "one or two MMA-issuing warps");
  // With two MMA warps, warp m issues super-tiles m, m + 2, ...; the activation ring (and the weight ring, when weights
  // stream) is split into one part per warp, filled by the producers according to the super-tile's parity.  An mbarrier
  // parity wait can only tell adjacent phases apart, so every ring needs a single in-order consumer.
  static_assert(A_STAGES % NMMA == 0 && (RESIDENT_B || B_STAGES % NMMA == 0), "rings are split between the MMA warps");
  static constexpr int kAHalf = A_STAGES / NMMA, kBHalf = B_STAGES / NMMA;
  static_assert(!UPS || UPS_WARPS == 4 || UPS_WARPS == 8, "4 (8 channels per thread) or 8 (4 channels per thread) producer warps");
  static constexpr int kRowBytes = BLOCK_K * 2;
  static constexpr int kEpiThreads = 128 * SETS;
  static constexpr int kUpsThreads = UPS ? 32 * UPS_WARPS : 0;
  static constexpr int kFrontThreads = 128;   // warp 0 TMA, warps 1..NMMA MMA issuers, the rest of the first four idle (sets stay 4-aligned)
  static constexpr int kThreads = kFrontThreads + kEpiThreads + kUpsThreads;
  static constexpr int kSrcW = PAR ? HALO_BW + 2 : UPS_SRC_W;           // low-res source columns: 10 for a 16-pixel super-tile
  static constexpr int kSrcBytes = kSrcW * UPS_SRC_H * kRowBytes;       // one source box (7680 B at BLOCK_K = 64; PAR: 12800 B)
  // The source boxes are the fused-upsample layers' ONLY global loads: their ring depth is the memory-level parallelism of
  // the CTA (2 boxes of 7.7 KB in flight per SM could not cover the DRAM latency; the plain 32->32 layer gained 18 % from
  // 4 -> 8 stages, profiles/README.md r02).  Measured for the fused-upsample layers themselves (r02 r2o): 6 boxes instead of 2
  // change nothing on 64->32 (0.852 ms either way: that layer is issue-bound), and trading weight-ring or resident-weight
  // space for them makes 256->128 and 128->64 slower (0.562 -> 0.586, 0.591 -> 1.025 ms): 2 it stays.
  static constexpr int kSrcStages = UPS_S_STAGES;
  static constexpr int kSrcRegion = UPS ? kSrcStages * kSrcBytes : 0;
  static constexpr int kHaloW = PAR ? HALO_BW + 1 : HALO_W;             // pixel rows per halo row in a box (PAR: 9 per lattice)
  static constexpr int kBoxRows = kHaloW * HALO_H;
  static constexpr int kBoxBytes = (kBoxRows * kRowBytes + 1023) / 1024 * 1024;   // every box 1024 B aligned (swizzle phase)
  static constexpr int kATx = (PAR ? 2 : 1) * kBoxRows * kRowBytes;     // bytes one stage receives
  static constexpr int kABytes = (PAR ? 2 : 1) * kBoxBytes;             // stage stride
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;           // weights of one (channel chunk, tap)
  static constexpr int kSlotCols = TPS * BLOCK_N;
  static constexpr int kTmemCols = halo_pow2_cols(SETS * kSlotCols);
  static_assert(SETS * kSlotCols <= 512, "TMEM has 512 columns");
  static constexpr int kBarBytes = 512;
  static_assert((2 * A_STAGES + 2 * B_STAGES + 2 * SETS + 2 * kSrcStages) * 8 + 8 <= kBarBytes, "barrier block overflow");
  static constexpr int kTableFloats = (2 + 10 * SETS) * BLOCK_N;  // nw, nb, then per set 2 slots x (d, m, 3 rgb rows)
  __host__ __device__ static size_t b_region(int num_b_tiles) { return (size_t)(RESIDENT_B ? num_b_tiles : B_STAGES) * kBBytes; }
  static size_t smem_bytes(int num_b_tiles) {
    return 1024 + (size_t)A_STAGES * kABytes + b_region(num_b_tiles) + kBarBytes + (size_t)kTableFloats * sizeof(float) + kSrcRegion;
  }
};

struct ConvHaloParams {
  int B, H, W, Ci, Co;
  int tiles_x, tiles_y, num_tiles;   // W/8, H/16, B*tiles_y*tiles_x
  int tx_shift, tpb_shift;           // log2(tiles_x), log2(tiles_x * tiles_y): H, W are powers of two
  int debug;                         // SX_HALO_DEBUG bitmask (bottleneck experiments; results are garbage when set):
                                     //   1 skip epilogue math/stores, 2 skip MMA issue, 4 skip activation TMA loads
  int kchunks, num_b_tiles;          // Ci/BLOCK_K, 9*kchunks
  int tps;                           // tiles per accumulator slot actually used: min(TPS, tiles per sample), a power of two
  unsigned long long* trace;         // SX_HALO_DEBUG_KNOBS builds: [8 roles][64 tiles] clock64 stamps of CTA (0,0); else null
  ConvEpilogue ep;
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

// The generator's epilogue on 4 accumulator columns of one pixel, two columns per instruction (see epi_fast32 in
// conv_tc.cuh for the arithmetic; identical values), with the per-channel table values already in registers.
// OUT: the modulated tensor is wanted (om), RAW: the raw one (orw).  om / orw: 2 packed bf16x2 words each.
struct EpiTab4 {
  float4 d, w, b, m, r0, r1, r2;
};
template <bool RGB, bool OUT, bool RAW>
__device__ __forceinline__ void epi_group4(const uint32_t* __restrict__ v, const EpiTab4& t, uint64_t nz2, uint32_t* __restrict__ om,
                                           uint32_t* __restrict__ orw, uint64_t* __restrict__ racc) {
  const uint64_t c02 = pk2(0.2f, 0.2f);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint64_t d2 = h ? pk2(t.d.z, t.d.w) : pk2(t.d.x, t.d.y);
    const uint64_t w2 = h ? pk2(t.w.z, t.w.w) : pk2(t.w.x, t.w.y);
    const uint64_t b2 = h ? pk2(t.b.z, t.b.w) : pk2(t.b.x, t.b.y);
    const uint64_t t2 = fma2(pk2(__uint_as_float(v[2 * h]), __uint_as_float(v[2 * h + 1])), d2, fma2(nz2, w2, b2));
    const uint64_t l2 = mul2(t2, c02);
    float t0, t1, l0, l1;
    upk2(t2, t0, t1);
    upk2(l2, l0, l1);
    if (RGB && !OUT && !RAW) {
      // the last block: nothing stores this feature map (only its ToRGB sum leaves the kernel), so there is no storage type
      // to round to -- the fp32 activation feeds the ToRGB accumulation directly (3 instructions fewer per channel pair,
      // and closer to the fp32 reference).  Full and suffix forwards run this same kernel: still bit-identical to each other.
      const uint64_t fr2 = pk2(fmaxf(t0, l0), fmaxf(t1, l1));
      racc[0] = fma2(fr2, h ? pk2(t.r0.z, t.r0.w) : pk2(t.r0.x, t.r0.y), racc[0]);
      racc[1] = fma2(fr2, h ? pk2(t.r1.z, t.r1.w) : pk2(t.r1.x, t.r1.y), racc[1]);
      racc[2] = fma2(fr2, h ? pk2(t.r2.z, t.r2.w) : pk2(t.r2.x, t.r2.y), racc[2]);
      continue;
    }
    const uint32_t raw = bf16x2_rn(fmaxf(t0, l0), fmaxf(t1, l1));
    if (RAW) orw[h] = raw;
    if (OUT || RGB) {
      const uint64_t fr2 = pk2(__uint_as_float(raw << 16), __uint_as_float(raw & 0xffff0000u));
      if (OUT) {
        const uint64_t m2 = h ? pk2(t.m.z, t.m.w) : pk2(t.m.x, t.m.y);
        float f0, f1;
        upk2(mul2(fr2, m2), f0, f1);
        om[h] = bf16x2_rn(f0, f1);
      }
      if (RGB) {
        racc[0] = fma2(fr2, h ? pk2(t.r0.z, t.r0.w) : pk2(t.r0.x, t.r0.y), racc[0]);
        racc[1] = fma2(fr2, h ? pk2(t.r1.z, t.r1.w) : pk2(t.r1.x, t.r1.y), racc[1]);
        racc[2] = fma2(fr2, h ? pk2(t.r2.z, t.r2.w) : pk2(t.r2.x, t.r2.y), racc[2]);
      }
    }
  }
}

// bf16x2 word -> packed fp32x2 (exact)
__device__ __forceinline__ uint64_t bf2_to_f2(uint32_t w) { return pk2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
__device__ __forceinline__ uint32_t f2_to_bf2(uint64_t v) {
  float lo, hi;
  upk2(v, lo, hi);
  return bf16x2_rn(lo, hi);
}

template <int BLOCK_N, int BLOCK_K, int A_STAGES, int B_STAGES, bool RESIDENT_B, bool FUSE_RGB, bool UPS, int SETS, int UPS_WARPS,
          int TPS, int NMMA, bool PAR = false>
__global__ void __launch_bounds__(HaloCfg<BLOCK_N, BLOCK_K, A_STAGES, B_STAGES, RESIDENT_B, UPS, SETS, UPS_WARPS, TPS, NMMA, PAR>::kThreads, 1)
conv_tc_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const ConvHaloParams p) {
  using Cfg = HaloCfg<BLOCK_N, BLOCK_K, A_STAGES, B_STAGES, RESIDENT_B, UPS, SETS, UPS_WARPS, TPS, NMMA, PAR>;
  static_assert(!UPS || BLOCK_K == 64, "the fused-upsample producer writes the SWIZZLE_128B layout");
  extern __shared__ uint8_t smem_raw[];
  // offset arithmetic on the __shared__ array (not an integer round trip) keeps the shared address space known to the
  // compiler: the table / source-box accesses below compile to LDS/STS instead of generic LD/ST (ncu: long_scoreboard)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + A_STAGES * Cfg::kABytes;
  uint8_t* after = smem_b + Cfg::b_region(p.num_b_tiles);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(after);
  uint64_t* a_empty = a_full + A_STAGES;
  uint64_t* b_full = a_empty + A_STAGES;          // [B_STAGES]  (resident: b_full[0] = "all weights landed")
  uint64_t* b_empty = b_full + B_STAGES;
  uint64_t* tmem_full = b_empty + B_STAGES;       // [SETS]
  uint64_t* tmem_empty = tmem_full + SETS;        // [SETS]
  uint64_t* s_full = tmem_empty + SETS;           // [UPS_S_STAGES]  (UPS only) source box landed
  uint64_t* s_empty = s_full + Cfg::kSrcStages;   // [kSrcStages]  (UPS only) producers are done with the source box
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(s_empty + Cfg::kSrcStages);
  float* s_nw = reinterpret_cast<float*>(after + Cfg::kBarBytes);
  float* s_nb = s_nw + BLOCK_N;
  float* s_tab = s_nb + BLOCK_N;          // [SETS][2 slots][d | m | rgb0 | rgb1 | rgb2][BLOCK_N]
  uint8_t* smem_src = reinterpret_cast<uint8_t*>(s_nw + Cfg::kTableFloats);   // (UPS only) [UPS_S_STAGES][10][6][BLOCK_K] bf16

  const int warp_id = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BLOCK_N;
  // Each CTA walks a CONTIGUOUS range of tiles (sample-major, then rows, then columns): the tile's sample -- and with it
  // the per-sample epilogue tables -- changes once per several hundred tiles instead of every other tile, and
  // consecutive tiles share halo columns in L2.
  // The unit of the walk is the super-tile (tps tiles: one accumulator slot).
  constexpr int tps = TPS;   // compile-time: the launcher guarantees tiles_x >= TPS
  const int num_super = p.num_tiles / tps;
  const int super_begin = (int)((long long)blockIdx.x * num_super / gridDim.x);
  const int super_end = (int)((long long)(blockIdx.x + 1) * num_super / gridDim.x);
  const int tile_begin = super_begin * tps, tile_end = super_end * tps;
  constexpr int tps_shift = TPS == 2 ? 1 : 0;
#ifdef SX_HALO_DEBUG_KNOBS   // bottleneck experiments only (profiles/README.md); never in the shipped library
  const int dbg = p.debug;
#else
  constexpr int dbg = 0;
#endif
#ifdef SX_HALO_DEBUG_KNOBS
#define SX_TRACE(role, idx)                                                                          \
  do {                                                                                               \
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && (idx) >= 0 && (idx) < 64) p.trace[(role) * 64 + (idx)] = clock64(); \
  } while (0)
#else
#define SX_TRACE(role, idx) do { } while (0)
#endif

  for (int i = threadIdx.x; i < BLOCK_N; i += Cfg::kThreads) {
    s_nw[i] = p.ep.noise ? __ldg(p.ep.noise_w + n0 + i) : 0.f;
    s_nb[i] = p.ep.noise ? __ldg(p.ep.noise_b + n0 + i) : 0.f;
  }
  if (warp_id == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < A_STAGES; ++s) { mbar_init(&a_full[s], UPS ? Cfg::kUpsThreads : 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < B_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < SETS; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 128); }
    if (UPS) for (int s = 0; s < Cfg::kSrcStages; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], Cfg::kUpsThreads); }
    fence_barrier_init();
  } else if (warp_id == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_id == 0) {
    // ===================== TMA producer (warp-uniform control flow, one elected lane issues) =====================
    {
      if (RESIDENT_B) {
        if (elect_one()) {
          mbar_arrive_expect_tx(&b_full[0], (uint32_t)(p.num_b_tiles * Cfg::kBBytes));
          for (int i = 0; i < p.num_b_tiles; ++i) {
            const int chunk = i / 9;
            int tap = i - chunk * 9;
            if (PAR) tap = (tap / 3) * 3 + (2 - tap % 3);   // the tiles of a filter row lie kx = 2, 1, 0 (see the MMA issuer)
            tma_load_2d(smem_b + (size_t)i * Cfg::kBBytes, &tmap_b, &b_full[0], tap * p.Ci + chunk * BLOCK_K, n0);
          }
        }
        __syncwarp();
      }
      // ring positions per half (index = tile parity); the source-box ring of the fused upsample is one in-order ring
      int as2[2] = {0, 0}, bs2[2] = {0, 0}, ss = 0;
      uint32_t aph2[2] = {0, 0}, bph2[2] = {0, 0}, sph = 0;
      for (int tile = tile_begin; tile < tile_end; tile += PAR ? 2 : 1) {   // PAR: one stage fill per 16-pixel super-tile
        const int b = tile >> p.tpb_shift;
        const int tr = tile - (b << p.tpb_shift);
        const int ty = tr >> p.tx_shift, tx = tr - (ty << p.tx_shift);
        const int x0 = tx * HALO_BW, y0 = ty * HALO_BH;
        const int rg = NMMA == 2 ? (((tile - tile_begin) >> tps_shift) & 1) : 0;
        int as = rg ? as2[1] : as2[0], bs = rg ? bs2[1] : bs2[0];
        uint32_t aph = rg ? aph2[1] : aph2[0], bph = rg ? bph2[1] : bph2[0];
        for (int chunk = 0; chunk < p.kchunks; ++chunk) {
          if (UPS) {
            // low-resolution source box of the halo box: rows y0/2-1 .. y0/2+8, columns x0/2-1 .. x0/2+4 (zero-filled
            // outside the image; the producers clamp instead, like torch's bilinear kernel)
            mbar_wait(&s_empty[ss], sph ^ 1, 10);
            if (elect_one()) {
              if (chunk == 0) SX_TRACE(0, tile - tile_begin);
              mbar_arrive_expect_tx(&s_full[ss], Cfg::kSrcBytes);   // (PAR: the tensor map's box is 10 columns wide)
              tma_load_4d(smem_src + ss * Cfg::kSrcBytes, &tmap_a, &s_full[ss], chunk * BLOCK_K, x0 / 2 - 1, y0 / 2 - 1, b);
            }
            __syncwarp();
            if (++ss == Cfg::kSrcStages) { ss = 0; sph ^= 1; }
          } else {
            const int st = rg * Cfg::kAHalf + as;
            mbar_wait(&a_empty[st], aph ^ 1, 10);
            if (elect_one()) {
              if (chunk == 0) SX_TRACE(0, tile - tile_begin);
              if (dbg & 4) {
                mbar_arrive(&a_full[st]);
              } else {
                mbar_arrive_expect_tx(&a_full[st], Cfg::kATx);
                tma_load_4d(smem_a + st * Cfg::kABytes, &tmap_a, &a_full[st], chunk * BLOCK_K, x0 - 1, y0 - 1, b);
                if (PAR)   // the map samples every other column: this box holds x0 - 1 + 2 m, the second one x0 + 2 m
                  tma_load_4d(smem_a + st * Cfg::kABytes + Cfg::kBoxBytes, &tmap_a, &a_full[st], chunk * BLOCK_K, x0, y0 - 1, b);
              }
            }
            __syncwarp();
            if (++as == Cfg::kAHalf) { as = 0; aph ^= 1; }
          }
          if (!RESIDENT_B) {
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
              const int st = rg * Cfg::kBHalf + bs;
              mbar_wait(&b_empty[st], bph ^ 1, 11);
              if (elect_one()) {
                mbar_arrive_expect_tx(&b_full[st], Cfg::kBBytes);
                tma_load_2d(smem_b + st * Cfg::kBBytes, &tmap_b, &b_full[st], tap * p.Ci + chunk * BLOCK_K, n0);
              }
              __syncwarp();
              if (++bs == Cfg::kBHalf) { bs = 0; bph ^= 1; }
            }
          }
        }
        if (rg) { as2[1] = as; aph2[1] = aph; bs2[1] = bs; bph2[1] = bph; }
        else { as2[0] = as; aph2[0] = aph; bs2[0] = bs; bph2[0] = bph; }
      }
    }
  } else if (warp_id >= 1 && warp_id <= NMMA) {
    // ===================== MMA issuers (warp-uniform control flow, one elected lane issues) =====================
    // Pipeline traces (profiles/README.md) showed ONE issuing warp pacing the narrow layers: tcgen05.mma issue blocks while
    // the (short) MMA queue is full, so the warp spends the tile's whole MMA time inside the issue loop and only then
    // runs its per-tile bookkeeping (barrier waits, fences, descriptors: 500-1000 cycles of dependent scalar code) with
    // the tensor pipe idle.  With NMMA = 2, warp m issues super-tiles m, m + 2, ... from its own part of the rings: while
    // one warp is blocked issuing, the other does its bookkeeping.
    {
      const int mw = warp_id - 1;
      constexpr uint32_t idesc = make_idesc(BLOCK_N);
      constexpr uint32_t idesc2 = make_idesc(2 * BLOCK_N);   // PAR: both pixels of a lane in one instruction
      constexpr uint32_t sbo = Cfg::kHaloW * Cfg::kRowBytes;
      if (RESIDENT_B) {
        mbar_wait(&b_full[0], 0, 12);
        tc_fence_after();
      }
      const uint32_t a_base0 = smem_u32(smem_a) + (uint32_t)(mw * Cfg::kAHalf * Cfg::kABytes);   // this warp's part of the A ring
      const uint32_t b_base0 = smem_u32(smem_b) + (RESIDENT_B ? 0u : (uint32_t)(mw * Cfg::kBHalf * Cfg::kBBytes));
      uint64_t* my_a_full = a_full + mw * Cfg::kAHalf;
      uint64_t* my_a_empty = a_empty + mw * Cfg::kAHalf;
      uint64_t* my_b_full = b_full + (RESIDENT_B ? 0 : mw * Cfg::kBHalf);
      uint64_t* my_b_empty = b_empty + (RESIDENT_B ? 0 : mw * Cfg::kBHalf);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      const int nsuper = super_end - super_begin;
      for (int si = mw; si < nsuper; si += NMMA) {
        const int acc = si % SETS;
        const uint32_t accph = (uint32_t)(si / SETS) & 1u;
        mbar_wait(&tmem_empty[acc], accph ^ 1, 13);   // this slot's epilogue set has drained the accumulators
        tc_fence_after();
        for (int sub = 0; sub < (PAR ? 1 : tps); ++sub) {
          const int ti = (si << tps_shift) + sub;
          (void)ti;
          if (lane == 0) SX_TRACE(1, ti);
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * Cfg::kSlotCols + sub * BLOCK_N);
          for (int chunk = 0; chunk < p.kchunks; ++chunk) {
            mbar_wait(&my_a_full[as], aph, 14);
            tc_fence_after();
            if (lane == 0 && chunk == 0) SX_TRACE(2, ti);
            // descriptors of tap 0; every other tap / k step is this plus a compile-time constant (fully unrolled)
            const uint64_t da0 = make_smem_desc_sbo<BLOCK_K>(a_base0 + (uint32_t)(as * Cfg::kABytes), sbo);
            const uint64_t db_res = make_smem_desc<BLOCK_K>(b_base0 + (uint32_t)(chunk * 9 * Cfg::kBBytes));
            if constexpr (PAR) {
              // column-parity form: per filter row 4 A operands (input columns 2 xc + u - 1, u = 0..3) instead of 6.
              // box 0 = odd lattice (columns x0 - 1 + 2 m), box 1 = even lattice (x0 + 2 m); weight tiles kx = 2, 1, 0.
              //   u = 3: even, m = xc + 1, W[kx=2]            -> delta = 1 half      (N = Co)
              //   u = 0: odd,  m = xc,     W[kx=0]            -> delta = 0 half      (N = Co)
              //   u = 1: even, m = xc,     W[kx=1] | W[kx=0]  -> both halves         (N = 2 Co)
              //   u = 2: odd,  m = xc + 1, W[kx=2] | W[kx=1]  -> both halves         (N = 2 Co)
              // the two N = Co instructions come first so that the very first write of each half can overwrite
              if (elect_one()) {
                if (!(dbg & 2)) {
                  const uint64_t da1 = da0 + (uint64_t)(Cfg::kBoxBytes >> 4);
                  constexpr uint64_t mshift = (uint64_t)(Cfg::kRowBytes >> 4);
                  constexpr uint64_t btile = (uint64_t)(Cfg::kBBytes >> 4);
#pragma unroll
                  for (int ky = 0; ky < 3; ++ky) {
                    const uint64_t row = (uint64_t)((ky * Cfg::kHaloW * Cfg::kRowBytes) >> 4);
                    const uint64_t dbk = db_res + (uint64_t)(ky * 3) * btile;   // tile of (ky, kx = 2)
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                      const uint32_t first = (chunk | ky | k) != 0 ? 1u : 0u;
                      umma_bf16(d_tmem + BLOCK_N, da1 + row + mshift + (uint64_t)(2 * k), dbk + (uint64_t)(2 * k), idesc, first);
                      umma_bf16(d_tmem, da0 + row + (uint64_t)(2 * k), dbk + 2 * btile + (uint64_t)(2 * k), idesc, first);
                    }
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                      umma_bf16(d_tmem, da1 + row + (uint64_t)(2 * k), dbk + btile + (uint64_t)(2 * k), idesc2, 1u);
                      umma_bf16(d_tmem, da0 + row + mshift + (uint64_t)(2 * k), dbk + (uint64_t)(2 * k), idesc2, 1u);
                    }
                  }
                }
                umma_commit(&my_a_empty[as]);
              }
              __syncwarp();
            } else if (RESIDENT_B) {
              // weights resident: nothing to wait for inside the chunk -- one election, 9 x BLOCK_K/16 back-to-back MMAs
              if (elect_one()) {
                if (!(dbg & 2)) {
#pragma unroll
                  for (int tap = 0; tap < 9; ++tap) {
                    const int ky = tap / 3, kx = tap - ky * 3;
                    const uint64_t da = da0 + (uint64_t)(((ky * HALO_W + kx) * Cfg::kRowBytes) >> 4);
                    const uint64_t db = db_res + (uint64_t)((tap * Cfg::kBBytes) >> 4);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                      umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (chunk | tap | k) != 0 ? 1u : 0u);
                  }
                }
                umma_commit(&my_a_empty[as]);
              }
              __syncwarp();
            } else {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                mbar_wait(&my_b_full[bs], bph, 15);
                tc_fence_after();
                const uint64_t db = make_smem_desc<BLOCK_K>(b_base0 + (uint32_t)(bs * Cfg::kBBytes));
                const int ky = tap / 3, kx = tap - ky * 3;
                const uint64_t da = da0 + (uint64_t)(((ky * HALO_W + kx) * Cfg::kRowBytes) >> 4);
                if (elect_one()) {
                  if (!(dbg & 2)) {
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                      umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (chunk | tap | k) != 0 ? 1u : 0u);
                  }
                  umma_commit(&my_b_empty[bs]);
                  if (tap == 8) umma_commit(&my_a_empty[as]);
                }
                __syncwarp();
                if (++bs == Cfg::kBHalf) { bs = 0; bph ^= 1; }
              }
            }
            if (++as == Cfg::kAHalf) { as = 0; aph ^= 1; }
          }
          if (lane == 0) SX_TRACE(3, ti);
        }
        if (elect_one()) umma_commit(&tmem_full[acc]);   // hand the super-tile to its epilogue set
        __syncwarp();
      }
    }
  } else if (UPS && warp_id >= 4 + 4 * SETS) {
    // ===================== fused bilinear 2x upsample: A-operand producers =====================
    // out halo pixel (hy, hx) <-> image pixel (y0-1+hy, x0-1+hx); y0, x0 are even, so halo rows (2k, 2k+1) interpolate
    // source-box rows (k, k+1) with weights (.75,.25) / (.25,.75) (align_corners=False, scale 2: src = o/2 - 0.25),
    // and likewise for columns.  Source indices are clamped to the image (torch's border rule); halo pixels outside
    // the image are the conv's zero padding.  Thread = one channel group (CH channels) x one column pair x one third of
    // the rows; horizontal pass per source row, then the vertical pass, all in (packed) fp32, one rounding to bf16.
    constexpr int CH = UPS_WARPS == 4 ? 8 : 4;      // channels per thread
    constexpr int NW = CH / 2;                      // bf16x2 words per thread
    constexpr int CGS = 64 / CH;                    // channel groups per pixel row
    const int pt = threadIdx.x - (Cfg::kFrontThreads + Cfg::kEpiThreads);
    const int cg = pt % CGS;
    // work units: (column pair j, row segment seg); 5 x 3 for an 8-pixel tile, 9 x 3 for the 16-pixel PAR super-tile.
    // A thread takes units pt / CGS, pt / CGS + units-per-pass, ...
    constexpr int kPairs = PAR ? HALO_BW + 1 : HALO_W / 2;           // halo columns 2j, 2j+1 <- source columns j, j+1
    constexpr int kUnits = 3 * kPairs;
    constexpr int kUnitsPerPass = Cfg::kUpsThreads / CGS;
    const uint32_t cbyte = (uint32_t)(cg * CH * 2);            // byte offset of the channel group inside a 128-byte pixel row
    const uint32_t c16 = cbyte >> 4, cin = cbyte & 15u;        // 16-byte chunk index (swizzled), offset inside the chunk
    const uint64_t q25 = pk2(0.25f, 0.25f), q75 = pk2(0.75f, 0.75f);
    int as2[2] = {0, 0}, ss = 0;     // A-ring position per half (tile parity), one in-order source-box ring
    uint32_t aph2[2] = {0, 0}, sph = 0;
    for (int tile = tile_begin; tile < tile_end; tile += PAR ? 2 : 1) {
      const int b = tile >> p.tpb_shift;
      const int tr = tile - (b << p.tpb_shift);
      const int ty = tr >> p.tx_shift, tx = tr - (ty << p.tx_shift);
      const int rg = NMMA == 2 ? (((tile - tile_begin) >> tps_shift) & 1) : 0;
      int as = rg ? as2[1] : as2[0];
      uint32_t aph = rg ? aph2[1] : aph2[0];
      const bool left = tx == 0, right = tx + (PAR ? 2 : 1) == p.tiles_x, top = ty == 0, bottom = ty == p.tiles_y - 1;
      const int lo_c = left ? 1 : 0, hi_c = right ? Cfg::kSrcW - 2 : Cfg::kSrcW - 1;
      const int lo_r = top ? 1 : 0, hi_r = bottom ? UPS_SRC_H - 2 : UPS_SRC_H - 1;
      for (int chunk = 0; chunk < p.kchunks; ++chunk) {
        const int st = rg * Cfg::kAHalf + as;
        mbar_wait(&s_full[ss], sph, 17);
        mbar_wait(&a_empty[st], aph ^ 1, 18);
#pragma unroll 1
        for (int unit = pt / CGS; unit < kUnits; unit += kUnitsPerPass) {
          const int j = unit % kPairs;                // halo columns 2j, 2j+1 <- source columns j, j+1
          const int seg = unit / kPairs;              // halo rows 6seg .. 6seg+5 <- source rows 3seg .. 3seg+3
          const int c0 = min(max(j, lo_c), hi_c), c1 = min(max(j + 1, lo_c), hi_c);
          const bool zero_e = left && j == 0;                 // halo column 0 is outside the image
          const bool zero_o = right && j == kPairs - 1;       // the last halo column is outside the image
          const uint8_t* src = smem_src + ss * Cfg::kSrcBytes + cbyte;
          uint8_t* dst = smem_a + st * Cfg::kABytes + cin;
          uint64_t he0[NW], ho0[NW], he1[NW], ho1[NW];
          auto ldv = [&](const uint8_t* ptr, uint32_t* w) {
            if constexpr (CH == 8) {
              const uint4 t = *reinterpret_cast<const uint4*>(ptr);
              w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
            } else {
              const uint2 t = *reinterpret_cast<const uint2*>(ptr);
              w[0] = t.x; w[1] = t.y;
            }
          };
          auto hrow = [&](int r, uint64_t* he, uint64_t* ho) {
            const int rr = min(max(r, lo_r), hi_r);
            uint32_t a[NW], c[NW];
            ldv(src + (rr * Cfg::kSrcW + c0) * Cfg::kRowBytes, a);
            ldv(src + (rr * Cfg::kSrcW + c1) * Cfg::kRowBytes, c);
#pragma unroll
            for (int k = 0; k < NW; ++k) {
              const uint64_t a2 = bf2_to_f2(a[k]), c2 = bf2_to_f2(c[k]);
              he[k] = fma2(q25, c2, mul2(q75, a2));   // odd image column 2m+1
              ho[k] = fma2(q75, c2, mul2(q25, a2));   // even image column 2m+2
            }
          };
          auto put = [&](int hy, int hx, const uint64_t* v, bool zero) {
            // PAR: halo column hx <-> image column x0 - 1 + hx lives in lattice box hx & 1 at coarse column hx >> 1
            const int pix = PAR ? hy * Cfg::kHaloW + (hx >> 1) : hy * HALO_W + hx;
            uint32_t w[NW];
#pragma unroll
            for (int k = 0; k < NW; ++k) w[k] = zero ? 0u : f2_to_bf2(v[k]);
            uint8_t* d = dst + (PAR ? (hx & 1) * Cfg::kBoxBytes : 0) + pix * Cfg::kRowBytes + ((c16 ^ (uint32_t)(pix & 7)) << 4);
            if constexpr (CH == 8) *reinterpret_cast<uint4*>(d) = make_uint4(w[0], w[1], w[2], w[3]);
            else *reinterpret_cast<uint2*>(d) = make_uint2(w[0], w[1]);
          };
          hrow(3 * seg, he0, ho0);
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int k = 3 * seg + i;
            hrow(k + 1, he1, ho1);
            const bool zt = top && k == 0;        // halo row 0 is outside the image
            const bool zb = bottom && k == 8;     // halo row 17 is outside the image
            uint64_t v[NW];
#pragma unroll
            for (int q = 0; q < NW; ++q) v[q] = fma2(q25, he1[q], mul2(q75, he0[q]));
            put(2 * k, 2 * j, v, zt || zero_e);
#pragma unroll
            for (int q = 0; q < NW; ++q) v[q] = fma2(q25, ho1[q], mul2(q75, ho0[q]));
            put(2 * k, 2 * j + 1, v, zt || zero_o);
#pragma unroll
            for (int q = 0; q < NW; ++q) v[q] = fma2(q75, he1[q], mul2(q25, he0[q]));
            put(2 * k + 1, 2 * j, v, zb || zero_e);
#pragma unroll
            for (int q = 0; q < NW; ++q) v[q] = fma2(q75, ho1[q], mul2(q25, ho0[q]));
            put(2 * k + 1, 2 * j + 1, v, zb || zero_o);
#pragma unroll
            for (int q = 0; q < NW; ++q) { he0[q] = he1[q]; ho0[q] = ho1[q]; }
          }
        }
        // generic-proxy writes -> visible to the async proxy (tcgen05.mma reads shared memory through it)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&a_full[st]);
        mbar_arrive(&s_empty[ss]);
        if (pt == 0 && chunk == 0) SX_TRACE(7, tile - tile_begin);
        if (++as == Cfg::kAHalf) { as = 0; aph ^= 1; }
        if (++ss == Cfg::kSrcStages) { ss = 0; sph ^= 1; }
      }
      if (rg) { as2[1] = as; aph2[1] = aph; } else { as2[0] = as; aph2[0] = aph; }
    }
  } else if (warp_id >= 4) {
    // ===================== epilogue: SETS sets of four warps, set s drains accumulator slot s =====================
    const ConvEpilogue& ep = p.ep;
    const int set = (warp_id - 4) >> 2;
    const int et = (int)threadIdx.x - Cfg::kFrontThreads - set * 128;   // 0..127 inside the set
    const int q = warp_id & 3;          // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;        // tile row = TMEM lane
    const int yy = r >> 3, xx = r & 7;
    // pixel j of this thread: x + JSTEP * j.  Plain: the same lane of TPS tiles 8 pixels apart; PAR: lane (yy, xc) owns the
    // two adjacent pixels x0 + 2 xc + {0, 1} of its 16-pixel super-tile
    constexpr int JSTEP = PAR ? 1 : HALO_BW;
    constexpr int XSCALE = PAR ? 2 : 1;
    float* tab = s_tab + set * (10 * BLOCK_N);
    const uint32_t bar_id = 1u + (uint32_t)set;
    const uint32_t HW = (uint32_t)(p.H * p.W);
    const uint32_t S = (uint32_t)ep.noise_size;
    constexpr bool fuse_rgb = FUSE_RGB;   // compile-time: the plain instantiation carries none of the ToRGB registers
    uint32_t accph = 0;
    auto tile_coords = [&](int tile, int& b, int& x, int& y) {
      b = tile >> p.tpb_shift;
      const int tr = tile - (b << p.tpb_shift);
      const int ty = tr >> p.tx_shift, tx = tr - (ty << p.tx_shift);
      x = tx * HALO_BW + XSCALE * xx;
      y = ty * HALO_BH + yy;
    };
    // 32-bit element offsets (the launcher checks every tensor is below 2^31 elements)
    auto load_noise = [&](int b, int x, int y) -> float {
      if (!ep.noise) return 0.f;
      return __ldg(ep.noise + ((ep.noise_batch == 1 ? 0u : (uint32_t)b * S * S) + (uint32_t)x * S + (uint32_t)y));
    };
    auto rgb_off = [&](int b, int x, int y) -> uint32_t { return ((uint32_t)b * 3u * (uint32_t)p.H + (uint32_t)y) * (uint32_t)p.W + (uint32_t)x; };
    auto load_rgb_prev = [&](int b, int x, int y, float* v) {
      v[0] = v[1] = v[2] = 0.f;
      if (fuse_rgb && ep.rgb_accumulate) {
        const float* src = ep.rgb_out + rgb_off(b, x, y);
        v[0] = __ldg(src);
        v[1] = __ldg(src + HW);
        v[2] = __ldg(src + 2u * HW);
      }
    };
    // per-sample tables (demod coefficients, next-layer style + 1, fused-ToRGB weights) live in two slots per set;
    // a sample change (rare: tiles are walked sample-major) fills the other slot and synchronises the set's four warps once
    auto write_tables = [&](int slot, int b) {
      float* t = tab + slot * (5 * BLOCK_N);
      for (int i = et; i < BLOCK_N; i += 128) {
        t[i] = ep.dcoef ? __ldg(ep.dcoef + (long long)b * ep.dcoef_stride + n0 + i) : 1.f;
        t[BLOCK_N + i] = ep.next_style ? __ldg(ep.next_style + (long long)b * ep.next_style_stride + n0 + i) + 1.f : 1.f;
      }
      if (fuse_rgb) {
        for (int i = et; i < 3 * BLOCK_N; i += 128) {
          const int o = i % BLOCK_N;
          t[2 * BLOCK_N + i] = (__ldg(ep.rgb_style + (long long)b * ep.rgb_style_stride + o) + 1.f) * __ldg(ep.rgb_w + (i / BLOCK_N) * p.Co + o);
        }
      }
    };
    // the generator's configuration (activation on, bf16 NHWC or no feature-map output) takes the packed fast path
    const bool fast = ep.act != 0 && !ep.out_nchw_f32;
    const bool has_out = ep.out != nullptr, has_raw = ep.out_raw != nullptr;
    // pixel j of this thread in the super-tile: same sample, same row, x advancing by 8 per tile
    int sup = super_begin + set;
    int b = 0, x = 0, y = 0, cur_b = -1, slot = 1;
    float nz[TPS];
    auto load_noises = [&](int b_, int x_, int y_, float* nzv) {
#pragma unroll
      for (int j = 0; j < TPS; ++j) nzv[j] = j < tps ? load_noise(b_, x_ + JSTEP * j, y_) : 0.f;
    };
    if (sup < super_end) {
      tile_coords(sup * tps, b, x, y);
      load_noises(b, x, y, nz);
    }
    for (; sup < super_end; sup += SETS) {
      if (b != cur_b) {
        slot ^= 1;
        write_tables(slot, b);
        cur_b = b;
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // tables visible to the set's four warps
      }
      const float* dd = tab + slot * (5 * BLOCK_N);
      const float* mm = dd + BLOCK_N;
      const float* rw = dd + 2 * BLOCK_N;
      const uint32_t pix = ((uint32_t)b * (uint32_t)p.H + (uint32_t)y) * (uint32_t)p.W + (uint32_t)x;
      const uint32_t roff = rgb_off(b, x, y);
      // prefetch the next super-tile's noise values (consumed after this one's TMEM drain); the previous rgb of THIS
      // super-tile is only needed at the very end (rgb = prev + sum), so its loads fly during the math
      int b2 = b, x2 = 0, y2 = 0;
      float nzn[TPS];
      if (sup + SETS < super_end) {
        tile_coords((sup + SETS) * tps, b2, x2, y2);
        load_noises(b2, x2, y2, nzn);
      } else {
#pragma unroll
        for (int j = 0; j < TPS; ++j) nzn[j] = 0.f;
      }
      float rgb_acc[TPS][3];   // starts as the previous rgb, ends as the result
#pragma unroll
      for (int j = 0; j < TPS; ++j) {
        rgb_acc[j][0] = rgb_acc[j][1] = rgb_acc[j][2] = 0.f;
        if (j < tps) load_rgb_prev(b, x + JSTEP * j, y, rgb_acc[j]);
      }
      if (et == 0) SX_TRACE(6, sup - super_begin);
      mbar_wait(&tmem_full[set], accph, 16);
      tc_fence_after();
      if (et == 0) SX_TRACE(4, sup - super_begin);
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * Cfg::kSlotCols);
      if (dbg & 1) {
      } else if (fast) {
        uint64_t racc[TPS][3];
#pragma unroll
        for (int j = 0; j < TPS; ++j) racc[j][0] = racc[j][1] = racc[j][2] = 0ull;
        __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(ep.out) + ((size_t)pix * (uint32_t)p.Co + (uint32_t)n0);
        __nv_bfloat16* rawp = reinterpret_cast<__nv_bfloat16*>(ep.out_raw) + ((size_t)pix * (uint32_t)p.Co + (uint32_t)n0);
        const uint32_t jstride = (uint32_t)(JSTEP * p.Co);   // elements between this thread's pixels j and j + 1
        // CW columns per pass, double-buffered: the TMEM loads of pass k + 1 are in flight during the math of pass k.
        // The register budget (threads per CTA) picks the width.
        constexpr int CW = Cfg::kThreads <= 384 ? 16 : 8;
        constexpr int NPASS = BLOCK_N / CW;
        uint32_t v[2][TPS][CW];
        auto issue_loads = [&](int pass, uint32_t (*dst)[CW]) {
#pragma unroll
          for (int j = 0; j < TPS; ++j) {
            if constexpr (CW == 16) tmem_ld16(tbase + (uint32_t)(j * BLOCK_N + pass * CW), dst[j]);
            else tmem_ld8(tbase + (uint32_t)(j * BLOCK_N + pass * CW), dst[j]);
          }
        };
        issue_loads(0, v[0]);
        // fully unrolled: the buffer index must be a compile-time constant (a run-time index, or passing the buffers
        // through a helper, sends v[] to local memory -- measured: 1.6x slower)
#pragma unroll
        for (int pass = 0; pass < NPASS; ++pass) {
          const int c0 = pass * CW;
          tmem_ld_wait();
          if (pass + 1 < NPASS) issue_loads(pass + 1, v[(pass + 1) & 1]);
          uint32_t (*vv)[CW] = v[pass & 1];
          uint32_t om[TPS][CW / 2], orw[TPS][CW / 2];
#pragma unroll
          for (int g = 0; g < CW / 4; ++g) {
            const int c = c0 + 4 * g;
            EpiTab4 t;
            t.d = *reinterpret_cast<const float4*>(dd + c);
            t.w = *reinterpret_cast<const float4*>(s_nw + c);
            t.b = *reinterpret_cast<const float4*>(s_nb + c);
            t.m = make_float4(1.f, 1.f, 1.f, 1.f);
            if (has_out) t.m = *reinterpret_cast<const float4*>(mm + c);
            t.r0 = t.r1 = t.r2 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (fuse_rgb) {
              t.r0 = *reinterpret_cast<const float4*>(rw + c);
              t.r1 = *reinterpret_cast<const float4*>(rw + BLOCK_N + c);
              t.r2 = *reinterpret_cast<const float4*>(rw + 2 * BLOCK_N + c);
            }
#pragma unroll
            for (int j = 0; j < TPS; ++j) {
              if (j < tps) {
                if (has_out) {
                  if (has_raw) epi_group4<fuse_rgb, true, true>(vv[j] + 4 * g, t, pk2(nz[j], nz[j]), om[j] + 2 * g, orw[j] + 2 * g, racc[j]);
                  else epi_group4<fuse_rgb, true, false>(vv[j] + 4 * g, t, pk2(nz[j], nz[j]), om[j] + 2 * g, orw[j] + 2 * g, racc[j]);
                } else {
                  epi_group4<fuse_rgb, false, false>(vv[j] + 4 * g, t, pk2(nz[j], nz[j]), om[j] + 2 * g, orw[j] + 2 * g, racc[j]);
                }
              }
            }
          }
          if (has_out) {
#pragma unroll
            for (int j = 0; j < TPS; ++j) {
              if (j < tps) {
                uint4* o4 = reinterpret_cast<uint4*>(outp + (size_t)j * jstride + c0);
                uint4* r4 = reinterpret_cast<uint4*>(rawp + (size_t)j * jstride + c0);
#pragma unroll
                for (int k = 0; k < CW / 8; ++k) {
                  o4[k] = make_uint4(om[j][4 * k], om[j][4 * k + 1], om[j][4 * k + 2], om[j][4 * k + 3]);
                  if (has_raw) r4[k] = make_uint4(orw[j][4 * k], orw[j][4 * k + 1], orw[j][4 * k + 2], orw[j][4 * k + 3]);
                }
              }
            }
          }
        }
        if (fuse_rgb) {
#pragma unroll
          for (int j = 0; j < TPS; ++j) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              float lo, hi;
              upk2(racc[j][c], lo, hi);
              rgb_acc[j][c] += lo + hi;
            }
          }
        }
      } else {
        // generic configuration (module-level Conv2DMod in bf16: no activation and / or NCHW fp32 output), 8 columns a time
#pragma unroll
        for (int j = 0; j < TPS; ++j) {
          if (j >= tps) break;
          const int xj = x + JSTEP * j;
          const size_t pixj = (size_t)pix + (size_t)(JSTEP * j);
#pragma unroll 1
          for (int c0 = 0; c0 < BLOCK_N; c0 += 8) {
            uint32_t v[8];
            tmem_ld8(tbase + (uint32_t)(j * BLOCK_N + c0), v);
            tmem_ld_wait();
            float f[8], fr[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float t = __uint_as_float(v[k]) * dd[c0 + k];
              t += nz[j] * s_nw[c0 + k] + s_nb[c0 + k];
              if (ep.act) t = lrelu02(t);
              if (!ep.out_nchw_f32) t = __bfloat162float(__float2bfloat16_rn(t));
              fr[k] = t;
              f[k] = t * mm[c0 + k];
              if (fuse_rgb) {
#pragma unroll
                for (int c = 0; c < 3; ++c) rgb_acc[j][c] = fmaf(t, rw[c * BLOCK_N + c0 + k], rgb_acc[j][c]);
              }
            }
            if (!ep.out) {
            } else if (ep.out_nchw_f32) {
              float* out = reinterpret_cast<float*>(ep.out);
#pragma unroll
              for (int k = 0; k < 8; ++k) out[(((size_t)b * p.Co + n0 + c0 + k) * p.H + y) * p.W + xj] = f[k];
            } else {
              uint4 pk;
              pack(f, pk);
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + (pixj * p.Co + n0 + c0)) = pk;
            }
            if (ep.out_raw) {
              uint4 pk;
              pack(fr, pk);
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out_raw) + (pixj * p.Co + n0 + c0)) = pk;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[set]);   // 128 arrivals release the accumulator slot to the MMA warp
      if (et == 0) SX_TRACE(5, sup - super_begin);
      if (fuse_rgb) {
#pragma unroll
        for (int j = 0; j < TPS; ++j) {
          if (j < tps) {
            float* dst = ep.rgb_out + roff + JSTEP * j;
            dst[0] = rgb_acc[j][0];
            dst[HW] = rgb_acc[j][1];
            dst[2u * HW] = rgb_acc[j][2];
          }
        }
      }
      accph ^= 1;
      b = b2; x = x2; y = y2;
#pragma unroll
      for (int j = 0; j < TPS; ++j) nz[j] = nzn[j];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_id == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

inline bool halo_shape_supported(int Ci, int Co, int H, int W) {
  auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  if (!(Co == 32 || Co == 64 || Co == 128)) return false;
  if (Ci % 32 != 0) return false;
  if (H != W || !pow2(W) || H < HALO_BH) return false;
  return true;
}

// x: the conv input [B,H,W,Ci] -- or, for UPS, the low-resolution tensor [B,H/2,W/2,Ci] the kernel upsamples itself
template <int BLOCK_N, int BLOCK_K, int A_STAGES, int B_STAGES, bool RESIDENT_B, bool FUSE_RGB, bool UPS, int SETS, int UPS_WARPS, int TPS,
          int NMMA, bool PAR = false>
int launch_conv_halo_cfg2(const __nv_bfloat16* x, const __nv_bfloat16* wk, ConvHaloParams p, cudaStream_t stream) {
  using Cfg = HaloCfg<BLOCK_N, BLOCK_K, A_STAGES, B_STAGES, RESIDENT_B, UPS, SETS, UPS_WARPS, TPS, NMMA, PAR>;
  // a super-tile never leaves its sample or its tile row: TPS divides tiles_x (a power of two >= 2, since W >= 16)
  if (p.tiles_x % TPS != 0) return fail(SX_EUNSUPPORTED, "conv_tc_halo: %d tiles per row is not a multiple of %d", p.tiles_x, TPS);
  p.tps = TPS;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(SX_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  // the epilogue / producers index with 32-bit element offsets
  const long long big = (long long)p.B * p.H * p.W * (p.Ci > p.Co ? p.Ci : p.Co);
  if (big >= (1ll << 31)) return fail(SX_EUNSUPPORTED, "conv_tc_halo: tensor of %lld elements exceeds 32-bit indexing (split the batch)", big);
  const CUtensorMapSwizzle swz = BLOCK_K == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMap ta, tb;
  {
    const int IH = UPS ? p.H / 2 : p.H, IW = UPS ? p.W / 2 : p.W;
    cuuint64_t gdim[4] = {(cuuint64_t)p.Ci, (cuuint64_t)IW, (cuuint64_t)IH, (cuuint64_t)p.B};
    cuuint64_t gstr[3] = {(cuuint64_t)p.Ci * 2, (cuuint64_t)IW * p.Ci * 2, (cuuint64_t)IH * IW * p.Ci * 2};
    // PAR without the fused upsample: every other column (traversal stride 2) over a span of 17 = 9 columns per box
    cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)(UPS ? Cfg::kSrcW : (PAR ? 2 * Cfg::kHaloW - 1 : HALO_W)),
                         (cuuint32_t)(UPS ? UPS_SRC_H : HALO_H), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)((PAR && !UPS) ? 2 : 1), 1, 1};
    CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(x), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, UPS ? CU_TENSOR_MAP_SWIZZLE_NONE : swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SX_ECUDA, "cuTensorMapEncodeTiled(A halo) failed: %d", (int)r);
  }
  {
    const cuuint64_t ktot = (cuuint64_t)9 * p.Ci;
    cuuint64_t gdim[2] = {ktot, (cuuint64_t)p.Co};
    cuuint64_t gstr[1] = {ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)BLOCK_N};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(wk), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SX_ECUDA, "cuTensorMapEncodeTiled(B) failed: %d", (int)r);
  }
  auto kern = conv_tc_halo_kernel<BLOCK_N, BLOCK_K, A_STAGES, B_STAGES, RESIDENT_B, FUSE_RGB, UPS, SETS, UPS_WARPS, TPS, NMMA, PAR>;
  const size_t smem = Cfg::smem_bytes(p.num_b_tiles);
  if (smem > 227 * 1024) return fail(SX_EUNSUPPORTED, "conv_tc_halo: %zu bytes of shared memory needed", smem);
  static size_t configured = 0;
  static int occ_cached = 0;
  if (smem > configured) {
    SX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
    occ_cached = 0;
  }
  if (occ_cached == 0) {
    // resident CTAs per SM, computed by hand (the occupancy API answered 1 for a kernel ncu showed could host 2:
    // it assumes the default shared-memory carve-out): 227 KB smem incl. 1 KB/CTA driver reserve, 64K registers,
    // 2048 threads, 512 TMEM columns.
    SX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    cudaFuncAttributes fa;
    SX_CUDA(cudaFuncGetAttributes(&fa, kern));
    const int regs = (fa.numRegs + 7) / 8 * 8;
    int occ = (int)((227 * 1024) / (smem + 1024));
    occ = std::min(occ, 65536 / (regs * Cfg::kThreads));
    occ = std::min(occ, 2048 / Cfg::kThreads);
    occ = std::min(occ, 512 / Cfg::kTmemCols);
    occ_cached = 1;   // every configuration is sized for one CTA per SM (__launch_bounds__(threads, 1))
    (void)occ;
  }
  int grid_x = occ_cached * num_sms();
  if (grid_x > p.num_tiles / p.tps) grid_x = p.num_tiles / p.tps;
  dim3 grid((unsigned)grid_x, (unsigned)(p.Co / BLOCK_N));
#ifdef SX_HALO_DEBUG_KNOBS
  // SX_HALO_TRACE=Ci_Co: clock64 stamps of the pipeline roles of CTA (0,0) for the first 64 tiles, printed for the 3rd
  // matching launch (after warm-up)
  static int trace_hits = 0;
  static unsigned long long* trace_dev = nullptr;
  bool tracing = false;
  if (const char* tr = getenv("SX_HALO_TRACE")) {
    char want[64];
    snprintf(want, sizeof(want), "%d_%d_%d", p.Ci, p.Co, (int)FUSE_RGB);
    if (!strcmp(tr, want) && ++trace_hits == 3) {
      if (!trace_dev) cudaMalloc(&trace_dev, 8 * 64 * sizeof(unsigned long long));
      cudaMemsetAsync(trace_dev, 0, 8 * 64 * sizeof(unsigned long long), stream);
      p.trace = trace_dev;
      tracing = true;
    }
  }
#endif
  kern<<<grid, Cfg::kThreads, smem, stream>>>(ta, tb, p);
  SX_CHECK_LAUNCH();
#ifdef SX_HALO_DEBUG_KNOBS
  if (tracing) {
    unsigned long long h[8 * 64];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull;
    for (int i = 0; i < 8 * 64; ++i) if (h[i] && h[i] < t0) t0 = h[i];
    fprintf(stderr, "HALOTRACE Ci=%d Co=%d rgb=%d tps=%d grid=%d threads=%d  (clk relative to first stamp; roles: 0 tma-issue 1 mma-slot-free 2 mma-a-ready 3 mma-issued 6 epi-at-wait 4 epi-tmem-ready 5 epi-done 7 ups-produced)\n",
            p.Ci, p.Co, (int)FUSE_RGB, p.tps, grid_x, Cfg::kThreads);
    for (int i = 0; i < 48; ++i) {
      fprintf(stderr, "HALOTRACE t%02d", i);
      for (int r = 0; r < 8; ++r) fprintf(stderr, " r%d=%lld", r, h[r * 64 + i] ? (long long)(h[r * 64 + i] - t0) : -1ll);
      fprintf(stderr, "\n");
    }
  }
#endif
  return SX_OK;
}

// plain (no fused upsample) configurations: ToRGB fusion is a run-time property of the epilogue descriptor
template <int BLOCK_N, int BLOCK_K, int A_STAGES, int B_STAGES, bool RESIDENT_B, int SETS, int TPS, int NMMA, bool PAR = false>
int launch_conv_halo_cfg(const __nv_bfloat16* x, const __nv_bfloat16* wk, ConvHaloParams p, cudaStream_t stream) {
  if (p.ep.rgb_style) return launch_conv_halo_cfg2<BLOCK_N, BLOCK_K, A_STAGES, B_STAGES, RESIDENT_B, true, false, SETS, 4, TPS, NMMA, PAR>(x, wk, p, stream);
  return launch_conv_halo_cfg2<BLOCK_N, BLOCK_K, A_STAGES, B_STAGES, RESIDENT_B, false, false, SETS, 4, TPS, NMMA, PAR>(x, wk, p, stream);
}

inline ConvHaloParams make_halo_params(int B, int Ci, int Co, int H, int W, int bk, const ConvEpilogue& ep) {
  ConvHaloParams p;
  p.B = B; p.H = H; p.W = W; p.Ci = Ci; p.Co = Co;
  p.tiles_x = W / HALO_BW; p.tiles_y = H / HALO_BH;
  p.num_tiles = B * p.tiles_x * p.tiles_y;
  p.tx_shift = 0;
  while ((1 << p.tx_shift) < p.tiles_x) ++p.tx_shift;
  p.tpb_shift = 0;
  while ((1 << p.tpb_shift) < p.tiles_x * p.tiles_y) ++p.tpb_shift;
  static const int dbg = getenv("SX_HALO_DEBUG") ? atoi(getenv("SX_HALO_DEBUG")) : 0;
  p.debug = dbg;
  p.kchunks = Ci / bk;
  p.num_b_tiles = 9 * p.kchunks;
  p.tps = 1;
  p.trace = nullptr;
  p.ep = ep;
  return p;
}

// SX_HALO_VARIANT (bitmask, tuning experiments): 1 = the 32 -> 32 layers run 2 epilogue sets (wide passes) instead of 4;
// 2 = the 64 -> 64 layers run 2 sets x 2 tiles with two MMA warps instead of 4 sets x 1 tile with one;
// 4 = the fused-upsample 128 -> 64 layer runs 2 tiles per slot with two MMA warps;
// 8 = the weight-streaming Co = 128 layers run the old 4-deep weight ring instead of 8.
// Ring depths tried and rejected in round 2 (256 px, batch 256): plain 64 -> 64 with 6 activation stages 0.372 -> 0.390 ms,
// plain 128 -> 128 with 4 activation stages + a 6-deep weight ring 0.282 -> 0.288 ms.
// SX_HALO_MAX_CO: plain (non-upsample) layers wider than this go to conv_tc_kernel instead (A/B of the two kernels)
// SX_HALO_PAR (bitmask, default 1): layers that run the column-parity form.  Measured at 256 px, batch 256 (profiles/README.md
// r02d/e): 1 = 32 -> 32 (+ToRGB): 0.615 -> 0.551 ms (default ON); 2 = 64 -> 64 (+ToRGB) with 32-channel chunks x 4 stages:
// 0.371 -> 0.492 ms (two pixels per lane at N = 64 spill in the ToRGB epilogue at the 96-register budget of 640 threads;
// 3 stages of 64 channels with one MMA warp: 0.505); 4 = fused-upsample 64 -> 32 with 3 stages / one MMA warp: 0.853 ->
// 0.835 ms (2 stages / two MMA warps: 1.04) -- that layer is issue-bound on its upsample producers + epilogue (ncu: 65 %
// issue-active), not on the operand reads.  2 and 4 stay selectable for A/B runs; the GPU tests cover them (SX_HALO_PAR=7).
inline int halo_par() {
  static const int v = getenv("SX_HALO_PAR") ? atoi(getenv("SX_HALO_PAR")) : 1;
  return v;
}
inline int halo_variant() {
  static const int v = getenv("SX_HALO_VARIANT") ? atoi(getenv("SX_HALO_VARIANT")) : 0;
  return v;
}

// *handled = false (and SX_OK) when the halo kernel does not cover the shape
inline int launch_conv_halo(const __nv_bfloat16* x, const __nv_bfloat16* wk, int B, int Ci, int Co, int H, int W,
                            const ConvEpilogue& ep, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (!halo_shape_supported(Ci, Co, H, W) || B == 0) return SX_OK;
  static const int max_co = getenv("SX_HALO_MAX_CO") ? atoi(getenv("SX_HALO_MAX_CO")) : 128;
  if (Co > max_co) return SX_OK;
  const int bk = Ci % 64 == 0 ? 64 : 32;
  const ConvHaloParams p = make_halo_params(B, Ci, Co, H, W, bk, ep);
  const size_t weight_bytes = (size_t)9 * Ci * Co * 2;
  const bool resident = weight_bytes <= 80 * 1024;
  *handled = true;
  //                                                                    N   K  A  B  resident SETS TPS MMA-warps
  if (Co == 32 && bk == 64 && resident) return launch_conv_halo_cfg<32, 64, 4, 2, true, 4, 2, 2>(x, wk, p, stream);
  if (Co == 32 && bk == 32 && resident) {
    // column-parity form (SX_HALO_PAR bit 1)
    // 8 stages of two 11 KB lattice boxes (4 stages: 0.523 ms at 256 px / batch 256; 8: 0.445 -- the layer is load-latency bound)
    if ((halo_par() & 8) && W >= 16) return launch_conv_halo_cfg<32, 32, 4, 2, true, 4, 2, 2, true>(x, wk, p, stream);   // A/B: 4 stages
    if ((halo_par() & 1) && W >= 16) return launch_conv_halo_cfg<32, 32, 8, 2, true, 4, 2, 2, true>(x, wk, p, stream);
    if (halo_variant() & 1) return launch_conv_halo_cfg<32, 32, 8, 2, true, 2, 2, 2>(x, wk, p, stream);
    return launch_conv_halo_cfg<32, 32, 8, 2, true, 4, 2, 2>(x, wk, p, stream);
  }
  if (Co == 64 && bk == 64 && resident) {
    // column-parity form (SX_HALO_PAR bit 2): 2 accumulator slots of 2 x 64 columns per set ... 4 sets x 128 = 512 columns
    if ((halo_par() & 2) && W >= 16)    // 32-channel chunks: four 22 KB stages instead of two 43 KB ones
      return launch_conv_halo_cfg<64, 32, 4, 2, true, 4, 2, 2, true>(x, wk, make_halo_params(B, Ci, Co, H, W, 32, ep), stream);
    if (halo_variant() & 2) return launch_conv_halo_cfg<64, 64, 4, 2, true, 2, 2, 2>(x, wk, p, stream);
    return launch_conv_halo_cfg<64, 64, 4, 2, true, 4, 1, 1>(x, wk, p, stream);
  }
  // 128 -> 64 channels (147 KB of weights): still resident, with a 2-deep activation ring
  if (Co == 64 && bk == 64 && weight_bytes <= 150 * 1024) return launch_conv_halo_cfg<64, 64, 2, 2, true, 4, 1, 1>(x, wk, p, stream);
  if (Co == 64 && bk == 64) return launch_conv_halo_cfg<64, 64, 3, 6, false, 4, 1, 1>(x, wk, p, stream);
  if (Co == 128 && bk == 64) {
    // 8-deep weight ring: with 4 stages (4 x 272 cycles of MMA = 0.57 us of look-ahead, about one loaded L2 round trip) the
    // streamed weights arrived late: 256->128@64 1017 -> 1136 TFLOP/s, 128->128@64 1041 -> 1095 (profiles/README.md r02a)
    if (halo_variant() & 8) return launch_conv_halo_cfg<128, 64, 3, 4, false, 2, 2, 1>(x, wk, p, stream);
    return launch_conv_halo_cfg<128, 64, 3, 8, false, 2, 2, 1>(x, wk, p, stream);
  }
  *handled = false;
  return SX_OK;
}

// Fused-upsample variant: xlow is the LOW-RESOLUTION input [B,H/2,W/2,Ci] (already modulated); H, W are the conv's
// (output) sizes.  The upsampled tensor is never written to memory: 4x less input traffic and no upsample kernel.
inline bool halo_ups_supported(int Ci, int Co, int H, int W) {
  if (!halo_shape_supported(Ci, Co, H, W) || Ci % 64 != 0 || H < 32) return false;
  const size_t weight_bytes = (size_t)9 * Ci * Co * 2;
  if (Co == 32) return weight_bytes <= 80 * 1024;
  return Co == 64 || Co == 128;
}
inline int launch_conv_halo_ups(const __nv_bfloat16* xlow, const __nv_bfloat16* wk, int B, int Ci, int Co, int H, int W,
                                const ConvEpilogue& ep, cudaStream_t stream) {
  if (!halo_ups_supported(Ci, Co, H, W)) return fail(SX_EUNSUPPORTED, "conv_tc_halo(ups): unsupported shape Ci=%d Co=%d H=%d", Ci, Co, H);
  if (ep.rgb_style) return fail(SX_EINVAL, "conv_tc_halo(ups): fused ToRGB is a conv2 feature");
  if (B == 0) return SX_OK;
  const ConvHaloParams p = make_halo_params(B, Ci, Co, H, W, 64, ep);
  const size_t weight_bytes = (size_t)9 * Ci * Co * 2;
  //                                              N   K  A  B  resident rgb   ups  SETS producer-warps TPS MMA-warps
  if (Co == 32 && (halo_par() & 4)) return launch_conv_halo_cfg2<32, 64, 3, 2, true, false, true, 4, 8, 2, 1, true>(xlow, wk, p, stream);
  if (Co == 32) return launch_conv_halo_cfg2<32, 64, 4, 2, true, false, true, 4, 8, 2, 2>(xlow, wk, p, stream);
  if (Co == 64 && weight_bytes <= 150 * 1024) {
    if (halo_variant() & 4) return launch_conv_halo_cfg2<64, 64, 2, 2, true, false, true, 4, 8, 2, 2>(xlow, wk, p, stream);
    return launch_conv_halo_cfg2<64, 64, 2, 2, true, false, true, 4, 8, 1, 1>(xlow, wk, p, stream);
  }
  if (Co == 64) return launch_conv_halo_cfg2<64, 64, 2, 6, false, false, true, 4, 8, 1, 1>(xlow, wk, p, stream);
  if (halo_variant() & 8) return launch_conv_halo_cfg2<128, 64, 3, 4, false, false, true, 2, 8, 2, 1>(xlow, wk, p, stream);
  return launch_conv_halo_cfg2<128, 64, 3, 8, false, false, true, 2, 8, 2, 1>(xlow, wk, p, stream);
}

// bf16 Conv2DMod dispatch: the halo-reusing persistent kernel where it applies, the per-tap kernel otherwise.
// SX_DISABLE_HALO=1 forces the per-tap kernel (A/B measurements).
inline int launch_conv_bf16(const __nv_bfloat16* x, const __nv_bfloat16* wk, int B, int Ci, int Co, int H, int W, int KS,
                            const ConvEpilogue& ep, cudaStream_t stream) {
  static const bool halo_off = getenv("SX_DISABLE_HALO") != nullptr;
  if (!halo_off && KS == 3) {
    bool handled = false;
    SX_TRY(launch_conv_halo(x, wk, B, Ci, Co, H, W, ep, stream, &handled));
    if (handled) return SX_OK;
  }
  return launch_conv_tc(x, wk, B, Ci, Co, H, W, KS, ep, stream);
}

}  // namespace tc
}  // namespace sx
