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

#include "conv_tc.cuh"

namespace sx {
namespace tc {

constexpr int HALO_BW = 8, HALO_BH = 16, HALO_W = HALO_BW + 2, HALO_H = HALO_BH + 2, HALO_ROWS = HALO_W * HALO_H;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// like make_smem_desc but with an explicit stride between 8-row groups
template <int BLOCK_K>
__device__ __forceinline__ uint64_t make_smem_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  constexpr uint64_t layout = BLOCK_K * 2 == 128 ? 2 : 4;  // SWIZZLE_128B = 2, SWIZZLE_64B = 4
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (layout << 61);
}

// fused-upsample variant (UPS): the conv input is the LOW-RESOLUTION tensor; a 10 x 6 pixel source box per channel chunk
// is TMA-loaded (un-swizzled) into its own ring and four producer warps write the bilinearly upsampled 18 x 10 halo box
// into the swizzled A stage (see the producer branch of the kernel).
constexpr int UPS_SRC_W = HALO_W / 2 + 1, UPS_SRC_H = HALO_H / 2 + 1;   // 6 x 10 source pixels
constexpr int UPS_S_STAGES = 2;
constexpr int UPS_THREADS = 128;

template <int BLOCK_N, int BLOCK_K, int A_STAGES, int B_STAGES, bool RESIDENT_B, bool UPS = false>
struct HaloCfg {
  static constexpr int kRowBytes = BLOCK_K * 2;
  static constexpr int kThreads = NUM_THREADS + (UPS ? UPS_THREADS : 0);
  static constexpr int kSrcBytes = UPS_SRC_W * UPS_SRC_H * kRowBytes;   // one source box (7680 B at BLOCK_K = 64)
  static constexpr int kSrcRegion = UPS ? UPS_S_STAGES * kSrcBytes : 0;
  static constexpr int kATx = HALO_ROWS * kRowBytes;              // bytes one halo box delivers
  static constexpr int kABytes = (kATx + 1023) / 1024 * 1024;     // stage stride (keeps every stage 1024 B aligned)
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;           // weights of one (channel chunk, tap)
  static constexpr int kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  static_assert((kTmemCols & (kTmemCols - 1)) == 0 && kTmemCols <= 512, "TMEM columns must be a power of two <= 512");
  __host__ __device__ static size_t b_region(int num_b_tiles) { return (size_t)(RESIDENT_B ? num_b_tiles : B_STAGES) * kBBytes; }
  static size_t smem_bytes(int num_b_tiles) {
    return 1024 + (size_t)A_STAGES * kABytes + b_region(num_b_tiles) + 256 + (size_t)(2 + 4 + 6) * BLOCK_N * sizeof(float) + kSrcRegion;
  }
};

struct ConvHaloParams {
  int B, H, W, Ci, Co;
  int tiles_x, tiles_y, num_tiles;   // W/8, H/16, B*tiles_y*tiles_x
  int tx_shift, tpb_shift;           // log2(tiles_x), log2(tiles_x * tiles_y): H, W are powers of two
  int debug;                         // SX_HALO_DEBUG bitmask (bottleneck experiments; results are garbage when set):
                                     //   1 skip epilogue math/stores, 2 skip MMA issue, 4 skip activation TMA loads
  int kchunks, num_b_tiles;          // Ci/BLOCK_K, 9*kchunks
  ConvEpilogue ep;
};

// resident CTAs per SM the register budget is compiled for: the narrow layers' per-tile chain (TMEM wait -> ld ->
// math -> barrier) is latency-bound, so more independent CTAs per SM is what hides it
template <int BLOCK_N, int BLOCK_K>
constexpr int halo_min_ctas() { return BLOCK_N == 32 ? 2 : 1; }   // 2 CTAs/SM where smem allows it; 3 (96 regs, spills) measured slower

template <int BLOCK_N, int BLOCK_K, int A_STAGES, int B_STAGES, bool RESIDENT_B, bool FUSE_RGB, bool UPS>
__global__ void __launch_bounds__(NUM_THREADS + (UPS ? UPS_THREADS : 0), halo_min_ctas<BLOCK_N, BLOCK_K>())
conv_tc_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const ConvHaloParams p) {
  using Cfg = HaloCfg<BLOCK_N, BLOCK_K, A_STAGES, B_STAGES, RESIDENT_B, UPS>;
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
  uint64_t* tmem_full = b_empty + B_STAGES;       // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2]
  uint64_t* s_full = tmem_empty + 2;              // [UPS_S_STAGES]  (UPS only) source box landed
  uint64_t* s_empty = s_full + UPS_S_STAGES;      // [UPS_S_STAGES]  (UPS only) producers are done with the source box
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(s_empty + UPS_S_STAGES);
  static_assert((2 * A_STAGES + 2 * B_STAGES + 4 + 2 * UPS_S_STAGES) * 8 + 8 <= 256, "barrier block overflow");
  float* s_nw = reinterpret_cast<float*>(after + 256);
  float* s_nb = s_nw + BLOCK_N;
  float* s_d = s_nb + BLOCK_N;            // [2][BLOCK_N] demod coefficients of the tile's sample (per accumulator slot)
  float* s_m = s_d + 2 * BLOCK_N;         // [2][BLOCK_N] next-layer (style+1)
  float* s_rgbw = s_m + 2 * BLOCK_N;      // [2][3][BLOCK_N] fused-ToRGB weights of the tile's sample
  uint8_t* smem_src = reinterpret_cast<uint8_t*>(s_rgbw + 6 * BLOCK_N);   // (UPS only) [UPS_S_STAGES][10][6][BLOCK_K] bf16

  const int warp_id = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BLOCK_N;
  // Each CTA walks a CONTIGUOUS range of tiles (sample-major, then rows, then columns): the tile's sample -- and with it
  // the per-sample epilogue tables -- changes once per several hundred tiles instead of every other tile, and
  // consecutive tiles share halo columns in L2.
  const int tile_begin = (int)((long long)blockIdx.x * p.num_tiles / gridDim.x);
  const int tile_end = (int)((long long)(blockIdx.x + 1) * p.num_tiles / gridDim.x);
#ifdef SX_HALO_DEBUG_KNOBS   // bottleneck experiments only (profiles/README.md); never in the shipped library
  const int dbg = p.debug;
#else
  constexpr int dbg = 0;
#endif

  if (warp_id == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < A_STAGES; ++s) { mbar_init(&a_full[s], UPS ? UPS_THREADS : 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < B_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 128); }
    if (UPS) for (int s = 0; s < UPS_S_STAGES; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], UPS_THREADS); }
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
            const int chunk = i / 9, tap = i - chunk * 9;
            tma_load_2d(smem_b + (size_t)i * Cfg::kBBytes, &tmap_b, &b_full[0], tap * p.Ci + chunk * BLOCK_K, n0);
          }
        }
        __syncwarp();
      }
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int b = tile >> p.tpb_shift;
        const int tr = tile - (b << p.tpb_shift);
        const int ty = tr >> p.tx_shift, tx = tr - (ty << p.tx_shift);
        const int x0 = tx * HALO_BW, y0 = ty * HALO_BH;
        for (int chunk = 0; chunk < p.kchunks; ++chunk) {
          if (UPS) {
            // low-resolution source box of the halo box: rows y0/2-1 .. y0/2+8, columns x0/2-1 .. x0/2+4 (zero-filled
            // outside the image; the producers clamp instead, like torch's bilinear kernel)
            mbar_wait(&s_empty[as], aph ^ 1, 10);
            if (elect_one()) {
              mbar_arrive_expect_tx(&s_full[as], Cfg::kSrcBytes);
              tma_load_4d(smem_src + as * Cfg::kSrcBytes, &tmap_a, &s_full[as], chunk * BLOCK_K, x0 / 2 - 1, y0 / 2 - 1, b);
            }
            __syncwarp();
            if (++as == UPS_S_STAGES) { as = 0; aph ^= 1; }
          } else {
            mbar_wait(&a_empty[as], aph ^ 1, 10);
            if (elect_one()) {
              if (dbg & 4) {
                mbar_arrive(&a_full[as]);
              } else {
                mbar_arrive_expect_tx(&a_full[as], Cfg::kATx);
                tma_load_4d(smem_a + as * Cfg::kABytes, &tmap_a, &a_full[as], chunk * BLOCK_K, x0 - 1, y0 - 1, b);
              }
            }
            __syncwarp();
            if (++as == A_STAGES) { as = 0; aph ^= 1; }
          }
          if (!RESIDENT_B) {
#pragma unroll 1
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&b_empty[bs], bph ^ 1, 11);
              if (elect_one()) {
                mbar_arrive_expect_tx(&b_full[bs], Cfg::kBBytes);
                tma_load_2d(smem_b + bs * Cfg::kBBytes, &tmap_b, &b_full[bs], tap * p.Ci + chunk * BLOCK_K, n0);
              }
              __syncwarp();
              if (++bs == B_STAGES) { bs = 0; bph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp_id == 1) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
    {
      constexpr uint32_t idesc = make_idesc(BLOCK_N);
      constexpr uint32_t sbo = HALO_W * Cfg::kRowBytes;
      if (RESIDENT_B) {
        mbar_wait(&b_full[0], 0, 12);
        tc_fence_after();
      }
      int as = 0, bs = 0, acc = 0;
      uint32_t aph = 0, bph = 0, accph = 0;
      const uint32_t a_base0 = smem_u32(smem_a), b_base0 = smem_u32(smem_b);
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        mbar_wait(&tmem_empty[acc], accph ^ 1, 13);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int chunk = 0; chunk < p.kchunks; ++chunk) {
          mbar_wait(&a_full[as], aph, 14);
          tc_fence_after();
          // descriptors of tap 0; every other tap / k step is this plus a compile-time constant (fully unrolled)
          const uint64_t da0 = make_smem_desc_sbo<BLOCK_K>(a_base0 + (uint32_t)(as * Cfg::kABytes), sbo);
          const uint64_t db_res = make_smem_desc<BLOCK_K>(b_base0 + (uint32_t)(chunk * 9 * Cfg::kBBytes));
          if (RESIDENT_B) {
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
            }
            __syncwarp();
          } else {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&b_full[bs], bph, 15);
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
                umma_commit(&b_empty[bs]);
              }
              __syncwarp();
              if (++bs == B_STAGES) { bs = 0; bph ^= 1; }
            }
          }
          if (elect_one()) umma_commit(&a_empty[as]);
          __syncwarp();
          if (++as == A_STAGES) { as = 0; aph ^= 1; }
        }
        if (elect_one()) umma_commit(&tmem_full[acc]);
        __syncwarp();
        acc ^= 1;
        if (acc == 0) accph ^= 1;
      }
    }
  } else if (UPS && warp_id >= 6) {
    // ===================== fused bilinear 2x upsample: A-operand producers (warps 6..9) =====================
    // out halo pixel (hy, hx) <-> image pixel (y0-1+hy, x0-1+hx); y0, x0 are even, so halo rows (2k, 2k+1) interpolate
    // source-box rows (k, k+1) with weights (.75,.25) / (.25,.75) (align_corners=False, scale 2: src = o/2 - 0.25),
    // and likewise for columns.  Source indices are clamped to the image (torch's border rule); halo pixels outside
    // the image are the conv's zero padding.  Thread = one 16-byte channel group x one column pair x one third of
    // the rows; horizontal pass per source row, then the vertical pass, all in fp32, one rounding to bf16.
    const int pt = threadIdx.x - NUM_THREADS;   // 0..127
    const int cg = pt & 7;
    const int unit = pt >> 3;                   // 0..15; unit 15 has no work
    const int j = unit % 5;                     // halo columns 2j, 2j+1 <- source columns j, j+1
    const int seg = unit / 5;                   // halo rows 6seg .. 6seg+5 <- source rows 3seg .. 3seg+3
    int as = 0, ss = 0;
    uint32_t aph = 0, sph = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int b = tile >> p.tpb_shift;
      const int tr = tile - (b << p.tpb_shift);
      const int ty = tr >> p.tx_shift, tx = tr - (ty << p.tx_shift);
      const bool left = tx == 0, right = tx == p.tiles_x - 1, top = ty == 0, bottom = ty == p.tiles_y - 1;
      const int lo_c = left ? 1 : 0, hi_c = right ? UPS_SRC_W - 2 : UPS_SRC_W - 1;
      const int lo_r = top ? 1 : 0, hi_r = bottom ? UPS_SRC_H - 2 : UPS_SRC_H - 1;
      const int c0 = min(max(j, lo_c), hi_c), c1 = min(max(j + 1, lo_c), hi_c);
      const bool zero_e = left && j == 0;        // halo column 0 is outside the image
      const bool zero_o = right && j == 4;       // halo column 9 is outside the image
      for (int chunk = 0; chunk < p.kchunks; ++chunk) {
        mbar_wait(&s_full[ss], sph, 17);
        mbar_wait(&a_empty[as], aph ^ 1, 18);
        if (seg < 3) {
          const uint8_t* src = smem_src + ss * Cfg::kSrcBytes + cg * 16;
          uint8_t* dst = smem_a + as * Cfg::kABytes;
          float he0[8], ho0[8], he1[8], ho1[8];
          auto hrow = [&](int r, float* he, float* ho) {
            const int rr = min(max(r, lo_r), hi_r);
            float a[8], c[8];
            unpack(*reinterpret_cast<const uint4*>(src + (rr * UPS_SRC_W + c0) * Cfg::kRowBytes), a);
            unpack(*reinterpret_cast<const uint4*>(src + (rr * UPS_SRC_W + c1) * Cfg::kRowBytes), c);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              he[k] = fmaf(0.25f, c[k], 0.75f * a[k]);   // odd image column 2m+1
              ho[k] = fmaf(0.75f, c[k], 0.25f * a[k]);   // even image column 2m+2
            }
          };
          auto put = [&](int hy, int hx, const float* v, bool zero) {
            const int pix = hy * HALO_W + hx;
            uint4 pk;
            if (zero) pk = make_uint4(0u, 0u, 0u, 0u);
            else pack(v, pk);
            *reinterpret_cast<uint4*>(dst + pix * Cfg::kRowBytes + ((cg ^ (pix & 7)) << 4)) = pk;
          };
          hrow(3 * seg, he0, ho0);
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int k = 3 * seg + i;
            hrow(k + 1, he1, ho1);
            const bool zt = top && k == 0;        // halo row 0 is outside the image
            const bool zb = bottom && k == 8;     // halo row 17 is outside the image
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = fmaf(0.25f, he1[q], 0.75f * he0[q]);
            put(2 * k, 2 * j, v, zt || zero_e);
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = fmaf(0.25f, ho1[q], 0.75f * ho0[q]);
            put(2 * k, 2 * j + 1, v, zt || zero_o);
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = fmaf(0.75f, he1[q], 0.25f * he0[q]);
            put(2 * k + 1, 2 * j, v, zb || zero_e);
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = fmaf(0.75f, ho1[q], 0.25f * ho0[q]);
            put(2 * k + 1, 2 * j + 1, v, zb || zero_o);
#pragma unroll
            for (int q = 0; q < 8; ++q) { he0[q] = he1[q]; ho0[q] = ho1[q]; }
          }
        }
        // generic-proxy writes -> visible to the async proxy (tcgen05.mma reads shared memory through it)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&a_full[as]);
        mbar_arrive(&s_empty[ss]);
        if (++as == A_STAGES) { as = 0; aph ^= 1; }
        if (++ss == UPS_S_STAGES) { ss = 0; sph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const ConvEpilogue& ep = p.ep;
    const int et = threadIdx.x - 64;  // 0..127
    for (int i = et; i < BLOCK_N; i += 128) {
      s_nw[i] = ep.noise ? __ldg(ep.noise_w + n0 + i) : 0.f;
      s_nb[i] = ep.noise ? __ldg(ep.noise_b + n0 + i) : 0.f;
    }
    const int q = warp_id & 3;          // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;        // tile row = TMEM lane
    const int yy = r >> 3, xx = r & 7;
    int acc = 0;
    uint32_t accph = 0;
    auto tile_coords = [&](int tile, int& b, int& x, int& y) {
      b = tile >> p.tpb_shift;
      const int tr = tile - (b << p.tpb_shift);
      const int ty = tr >> p.tx_shift, tx = tr - (ty << p.tx_shift);
      x = tx * HALO_BW + xx;
      y = ty * HALO_BH + yy;
    };
    auto load_noise = [&](int b, int x, int y) -> float {
      if (!ep.noise) return 0.f;
      const int S = ep.noise_size;
      return __ldg(ep.noise + (long long)(ep.noise_batch == 1 ? 0 : b) * S * S + (long long)x * S + y);
    };
    constexpr bool fuse_rgb = FUSE_RGB;   // compile-time: the plain instantiation carries none of the ToRGB registers
    const long long HWl = (long long)p.H * p.W;
    auto load_rgb_prev = [&](int b, int x, int y, float* v) {
      v[0] = v[1] = v[2] = 0.f;
      if (fuse_rgb && ep.rgb_accumulate) {
        const float* src = ep.rgb_out + ((long long)b * 3 * p.H + y) * p.W + x;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = __ldg(src + c * HWl);
      }
    };
    // per-sample tables (demod coefficients, next-layer style + 1, fused-ToRGB weights) live in two shared-memory slots;
    // a sample change (rare: tiles are walked sample-major) fills the other slot and synchronises the four warps once
    auto write_tables = [&](int slot, int b) {
      for (int i = et; i < BLOCK_N; i += 128) {
        s_d[slot * BLOCK_N + i] = ep.dcoef ? __ldg(ep.dcoef + (long long)b * ep.dcoef_stride + n0 + i) : 1.f;
        s_m[slot * BLOCK_N + i] = ep.next_style ? __ldg(ep.next_style + (long long)b * ep.next_style_stride + n0 + i) + 1.f : 1.f;
      }
      if (fuse_rgb) {
        for (int i = et; i < 3 * BLOCK_N; i += 128) {
          const int o = i % BLOCK_N;
          s_rgbw[slot * 3 * BLOCK_N + i] =
              (__ldg(ep.rgb_style + (long long)b * ep.rgb_style_stride + o) + 1.f) * __ldg(ep.rgb_w + (i / BLOCK_N) * p.Co + o);
        }
      }
    };
    // the generator's configuration (activation on, bf16 NHWC or no feature-map output) takes the packed fast path
    const bool fast = ep.act != 0 && !ep.out_nchw_f32;
    int b = 0, x = 0, y = 0, cur_b = -1, slot = 1;
    float nz = 0.f, rgbp[3] = {0.f, 0.f, 0.f};
    if (tile_begin < tile_end) {
      tile_coords(tile_begin, b, x, y);
      nz = load_noise(b, x, y);
      load_rgb_prev(b, x, y, rgbp);
    }
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      if (b != cur_b) {
        slot ^= 1;
        write_tables(slot, b);
        cur_b = b;
        asm volatile("bar.sync 1, 128;" ::: "memory");   // tables (and, the first time, s_nw / s_nb) visible to all four warps
      }
      const float* dd = s_d + slot * BLOCK_N;
      const float* mm = s_m + slot * BLOCK_N;
      const float* rw = s_rgbw + slot * 3 * BLOCK_N;
      const long long pix = ((long long)b * p.H + y) * p.W + x;
      // prefetch the next tile's per-pixel operands (consumed after this tile's TMEM drain)
      int b2 = b, x2 = 0, y2 = 0;
      float nz2 = 0.f, rgbp2[3] = {0.f, 0.f, 0.f};
      if (tile + 1 < tile_end) {
        tile_coords(tile + 1, b2, x2, y2);
        nz2 = load_noise(b2, x2, y2);
        load_rgb_prev(b2, x2, y2, rgbp2);
      }
      mbar_wait(&tmem_full[acc], accph, 16);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
      float rgb_acc[3] = {rgbp[0], rgbp[1], rgbp[2]};
      if (dbg & 1) {
      } else if (fast) {
        constexpr int NCH = BLOCK_N / 32;
        uint32_t v[NCH > 1 ? 2 : 1][32];
        uint64_t racc[3] = {pk2(rgbp[0], 0.f), pk2(rgbp[1], 0.f), pk2(rgbp[2], 0.f)};
        tmem_ld32(tbase, v[0]);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const int c0 = ch * 32;
          tmem_ld_wait();
          if (ch + 1 < NCH) tmem_ld32(tbase + (uint32_t)(c0 + 32), v[(ch + 1) & 1]);   // next chunk in flight during the math
          uint32_t om[16], orw[16];
          epi_fast32<fuse_rgb>(v[ch & 1], dd + c0, s_nw + c0, s_nb + c0, mm + c0, nz, om, orw, rw + c0, BLOCK_N, racc);
          if (ep.out) {
            uint4* out = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + pix * p.Co + n0 + c0);
#pragma unroll
            for (int k = 0; k < 4; ++k) out[k] = make_uint4(om[4 * k], om[4 * k + 1], om[4 * k + 2], om[4 * k + 3]);
          }
          if (ep.out_raw) {
            uint4* out = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out_raw) + pix * p.Co + n0 + c0);
#pragma unroll
            for (int k = 0; k < 4; ++k) out[k] = make_uint4(orw[4 * k], orw[4 * k + 1], orw[4 * k + 2], orw[4 * k + 3]);
          }
        }
        if (fuse_rgb) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float lo, hi;
            upk2(racc[c], lo, hi);
            rgb_acc[c] = lo + hi;
          }
        }
      } else {
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tbase + (uint32_t)c0, v);
          tmem_ld_wait();
          float f[32], fr[32];
          epi_chunk32(v, dd + c0, s_nw + c0, s_nb + c0, mm + c0, nz, ep.act, !ep.out_nchw_f32, f, fr);
          if (fuse_rgb) rgb_chunk32(fr, rw + c0, BLOCK_N, rgb_acc);
          if (!ep.out) {
          } else if (ep.out_nchw_f32) {
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
          if (ep.out_raw) {
            __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(ep.out_raw) + pix * p.Co + n0 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 pk;
              pack(fr + j, pk);
              *reinterpret_cast<uint4*>(out + j) = pk;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);   // 128 arrivals release the accumulator to the MMA warp
      if (fuse_rgb) {
        float* dst = ep.rgb_out + ((long long)b * 3 * p.H + y) * p.W + x;
#pragma unroll
        for (int c = 0; c < 3; ++c) dst[c * HWl] = rgb_acc[c];
      }
      acc ^= 1;
      if (acc == 0) accph ^= 1;
      b = b2; x = x2; y = y2; nz = nz2;
      rgbp[0] = rgbp2[0]; rgbp[1] = rgbp2[1]; rgbp[2] = rgbp2[2];
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

template <int BLOCK_N, int BLOCK_K, int A_STAGES, int B_STAGES, bool RESIDENT_B, bool FUSE_RGB, bool UPS>
int launch_conv_halo_cfg2(const __nv_bfloat16* x, const __nv_bfloat16* wk, ConvHaloParams p, cudaStream_t stream);

template <int BLOCK_N, int BLOCK_K, int A_STAGES, int B_STAGES, bool RESIDENT_B>
int launch_conv_halo_cfg(const __nv_bfloat16* x, const __nv_bfloat16* wk, ConvHaloParams p, cudaStream_t stream) {
  if (p.ep.rgb_style) return launch_conv_halo_cfg2<BLOCK_N, BLOCK_K, A_STAGES, B_STAGES, RESIDENT_B, true, false>(x, wk, p, stream);
  return launch_conv_halo_cfg2<BLOCK_N, BLOCK_K, A_STAGES, B_STAGES, RESIDENT_B, false, false>(x, wk, p, stream);
}

// x: the conv input [B,H,W,Ci] -- or, for UPS, the low-resolution tensor [B,H/2,W/2,Ci] the kernel upsamples itself
template <int BLOCK_N, int BLOCK_K, int A_STAGES, int B_STAGES, bool RESIDENT_B, bool FUSE_RGB, bool UPS>
int launch_conv_halo_cfg2(const __nv_bfloat16* x, const __nv_bfloat16* wk, ConvHaloParams p, cudaStream_t stream) {
  using Cfg = HaloCfg<BLOCK_N, BLOCK_K, A_STAGES, B_STAGES, RESIDENT_B, UPS>;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(SX_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  const CUtensorMapSwizzle swz = BLOCK_K == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMap ta, tb;
  {
    const int IH = UPS ? p.H / 2 : p.H, IW = UPS ? p.W / 2 : p.W;
    cuuint64_t gdim[4] = {(cuuint64_t)p.Ci, (cuuint64_t)IW, (cuuint64_t)IH, (cuuint64_t)p.B};
    cuuint64_t gstr[3] = {(cuuint64_t)p.Ci * 2, (cuuint64_t)IW * p.Ci * 2, (cuuint64_t)IH * IW * p.Ci * 2};
    cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)(UPS ? UPS_SRC_W : HALO_W), (cuuint32_t)(UPS ? UPS_SRC_H : HALO_H), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
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
  auto kern = conv_tc_halo_kernel<BLOCK_N, BLOCK_K, A_STAGES, B_STAGES, RESIDENT_B, FUSE_RGB, UPS>;
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
    // 512 TMEM columns.
    SX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    cudaFuncAttributes fa;
    SX_CUDA(cudaFuncGetAttributes(&fa, kern));
    const int regs = (fa.numRegs + 7) / 8 * 8;
    const int occ_smem = (int)((227 * 1024) / (smem + 1024));
    const int occ_regs = 65536 / (regs * Cfg::kThreads);
    const int occ_tmem = 512 / Cfg::kTmemCols;
    int occ = occ_smem < occ_regs ? occ_smem : occ_regs;
    occ = occ < occ_tmem ? occ : occ_tmem;
    occ_cached = occ < 1 ? 1 : (occ > 2 ? 2 : occ);
  }
  int grid_x = occ_cached * num_sms();
  if (grid_x > p.num_tiles) grid_x = p.num_tiles;
  dim3 grid((unsigned)grid_x, (unsigned)(p.Co / BLOCK_N));
  kern<<<grid, Cfg::kThreads, smem, stream>>>(ta, tb, p);
  SX_CHECK_LAUNCH();
  return SX_OK;
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
  p.ep = ep;
  return p;
}

// *handled = false (and SX_OK) when the halo kernel does not cover the shape
inline int launch_conv_halo(const __nv_bfloat16* x, const __nv_bfloat16* wk, int B, int Ci, int Co, int H, int W,
                            const ConvEpilogue& ep, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (!halo_shape_supported(Ci, Co, H, W) || B == 0) return SX_OK;
  const int bk = Ci % 64 == 0 ? 64 : 32;
  const ConvHaloParams p = make_halo_params(B, Ci, Co, H, W, bk, ep);
  const size_t weight_bytes = (size_t)9 * Ci * Co * 2;
  const bool resident = weight_bytes <= 80 * 1024;
  *handled = true;
  if (Co == 32 && bk == 64 && resident) return launch_conv_halo_cfg<32, 64, 3, 2, true>(x, wk, p, stream);
  if (Co == 32 && bk == 32 && resident) return launch_conv_halo_cfg<32, 32, 3, 2, true>(x, wk, p, stream);
  if (Co == 64 && bk == 64 && resident) return launch_conv_halo_cfg<64, 64, 3, 2, true>(x, wk, p, stream);
  // 128 -> 64 channels (147 KB of weights): still resident, with a 2-deep activation ring (197 KB of smem, 1 CTA/SM)
  if (Co == 64 && bk == 64 && weight_bytes <= 150 * 1024) return launch_conv_halo_cfg<64, 64, 2, 2, true>(x, wk, p, stream);
  if (Co == 64 && bk == 64) return launch_conv_halo_cfg<64, 64, 3, 6, false>(x, wk, p, stream);
  if (Co == 128 && bk == 64) return launch_conv_halo_cfg<128, 64, 3, 4, false>(x, wk, p, stream);
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
  if (Co == 32) return launch_conv_halo_cfg2<32, 64, 2, 2, true, false, true>(xlow, wk, p, stream);
  if (Co == 64 && weight_bytes <= 150 * 1024) return launch_conv_halo_cfg2<64, 64, 2, 2, true, false, true>(xlow, wk, p, stream);
  if (Co == 64) return launch_conv_halo_cfg2<64, 64, 2, 6, false, false, true>(xlow, wk, p, stream);
  return launch_conv_halo_cfg2<128, 64, 3, 4, false, false, true>(xlow, wk, p, stream);
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
