// common.cuh -- error plumbing, launch accounting and small device helpers shared by all kernels.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <map>
#include <mutex>
#include <vector>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/stylex_b200.h"

namespace sx {

// ---- errors ------------------------------------------------------------------------------------
inline std::string& last_error() {
  static thread_local std::string e;
  return e;
}

inline int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

#define SX_CUDA(expr)                                                                            \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return ::sx::fail(SX_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

#define SX_CHECK_LAUNCH()                                                                        \
  do {                                                                                           \
    ::sx::launch_counter().fetch_add(1, std::memory_order_relaxed);                              \
    cudaError_t _e = cudaGetLastError();                                                         \
    if (_e != cudaSuccess)                                                                       \
      return ::sx::fail(SX_ECUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
  } while (0)

#define SX_TRY(expr)          \
  do {                        \
    int _r = (expr);          \
    if (_r != SX_OK) return _r; \
  } while (0)

#define SX_REQUIRE(cond, ...)                       \
  do {                                              \
    if (!(cond)) return ::sx::fail(SX_EINVAL, __VA_ARGS__); \
  } while (0)

inline std::atomic<unsigned long long>& launch_counter() {
  static std::atomic<unsigned long long> c{0};
  return c;
}

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// grid for a grid-stride elementwise kernel: enough CTAs for the work, capped at a few waves of the
// 148 SMs (multiple of the SM count so the last wave is full).
inline int ew_grid(long long work_items, int threads, int ctas_per_sm = 8) {
  long long need = (work_items + threads - 1) / threads;
  long long cap = (long long)num_sms() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- optional per-kernel timing (bench.py roofline): CUDA events on the launching stream -------------
// kinds: 0..19 conv index of the generator plan, 32 modulate, 33 upsample, 34 torgb, 35 demod, 36 styles,
//        40 make_styles, 41 scatter, 42 select, 43 minmax, 50 op-level conv, 51 other op-level
struct ProfRec {
  int kind;
  double flops, bytes;
  cudaEvent_t a, b;
};
struct Profiler {
  bool on = false;
  std::vector<ProfRec> recs;
  std::mutex m;
};
inline Profiler& profiler() {
  static Profiler p;
  return p;
}
struct ProfScope {
  bool active;
  ProfRec r;
  cudaStream_t st;
  ProfScope(int kind, double flops, double bytes, cudaStream_t s) : active(profiler().on), st(s) {
    if (!active) return;
    r.kind = kind; r.flops = flops; r.bytes = bytes;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) { active = false; return; }
    cudaEventRecord(r.a, st);
  }
  ~ProfScope() {
    if (!active) return;
    cudaEventRecord(r.b, st);
    std::lock_guard<std::mutex> g(profiler().m);
    profiler().recs.push_back(r);
  }
};

// ---- element types -----------------------------------------------------------------------------
template <typename T>
struct Elem;
template <>
struct Elem<float> {
  static constexpr int kVec = 4;  // elements per 16-byte vector
  using vec_t = float4;
};
template <>
struct Elem<__nv_bfloat16> {
  static constexpr int kVec = 8;
  using vec_t = uint4;
};

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte vector <-> float[kVec]
__device__ __forceinline__ void unpack(const float4& v, float* f) {
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
__device__ __forceinline__ void unpack(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void pack(const float* f, float4& v) { v = make_float4(f[0], f[1], f[2], f[3]); }
__device__ __forceinline__ void pack(const float* f, uint4& v) {
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
}

__device__ __forceinline__ float lrelu02(float v) { return v > 0.f ? v : 0.2f * v; }

// ---- epilogue shared by the FFMA and the tcgen05 convolution kernels -----------------------------
// out[b, pix, o] = next( act( acc * d[b,o] + noise[b|0, x, y] * nw[o] + nb[o] ) )
struct ConvEpilogue {
  const float* dcoef;        // [B, dcoef_stride] demodulation coefficients (pre-offset to this conv); null = none
  int dcoef_stride;
  const float* noise;        // [noise_batch, S, S] full-resolution noise map; null = no noise term
  int noise_batch;           // 1 (broadcast) or B
  int noise_size;            // S
  const float* noise_w;      // [Co]
  const float* noise_b;      // [Co]
  int act;                   // 1: leaky_relu(0.2)
  const float* next_style;   // [B, next_style_stride] style of the consuming conv (pre-offset): out *= (s+1); null = none
  int next_style_stride;
  void* out;                 // NHWC in the activation dtype, or NCHW fp32 (out_nchw_f32)
  int out_nchw_f32;
  void* out_raw;             // optional NHWC copy WITHOUT the next_style factor (clean-prefix cache); null = none
  // fused ToRGB (RGBBlock ST:618-624) in the epilogue of a block's conv2 (tcgen05 kernels, one N tile, one sample per
  // M tile):  rgb_out[b,c,y,x] (+)= sum_o act[b,y,x,o] * (rgb_style[b,o] + 1) * rgb_w[c,o]
  // rgb_out is pre-filled with blur(upsample2x(previous rgb)) when rgb_accumulate != 0.  `out` may then be null
  // (last block: nothing downstream reads the feature map).
  const float* rgb_style;    // [B, rgb_style_stride] (pre-offset); null = no fused ToRGB
  int rgb_style_stride;
  const float* rgb_w;        // [3][Co]
  float* rgb_out;            // [B,3,H,W] planar fp32
  int rgb_accumulate;
};

}  // namespace sx
