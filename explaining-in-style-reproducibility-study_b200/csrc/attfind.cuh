// attfind.cuh -- the AttFind sweep / selection kernels (reference notebook cells 5 and 15).
//
//   minmax          get_min_max_style_vectors           NB:237-252
//   make_styles     the per-coordinate shift injection   NB:358-381 (as data, never by patching weights)
//   scatter_effects logit delta store                     NB:385
//   select_*        class split + greedy top-k            NB:695-714, NB:731-758
//
// Selection exactness: numpy evaluates np.mean(E[mask], axis=0) on a float64 C-contiguous matrix, i.e. a
// sequential row-by-row float64 accumulation per column followed by one division.  colmean_kernel does
// exactly that (one thread per column, rows in image order), so the column means -- and therefore every
// argmax including its first-index tie-break -- are bit-identical to the notebook's.
#pragma once

#include "common.cuh"

namespace sx {

__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ sc, int N, int S, int stride,
                                                     float* __restrict__ mn, float* __restrict__ mx) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float lo = sc[s], hi = lo;
  for (int n = 1; n < N; ++n) {
    const float v = sc[(long long)n * stride + s];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  mn[s] = lo;
  mx[s] = hi;
}

// one CTA row per coord-eval j: copy the base style row, then overwrite coordinate s_j
__global__ void __launch_bounds__(256) make_styles_kernel(const float* __restrict__ base, const float* __restrict__ mn,
                                                          const float* __restrict__ mx, float* __restrict__ out, int row,
                                                          int first_s, float shift_size) {
  const int j = blockIdx.x;
  const int s = first_s + (j >> 1);
  const int d = j & 1;
  float* o = out + (long long)j * row;
  for (int i = threadIdx.x; i < row; i += blockDim.x) {
    float v = base[i];
    if (i == s) {
      const float target = d == 0 ? mn[s] : mx[s];
      v = v + (target - v) * shift_size;  // NB:374-375,381
    }
    o[i] = v;
  }
}

// the same for arbitrary (latent n_j, flat column x_j = d_j * S + s_j) pairs: row j starts as the style row of latent n_j
__global__ void __launch_bounds__(256) make_styles_pairs_kernel(const float* __restrict__ styles_all, long long row_stride,
                                                                const float* __restrict__ mn, const float* __restrict__ mx,
                                                                float* __restrict__ out, int row, int S,
                                                                const int* __restrict__ latent_idx, const int* __restrict__ columns,
                                                                float shift_size) {
  const int j = blockIdx.x;
  const int x = columns[j];
  const int s = x % S, d = x / S;
  const float* base = styles_all + (long long)latent_idx[j] * row_stride;
  float* o = out + (long long)j * row;
  for (int i = threadIdx.x; i < row; i += blockDim.x) {
    float v = base[i];
    if (i == s) {
      const float target = d == 0 ? mn[s] : mx[s];
      v = v + (target - v) * shift_size;  // NB:374-375,381
    }
    o[i] = v;
  }
}

__global__ void __launch_bounds__(256) scatter_effects_kernel(const float* __restrict__ logits, const float* __restrict__ base,
                                                              float* __restrict__ effects, int n, int S, int first_s,
                                                              int count) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  const int s = first_s + (j >> 1), d = j & 1;
  const float b0 = base[2 * n], b1 = base[2 * n + 1];
  float2 v = make_float2(logits[2 * j] - b0, logits[2 * j + 1] - b1);
  *reinterpret_cast<float2*>(effects + (((long long)n * 2 + d) * S + s) * 2) = v;
}

// ---- selection ----------------------------------------------------------------------------------
struct SelectState {  // lives in the caller's workspace
  double* colmean;       // [2S]
  double* images_effect; // [N]
  int* row_class;        // [N]  argmax(base_logits) (first max wins, like np.argmax)
  int* picked;           // [k]  flat column indices chosen so far
  int* num_picked;       // [1]
  unsigned long long* best;  // [1] scratch for the argmax reduction
};

__global__ void select_init_kernel(const float* __restrict__ base_logits, int N, int class_index, SelectState st) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n == 0) *st.num_picked = 0;
  if (n >= N) return;
  st.images_effect[n] = 0.0;
  st.row_class[n] = base_logits ? (base_logits[2 * n + 1] > base_logits[2 * n] ? 1 : 0) : class_index;
}

// colmean[x] = mean over rows {n : class(n)==c, images_effect[n] < max_effect} of E[n][x],
// E[n][x] = 0 if x already picked else max(0, effects[n, d, s, c]);  NaN when the mask is empty (0/0).
template <typename E>
__global__ void __launch_bounds__(128) colmean_kernel(const E* __restrict__ effects, int N, int S, int c, double max_effect,
                                                      SelectState st) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= 2 * S) return;
  bool was_picked = false;
  const int np = *st.num_picked;
  for (int i = 0; i < np; ++i) was_picked |= (st.picked[i] == x);
  double acc = 0.0;
  long long cnt = 0;
  for (int n = 0; n < N; ++n) {
    if (st.row_class[n] != c || !(st.images_effect[n] < max_effect)) continue;
    ++cnt;
    if (!was_picked) {
      const double v = (double)__ldg(effects + ((long long)n * 2 * S + x) * 2 + c);
      acc += v > 0.0 ? v : 0.0;  // np.maximum(0, .) on the float64 copy (NB:703-707,745)
    }
  }
  st.colmean[x] = cnt > 0 ? acc / (double)cnt : __longlong_as_double(0x7ff8000000000000LL);
}

// np.argmax semantics: first maximum; a NaN counts as the maximum (first NaN wins).  One CTA.
template <typename E>
__global__ void __launch_bounds__(1024) argmax_update_kernel(const E* __restrict__ effects, int N, int S, int c,
                                                             SelectState st, int* __restrict__ picks_out, int round) {
  __shared__ double s_val[32];
  __shared__ int s_idx[32];
  __shared__ int s_best;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double best = 0.0;
  int bi = -1;
  bool best_nan = false;
  for (int x = tid; x < 2 * S; x += blockDim.x) {  // ascending x per thread
    const double v = st.colmean[x];
    const bool vn = v != v;
    if (bi < 0 || (!best_nan && (vn || v > best))) {
      best = v; bi = x; best_nan = vn;
    }
  }
  auto better = [](double v, int i, double bv, int b_i) {  // is (v,i) ahead of (bv,b_i)?
    if (i < 0) return false;
    if (b_i < 0) return true;
    const bool vn = v != v, bn = bv != bv;
    if (vn != bn) return vn;
    if (vn && bn) return i < b_i;
    if (v != bv) return v > bv;
    return i < b_i;
  };
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, best, off);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
    if (better(ov, oi, best, bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) { s_val[wid] = best; s_idx[wid] = bi; }
  __syncthreads();
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    best = lane < nw ? s_val[lane] : 0.0;
    bi = lane < nw ? s_idx[lane] : -1;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (better(ov, oi, best, bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) {
      s_best = bi;
      picks_out[2 * round] = bi / S;      // direction   NB:758
      picks_out[2 * round + 1] = bi % S;  // sindex
    }
  }
  __syncthreads();
  const int x = s_best;
  // images_effect += E[:, x]; E[:, x] = 0     NB:755-756 (only rows of this class exist in the notebook's matrix)
  bool was_picked = false;
  const int np = *st.num_picked;
  for (int i = 0; i < np; ++i) was_picked |= (st.picked[i] == x);
  if (!was_picked) {
    for (int n = tid; n < N; n += blockDim.x) {
      if (st.row_class[n] != c) continue;
      const double v = (double)effects[((long long)n * 2 * S + x) * 2 + c];
      st.images_effect[n] += v > 0.0 ? v : 0.0;
    }
  }
  __syncthreads();
  if (tid == 0) {
    st.picked[np] = x;
    *st.num_picked = np + 1;
  }
}

}  // namespace sx
