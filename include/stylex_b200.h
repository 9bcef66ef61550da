/*
 * stylex_b200.h -- C ABI of libstylex_b200.so: the B200 (sm_100a) kernels under the reference's
 * AttFind hot path.  Plain pointers and sizes only; no torch / ATen / pybind types.
 *
 * The reference (NoahVl/Explaining-In-Style-Reproducibility-Study, "R/") is pure Python and has no
 * FFI of its own; its boundary is the Python class surface of R/stylex/stylex_train.py ("ST") and the
 * AttFind notebook R/stylex/run_attfind_combined.ipynb ("NB", raw JSON line numbers).  Each entry
 * point below names the reference interface it sits under.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into memory owned by the caller (PyTorch), unless it is
 *     documented as host; the library never frees or keeps caller memory beyond the call, except
 *     sx_generator_load(), which copies (packs) the weights into memory owned by the handle.
 *   - all work is enqueued on `stream` (a cudaStream_t; pass torch.cuda.current_stream().cuda_stream)
 *     and no entry point synchronises the device (sx_generator_load / sx_tc_selftest excepted).
 *   - return value: SX_OK (0) or a negative SX_E* code; the message is in sx_last_error()
 *     (thread local).  Nothing throws, nothing exits.  There is NO CPU fallback: a missing GPU, a
 *     non-sm_100 device or an unsupported shape is an error.
 *   - boundary tensors use the reference's layout (NCHW, fp32).  Inside the generator plan
 *     activations are NHWC in fp32 (SX_PREC_FP32, CUDA-core FFMA, <=1e-4 parity mode) or bf16
 *     (SX_PREC_BF16, tcgen05/TMEM/TMA implicit GEMM with fp32 accumulation).
 */
#ifndef STYLEX_B200_H
#define STYLEX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SX_VERSION 102

#define SX_OK 0
#define SX_EINVAL (-1)       /* bad argument */
#define SX_ECUDA (-2)        /* CUDA runtime / driver error */
#define SX_EUNSUPPORTED (-3) /* shape or device not supported by the sm_100a kernels */
#define SX_ENOMEM (-4)       /* caller workspace too small */
#define SX_ESTATE (-5)       /* handle not loaded / cache not primed */

#define SX_PREC_FP32 0
#define SX_PREC_BF16 1

#define SX_MAX_BLOCKS 10

typedef void* sx_stream_t;             /* cudaStream_t */
typedef struct sx_generator sx_generator_t;

int sx_version(void);
const char* sx_last_error(void);
/* 0 when the current device is a compute-capability 10.x GPU, SX_EUNSUPPORTED / SX_ECUDA otherwise. */
int sx_device_check(void);
/* number of kernels this library has launched in this process (bench.py "gpu_launches"). */
unsigned long long sx_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * L1 ops in the reference's tensor layout (NCHW fp32).  These sit under the nn.Module forwards.
 * ---------------------------------------------------------------------------------------------- */

/* Conv2DMod.forward(x, y) -- ST:647-667.  x[B,Ci,H,W], weight[Co,Ci,k,k] (k = 1 or 3), style[B,Ci]
 * -> out[B,Co,H,W].  out = d[b,o] * conv(W, x * (style+1)), d = rsqrt(sum((W*(style+1))^2) + eps) if
 * demod else 1 ("same" zero padding, stride 1, dilation 1).  precision selects the FFMA fp32 kernel or
 * the bf16 tcgen05 kernel (needs k == 3, Ci % 32 == 0, Co in {32,64,128,256,512}). */
size_t sx_conv2dmod_workspace_bytes(int B, int Ci, int Co, int H, int W, int k, int precision);
int sx_conv2dmod_fwd(const float* x, const float* weight, const float* style, float* out,
                     int B, int Ci, int Co, int H, int W, int k, int demod, float eps, int precision,
                     void* workspace, size_t workspace_bytes, sx_stream_t stream);

/* First-order backward of Conv2DMod.forward (what autograd derives for ST:647-667; training step).
 * Inputs: the forward's x, weight, style, its output `out` (read only when demod != 0) and grad_out = dL/dout
 * [B,Co,H,W].  Outputs (all written): grad_x [B,Ci,H,W], grad_weight [Co,Ci,k,k] (summed over the batch: the weight is
 * shared), grad_style [B,Ci] (through the modulation and, with demod, through the demodulation coefficients).
 * precision SX_PREC_FP32: FFMA implicit GEMMs.  SX_PREC_BF16: dgrad on the tcgen05 conv kernels (transposed weights) and
 * wgrad as a tcgen05 GEMM over K = pixels where the shapes allow (dgrad: k == 3, channel counts as in sx_conv2dmod_fwd;
 * wgrad: square power-of-two maps >= 64), FFMA otherwise; bf16 operands, fp32 accumulation and fp32 per-sample reductions.
 * Deterministic (fixed-order split-K reduction, no atomics).  B == 0 writes a zero grad_weight. */
size_t sx_conv2dmod_bwd_workspace_bytes(int B, int Ci, int Co, int H, int W, int k, int precision);
int sx_conv2dmod_bwd(const float* x, const float* weight, const float* style, const float* out, const float* grad_out,
                     float* grad_x, float* grad_weight, float* grad_style,
                     int B, int Ci, int Co, int H, int W, int k, int demod, float eps, int precision,
                     void* workspace, size_t workspace_bytes, sx_stream_t stream);

/* nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False) -- ST:614,679.  [B,C,H,W] -> [B,C,2H,2W] */
int sx_upsample2x_bilinear(const float* x, float* out, int B, int C, int H, int W, sx_stream_t stream);

/* Blur.forward -- ST:144-153 (kornia filter2d, [1,2,1]x[1,2,1]/16, reflect border).  Same shape. */
int sx_blur3x3_reflect(const float* x, float* out, int B, int C, int H, int W, sx_stream_t stream);

/* GeneratorBlock noise + activation -- ST:696-698,705,714:
 * out[b,c,y,x] = leaky_relu_0.2( x[b,c,y,x] + inoise[b|0, x, y] * noise_w[c] + noise_b[c] )
 * inoise is the full-resolution map [noise_batch, noise_size, noise_size] (the trailing 1 of the
 * reference's [B,S,S,1] dropped); note the TRANSPOSED spatial indexing of the reference (quirk Q1). */
int sx_noise_lrelu(const float* x, const float* inoise, const float* noise_w, const float* noise_b, float* out,
                   int B, int C, int H, int W, int noise_batch, int noise_size, sx_stream_t stream);

/* RGBBlock tail -- ST:623-627: out = [blur(upsample2x(] rgb + prev [))].  prev may be NULL.
 * rgb/prev [B,C,H,W]; out [B,C,2H,2W] when upsample != 0 else [B,C,H,W]. */
int sx_rgb_add_upsample_blur(const float* rgb, const float* prev, float* out, int B, int C, int H, int W,
                             int upsample, sx_stream_t stream);

/* RGBBlock.upsample alone -- ST:613-616,626-627: out[b] = blur(upsample2x(prev[b | 0])) for 3-channel planes, the pre-fill
 * of the rgb plane a fused-ToRGB conv2 epilogue accumulates into (generator plan).  prev [prev_batch,3,h,w] with
 * prev_batch 1 (one cached plane broadcast to the batch, AttFind suffix forwards) or B; out [B,3,2h,2w]. */
int sx_rgb_prefill_upsample_blur(const float* prev, int prev_batch, float* out, int B, int h, int w, sx_stream_t stream);

/* nn.Linear -- to_style1/2, RGBBlock.to_style (ST:608,681,685).  x[B,in], weight[out,in], bias[out] -> out[B,out] */
int sx_linear_fwd(const float* x, const float* weight, const float* bias, float* out, int B, int in_features,
                  int out_features, sx_stream_t stream);

/* ---- first-order backward of the ops above (training step).  All deterministic gathers. ---- */

/* nn.Linear: grad_x[B,in] = g W, grad_weight[out,in] = g^T x, grad_bias[out] = sum_b g (grad_bias may be NULL). */
int sx_linear_bwd(const float* x, const float* weight, const float* grad_out, float* grad_x, float* grad_weight,
                  float* grad_bias, int B, int in_features, int out_features, sx_stream_t stream);

/* noise + leaky-ReLU (sx_noise_lrelu): from the forward OUTPUT and grad_out -> grad_x[B,C,H,W] and the gradients of
 * the to_noise Linear(1, C): grad_noise_w[C] = sum grad_x * inoise[b|0, x, y] (transposed, quirk Q1), grad_noise_b[C]. */
size_t sx_noise_lrelu_bwd_workspace_bytes(int B, int C);
int sx_noise_lrelu_bwd(const float* out, const float* grad_out, const float* inoise, float* grad_x, float* grad_noise_w,
                       float* grad_noise_b, int B, int C, int H, int W, int noise_batch, int noise_size,
                       void* workspace, size_t workspace_bytes, sx_stream_t stream);

/* adjoint of sx_upsample2x_bilinear: grad_out[B,C,2H,2W] -> grad_x[B,C,H,W]. */
int sx_upsample2x_bilinear_bwd(const float* grad_out, float* grad_x, int B, int C, int H, int W, sx_stream_t stream);

/* adjoint of sx_blur3x3_reflect (reflected border taps fold back inside): same shape. */
int sx_blur3x3_reflect_bwd(const float* grad_out, float* grad_x, int B, int C, int H, int W, sx_stream_t stream);

/* The tensor branch of ResNet.classify_images' preprocessing (resnet_classifier.py:60-68) in one pass:
 * torchvision resize(images, [OH,OW]) (bilinear, antialias=True: ATen's _upsample_bilinear2d_aa arithmetic) ->
 * Normalize(mean, std) (optional) -> cast -> channels_last.  in [B,3,IH,IW] fp32 NCHW; out [B,OH,OW,3] fp32 or bf16,
 * i.e. the memory of a channels_last [B,3,OH,OW] tensor.  mean3/std3: HOST arrays of 3 floats.  The network stays PyTorch. */
int sx_resize_aa_normalize(const float* in, void* out, int out_bf16, int B, int IH, int IW, int OH, int OW, int normalize,
                           const float* mean3, const float* std3, sx_stream_t stream);

/* Same preprocessing, written as the 2x2 space-to-depth image the re-expressed ResNet stem consumes (the 7x7/stride-2/
 * pad-3 convolution of torchvision's resnet conv1 == a 4x4/stride-1/pad-0 convolution on this tensor; classifiers.py):
 * out[b, Y, X, (dy*2+dx)*3 + c] = preprocessed[b, c, 2(Y-2)+dy, 2(X-2)+dx], zero outside the image and in channels
 * 12..15.  out: [B, OH/2+3, OW/2+3, 16] fp32 or bf16 = the memory of a channels_last [B,16,OH/2+3,OW/2+3] tensor. */
int sx_resize_aa_normalize_s2d(const float* in, void* out, int out_bf16, int B, int IH, int IW, int OH, int OW, int normalize,
                               const float* mean3, const float* std3, sx_stream_t stream);

/* torch.nn.functional.max_pool2d(x, kernel_size=3, stride=2, padding=1) on the memory of a channels_last [B,C,H,W]
 * tensor (= [B,H,W,C]); out [B,(H-1)/2+1,(W-1)/2+1,C].  The pool between the ResNet stem and layer1 (torchvision
 * resnet.py `self.maxpool`, reached from resnet_classifier.py:71) -- bit-identical to ATen's kernel, NaNs propagate.
 * C must be a multiple of 8 (bf16) / 4 (fp32). */
int sx_maxpool3x3s2_nhwc(const void* in, void* out, int is_bf16, int B, int H, int W, int C, sx_stream_t stream);

/* The ResNet stem of the classifier wrapper (resnet_classifier.py:56-71 -> torchvision resnet.py conv1 + bn1 + relu, and
 * with fuse_pool the 3x3/2/1 maxpool after it) as one tcgen05 kernel on the space-to-depth input that
 * sx_resize_aa_normalize_s2d writes.  x: bf16 [B, Hin, Win, 16]; w_taps: bf16 [4][4][64][16] = the BatchNorm-folded 7x7
 * stride-2 weights re-indexed as 4x4 stride-1 taps (ky, kx, co, ci); bias: fp32 [64]; out: bf16 NHWC
 * [B, Hin-3, Win-3, 64], or with fuse_pool [B, (Hin-4)/2+1, (Win-4)/2+1, 64].  Values: relu(conv + bias) accumulated in
 * fp32 and rounded to bf16 once (what cudnn_convolution_relu stores), then the max. */
int sx_stem_s2d_conv_relu(const void* x, const void* w_taps, const float* bias, void* out, int B, int Hin, int Win, int fuse_pool,
                          sx_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Generator plan -- sits under Generator.forward (ST:794-825) and the notebook's repeated
 * stylex.G(w, noise) calls (NB:318,382).
 * ---------------------------------------------------------------------------------------------- */

/* per-GeneratorBlock parameters, each a device pointer to an fp32 tensor laid out exactly as in the
 * reference's state dict. */
typedef struct sx_block_params {
  const float* to_style1_w;  /* [Ci, latent]   blocks.i.to_style1.weight */
  const float* to_style1_b;  /* [Ci]                                     */
  const float* to_noise1_w;  /* [Co, 1]        blocks.i.to_noise1.weight */
  const float* to_noise1_b;  /* [Co]                                     */
  const float* conv1_w;      /* [Co, Ci, 3, 3] blocks.i.conv1.weight     */
  const float* to_style2_w;  /* [Co, latent]                             */
  const float* to_style2_b;  /* [Co]                                     */
  const float* to_noise2_w;  /* [Co, 1]                                  */
  const float* to_noise2_b;  /* [Co]                                     */
  const float* conv2_w;      /* [Co, Co, 3, 3]                           */
  const float* rgb_style_w;  /* [Co, latent]   blocks.i.to_rgb.to_style.weight */
  const float* rgb_style_b;  /* [Co]                                     */
  const float* rgb_conv_w;   /* [3, Co, 1, 1]  blocks.i.to_rgb.conv.weight */
} sx_block_params;

/* ci/co: host arrays of the (input_channels, filters) pair of each GeneratorBlock (ST:775-792);
 * image_size = 4 << (num_blocks-1). */
int sx_generator_create(const int* ci, const int* co, int num_blocks, int latent_dim, sx_generator_t** out);
void sx_generator_destroy(sx_generator_t* g);

/* Copy + pack the weights (fp32 [tap][Ci][Co] for the FFMA kernel, bf16 [Co][tap*Ci] K-major for the
 * tcgen05 kernel, sum_k W^2 tables for demodulation), fold the batch-invariant
 * initial_conv(initial_block) (ST:802,806, quirk Q13).  Synchronises `stream` before returning. */
int sx_generator_load(sx_generator_t* g, const float* initial_block, const float* initial_conv_w,
                      const float* initial_conv_b, const sx_block_params* blocks, sx_stream_t stream);

/* StyleSpace width S = sum(Ci+Co) (the reference's num_style_coords, NB:471) and the width of the full
 * per-sample style row  [ style coords (S) | ToRGB styles (sum Co) ]. */
int sx_generator_num_style_coords(const sx_generator_t* g);
int sx_generator_style_row(const sx_generator_t* g);

/* K5: all W->S affine layers of all blocks at once.  w[B, num_blocks, latent] (the reference's `styles`
 * argument) -> styles[B, style_row].  The first S columns are exactly Generator.forward's
 * `style_coords` output (ST:709,821). */
int sx_generator_styles(const sx_generator_t* g, const float* w, float* styles, int B, sx_stream_t stream);

size_t sx_generator_workspace_bytes(const sx_generator_t* g, int max_batch, int precision);

/* Generator.forward from StyleSpace.  styles[B, style_row]; inoise[noise_batch(1|B), S, S];
 * rgb_out[B,3,S,S] fp32 NCHW, unclamped (ST:825).
 *   start_conv == 0         : full forward.  If save_cache != 0 (requires B == 1) the raw input of every
 *                             conv and every block's rgb are kept in the workspace (clean prefix).
 *   start_conv == c (> 0)   : AttFind suffix: conv c = 2*block + {0: conv1, 1: conv2} is the first one
 *                             whose style differs from the cached latent; everything before it is taken
 *                             from the cache primed by the last save_cache call on this workspace. */
int sx_generator_forward(sx_generator_t* g, const float* styles, const float* inoise, int noise_batch,
                         float* rgb_out, int B, int start_conv, int save_cache, int precision,
                         void* workspace, size_t workspace_bytes, sx_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * AttFind sweep + selection -- sits under attfind_extraction (NB:269-417) and
 * find_significant_styles (NB:731-758).
 * ---------------------------------------------------------------------------------------------- */

/* get_min_max_style_vectors -- NB:237-252.  style_coords[N, row_stride] (first S columns used). */
int sx_attfind_minmax(const float* style_coords, int N, int S, int row_stride, float* minima, float* maxima,
                      sx_stream_t stream);

/* The shift injection of NB:358-381 for a whole batch of coord-evals of ONE latent:
 * out[j, :] = base_row[:];  out[j, s_j] = base_row[s_j] + (m_dj[s_j] - base_row[s_j]) * shift_size
 * with s_j = first_sindex + j / 2, d_j = j % 2 (0: towards minima, 1: towards maxima), j < 2*num_coords. */
int sx_attfind_make_styles(const float* base_row, const float* minima, const float* maxima, float* out,
                           int style_row, int first_sindex, int num_coords, float shift_size, sx_stream_t stream);

/* The same shift injection for ARBITRARY (latent, direction, coordinate) triples -- the exact re-evaluation of the top-k
 * candidates (attfind.attfind_verify_topk), batched across latents: styles_all [N, row_stride] are the style rows of all
 * latents, latent_idx[j] / columns[j] (device int32) name the latent and the flat column d_j * S + s_j (the index of
 * NB:745 / 758); out[j, :] = styles_all[latent_idx[j], :style_row] with coordinate s_j moved towards minima (d_j = 0) or
 * maxima (d_j = 1) by shift_size (NB:374-381). */
int sx_attfind_make_styles_pairs(const float* styles_all, long long row_stride, const float* minima, const float* maxima,
                                 float* out, int style_row, int S, const int* latent_idx, const int* columns, int count,
                                 float shift_size, sx_stream_t stream);

/* NB:385: effects[n, d_j, s_j, :] = logits[j, :] - base_logits[n, :]  (effects [N,2,S,2], logits [2*num_coords, 2]) */
int sx_attfind_scatter_effects(const float* logits, const float* base_logits, float* effects, int n, int S,
                               int first_sindex, int num_coords, sx_stream_t stream);

/* find_significant_styles -- NB:731-758 + the class split NB:695-714, on device.
 * effects[N,2,S,2] (fp32, or fp64 when effects_f64 != 0 -- the notebook hands the function a float64 copy),
 * base_logits[N,2] fp32 (label = argmax, first max wins) or NULL = every row belongs to `class_index` (the
 * caller already split the classes, as the notebook does).  Greedy k rounds over the images of class
 * `class_index`; column means accumulate in float64 in image order exactly like numpy.
 * picks (device int32 [k*2]) receives (direction, sindex) pairs.  workspace: sx_attfind_select_workspace_bytes. */
size_t sx_attfind_select_workspace_bytes(int N, int S);
int sx_attfind_select(const void* effects, int effects_f64, const float* base_logits, int N, int S, int k,
                      double max_image_effect, int class_index, int* picks, void* workspace, size_t workspace_bytes,
                      sx_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Diagnostics
 * ---------------------------------------------------------------------------------------------- */
/* Per-kernel timing for bench.py's roofline: while enabled, every kernel of the generator plan is bracketed by
 * CUDA events on its own stream.  sx_profile_collect synchronises those events and writes one row of 5 doubles
 * per kernel kind: {kind, launches, total ms, algorithmic FLOPs, algorithmic bytes}; kinds 0..19 = conv index
 * (2*block + {0,1}) of the plan, 32 modulate, 33 upsample+modulate, 34 ToRGB, 35 demod.  rows/n_rows: HOST. */
int sx_profile_enable(int on);
int sx_profile_collect(double* rows, int max_rows, int* n_rows);

/* launches the tcgen05 kernel on a tiny problem and compares with the FFMA kernel; returns SX_OK when
 * max-abs error <= tol.  Host-synchronous.  max_err_out (host pointer) may be NULL. */
int sx_tc_selftest(float tol, float* max_err_out);

#ifdef __cplusplus
}
#endif
#endif /* STYLEX_B200_H */
