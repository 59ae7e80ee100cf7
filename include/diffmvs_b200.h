/*
 * diffmvs_b200 - C ABI of the B200 (sm_100a) kernels behind the DiffMVS / CasDiffMVS hot path.
 *
 * The reference (cvg/diffmvs) is pure PyTorch and has no FFI of its own; the
 * boundary it offers is the Python `models/` operator surface.  Every entry point below replaces the
 * ATen call sequence of one reference operator (file:line cited per function) and is bound from
 * `diffmvs_b200/_cabi.py` with ctypes - plain pointers and sizes only, no torch types.
 *
 * Conventions
 *   - all tensors are fp32, device memory, channels-last: 2-D maps are [N][H][W][C], volumes are
 *     [N][D][H][W][C] (exceptions are stated per function: uint8 images, the float64 camera matrices
 *     and integer masks of the fusion entry points, the host-side weight block of dmvs_conv3d_to1_f32).  A "pixel stride" (`*_ps`) is the distance in floats between consecutive
 *     pixels, so a channel slice of a wider buffer is addressed with ptr+offset and the wide stride
 *     (this is how every `torch.cat`/`torch.split` of the reference is made free).
 *   - every function only enqueues work on `stream` (a cudaStream_t passed as void*); it never
 *     allocates, synchronises or throws.  Return value: 0 on success, DMVS_ERR_* (<0) for invalid
 *     arguments, or a positive cudaError_t from the launch.
 *   - outputs are written in full; `*_stats` accumulators must be zeroed by the caller.
 */
#ifndef DIFFMVS_B200_H
#define DIFFMVS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMVS_ABI_VERSION 4

#define DMVS_ERR_ARG (-1)      /* null pointer / non-positive size / unsupported combination */
#define DMVS_ERR_ALIGN (-2)    /* pointer or stride not aligned as documented */
#define DMVS_ERR_UNSUPPORTED (-3)

/* activation codes */
#define DMVS_ACT_NONE 0
#define DMVS_ACT_RELU 1
#define DMVS_ACT_SIGMOID 2
#define DMVS_ACT_TANH 3
#define DMVS_ACT_SILU 4

/* residual placement */
#define DMVS_RES_NONE 0
#define DMVS_RES_PRE_ACT 1  /* y = act(conv + bias + res)   ResidualBlock, module.py:315-319 */
#define DMVS_RES_POST_ACT 2 /* y = act(conv + bias) + res   CostRegNet skips, module.py:445-446 */

/* arithmetic of dmvs_conv_f32 (storage is always fp32) */
#define DMVS_PREC_FP32 0   /* CUDA-core FFMA, fp32 products */
#define DMVS_PREC_TF32X3 1 /* tensor cores, each operand split hi+lo: a*b ~ ah*bh + al*bh + ah*bl (fp32-class) */
#define DMVS_PREC_TF32 2   /* tensor cores, operands rounded to TF32 (what cuDNN does by torch default) */
#define DMVS_PREC_TC_TF32X3 3 /* as TF32X3 on the tcgen05/TMEM back end (stride-1 layers; others fall back to 1) */
#define DMVS_PREC_AUTO 5      /* fp32-class; per layer: tcgen05 3xTF32 where it is faster, FFMA elsewhere */
#define DMVS_PREC_TC_TF32 4   /* as TF32 on the tcgen05/TMEM back end (stride-1 layers; others fall back to 2) */
#define DMVS_PREC_WS_TF32X3 6 /* 3xTF32 on the width-stacked tcgen05 back end (conv_ws.cu) wherever it applies, FFMA elsewhere */
#define DMVS_PREC_WS_TF32 7   /* plain TF32 on the width-stacked tcgen05 back end, legacy TF32 elsewhere */
#define DMVS_PREC_WS2_TF32X3 8 /* 3xTF32, width-stacked tcgen05 arithmetic behind the TMA-fed pipeline (conv_ws2.cu): one
                                  persistent CTA per SM, cp.async.bulk.tensor producer, split / MMA / epilogue warps, two
                                  TMEM accumulator sets; conv_ws.cu for nearest-upsampled inputs, FFMA elsewhere */

#define DMVS_PREC_WS2_TF32_F16C 9 /* as WS2_TF32X3 with the two correction products A_lo*B_hi + A_hi*B_lo computed by one
                                     kind::f16 MMA (K = 16) from fp16 copies of the correction operands: two MMAs per kernel
                                     row instead of three, the same 11 significant bits in every factor (`w_ws16` required;
                                     layers it does not cover - <= 4 input channels, no `w_ws16` - run as WS2_TF32X3) */

/* epilogue kinds */
#define DMVS_EPI_STD 0
#define DMVS_EPI_GRU_ZR 1 /* c<hid: z=sigmoid(v); c>=hid: r*h = sigmoid(v)*aux1[c-hid]   module.py:166-168 */
#define DMVS_EPI_GRU_Q 2  /* h' = (1-z)*h + z*tanh(v), z=aux1, h=aux2                     module.py:169-170 */

int dmvs_abi_version(void);
/* Human-readable build string ("sm_100a, nvcc 12.9, ..."). */
const char* dmvs_build_info(void);
/* Kernels launched by this library since it was loaded (host-side counter, for bench.py's gpu_launches). */
uint64_t dmvs_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Direct convolution, 2-D or 3-D (KD=1,D=1 for 2-D), stride 1 or 2 in every spatial dim, zero pad.
 * Replaces nn.Conv2d / nn.Conv3d (+ folded eval BatchNorm, + activation, + residual) everywhere
 * in the reference: module.py:42-58,88-102,282-301,332-336,364-396,427-439,455-456,481-485;
 * update.py:44-48,94,186,222,237-243,281-287,335-339; SepConvGRU module.py:156-177.
 *
 * Weights are pre-packed by the host as [KD][KH][KW][cin_pad][cout_pad] (cin_pad = (C1+C2) rounded
 * up to 4, cout_pad = Cout rounded up to 4, zero filled), BN scale folded in, BN shift in `bias`.
 * ------------------------------------------------------------------------------------------- */
typedef struct dmvs_conv_desc {
  /* input (virtual concat of x[.., C1] and x2[.., C2] along channels; x2 may be NULL) */
  const float* x;
  const float* x2;
  int32_t N, D, H, W;       /* logical input size (after the optional nearest x2 upsample) */
  int32_t C1, C2;
  int32_t x_ps, x2_ps;      /* pixel strides of x / x2 in floats */
  int32_t in_up2;           /* 1: x is stored at (H/2, W/2) and nearest-upsampled on load (update.py:38-42) */
  /* optional GroupNorm(4 groups)+affine+SiLU applied to x on load (update.py:117-133):
   * v = silu((v - mean_g) * rstd_g * in_g1[c] + in_g0[c]); stats = [N][4][2] (sum, sumsq) as 64-bit fixed-point
   * integers in units of 2^-20, the accumulators an earlier call filled through `out_stats` */
  const int64_t* in_stats;
  const float* in_g1;
  const float* in_g0;
  float in_inv_count;       /* 1 / (elements per (sample, group)) */
  /* weights */
  const float* w;
  const float* w_t;         /* tensor-core layout [KD][KH][KW][cout_pad8][cin_pad8] (pads to 8, zero filled); may be
                               NULL when precision == DMVS_PREC_FP32 */
  const float* w_tc;        /* tcgen05 layout: two planes (hi = rna_tf32(w), lo = rna_tf32(w - hi)), each
                               [KD][KH*KW][cin_pad8/4][cout_pad16][4]; may be NULL unless precision is DMVS_PREC_TC_* */
  const float* w_ws;        /* width-stacked tcgen05 layout (conv_ws.cu), packed for THIS stride and padding: a stride-S
                               convolution runs as S*S stride-1 phases; with KHe, KWe the kernel extent in phase-plane
                               shifts (KH, KW for S = 1), per output-channel chunk (CC <= min(64, (256/KWe) & ~7)
                               channels, N = KWe*CC rounded up to 16) two planes (hi, lo), each
                               [KD][S*S][cin_pad8/8][KHe][2 channel quads][N][4] with column kw'*CC + c, zero where a
                               phase has no tap (packing.pack_ws); may be NULL */
  const float* w_ws_pair;   /* optional, layers with <= 4 input channels and KH >= 2 (stride 1): the width-stacked layout with
                               kernel rows PAIRED along the MMA's K dimension, per output-channel chunk two planes (hi, lo),
                               each [KD][ceil(KH/2)][2][N][4]: quad q of pair j holds kernel row 2j+q (zero past KH-1) of
                               input channels 0..3 (packing.pack_ws_pair).  The TMA-fed back end then issues ceil(KH/2)
                               row-MMAs per stage instead of KH and stages one channel quad instead of two. */
  const float* w_ws16;      /* optional: the `w_ws` layout with the lo plane of every slab replaced by fp16 correction
                               operands for DMVS_PREC_WS2_TF32_F16C - per (kernel row, column n) two 16-byte units of 8
                               halves: unit 0 = fp16(hi / 16) of input channels 0..7 of the chunk, unit 1 = fp16(16 (w - hi))
                               (packing.pack_ws(..., corr16=True)); packed for THIS stride and padding like `w_ws` */
  int32_t precision;        /* DMVS_PREC_* */
  const float* bias;        /* [Cout] or NULL */
  int32_t KD, KH, KW, stride, pad_d, pad_h, pad_w;
  /* output */
  float* y;
  int32_t Do, Ho, Wo, Cout;
  int32_t y_ps;
  /* epilogue */
  int32_t act, act_c0;      /* activation applied to channels >= act_c0 */
  int32_t res_mode;
  const float* res;
  int32_t res_ps, res_up2;  /* res_up2: residual stored at half resolution, nearest-upsampled (module.py:409-416) */
  int32_t epi;
  const float* aux1;
  const float* aux2;
  int32_t aux1_ps, aux2_ps, gru_hidden;
  /* Phase launches of "3x3 convolution of a nearest x2 upsampled map" (ops.conv_up2): the output rows of one launch are
   * every other row of the real tensor and the padding is one-sided.  Only the TMA-fed width-stacked back end accepts
   * these (dmvs_conv_backends reports bit 4 alone); 0 everywhere else. */
  int32_t explicit_extent;  /* 1: (Do, Ho, Wo) are taken as given - output o reads inputs o*stride - pad + k, zero outside
                               the input - instead of being checked against the symmetric-padding formula */
  int32_t y_row_stride;     /* floats between consecutive output rows (image n / slice d start at (n*Do + d)*Ho rows);
                               0 = dense (Wo * y_ps) */
  int32_t res_row_stride;   /* same for `res` (not with res_up2) */
  int64_t* out_stats;       /* optional [N][4][2] sum / sumsq of the written values per GroupNorm group, accumulated as
                               64-bit fixed point (2^-20 units; integer adds: the result is independent of the order in
                               which CTAs arrive, runs are bit-reproducible); must be zeroed by the caller */
} dmvs_conv_desc;

int dmvs_conv_f32(const dmvs_conv_desc* desc, void* stream);

/* Which arithmetic back ends can run `desc` (bit 0: FFMA, bit 1: mma.sync, bit 2: tcgen05 with taps as descriptor
 * offsets, bit 3: tcgen05 width-stacked, bit 4: tcgen05 width-stacked behind the TMA pipeline).  Used by the host-side per-layer autotuner (the counterpart of the
 * reference's `cudnn.benchmark = True`, test.py:18).  Launches nothing. */
int dmvs_conv_backends(const dmvs_conv_desc* desc);

/* Tile plan the width-stacked tcgen05 back end would use for `desc` (host only, launches nothing): per kernel launch
 * eight ints {CC, N, TH, TW, M blocks, ring depth, CTAs per SM, shared-memory bytes} are written to out (capacity
 * `cap` launches).  Returns the number of launches or DMVS_ERR_*.  Pointers in `desc` are only checked for alignment. */
int dmvs_conv_ws_plan(const dmvs_conv_desc* desc, int32_t* out, int32_t cap);
/* Same for the TMA-fed width-stacked back end (CTAs per SM is always 1). */
int dmvs_conv_ws2_plan(const dmvs_conv_desc* desc, int32_t* out, int32_t cap);
/* Tuning aid: with DMVS_WS2_DBG=1 in the environment the TMA-fed kernel stamps the SM clock of CTA 0's first 64 pipeline
 * events per role (8 roles x 64 slots: stage issued / landed / split / operands ready / MMAs issued / accumulators ready /
 * tile stored / accumulator set free); this call synchronises and copies the stamps of the last launch.  Returns the
 * number of values written, 0 when the facility is off. */
int dmvs_conv_ws2_timeline(int64_t* out, int32_t count);

/* ConvTranspose3d(k=3, s=2, p=1, output_padding=1) + folded BN + ReLU + skip add
 * (module.Deconv3d as used by CostRegNet_small, module.py:110-144,436-437,445-446).
 * x [N][D][H][W][Cin] -> y [N][2D][2H][2W][Cout]; w packed [27][Cin][Cout]; skip has y's shape. */
int dmvs_deconv3d_f32(const float* x, const float* w, const float* bias, const float* skip, float* y,
                      int32_t N, int32_t D, int32_t H, int32_t W, int32_t Cin, int32_t Cout, void* stream);

/* Conv3d(8 -> 1, k=3, s=1, p=1) marching along depth (conv3d_to1.cu): the output layer of PixelViewWeight
 * (module.py:454-457) and CostRegNet_small.prob (module.py:439,447).  x [N][D][H][W][x_ps >= 8] (16-byte aligned
 * pixels); w_host = the 216 weights [kd][kh][kw][ci] in HOST memory (they travel as launch parameters and become
 * constant-bank operands; read before the call returns).
 * mode 0: y [N][D][H][W] = conv + bias.
 * mode 1: y [N][H][W] = max over depth of sigmoid(conv + bias) - PixelViewWeight.forward's tail (module.py:459-463)
 *         fused, the logits never reach memory. */
int dmvs_conv3d_to1_f32(const float* x, int32_t x_ps, const float* w_host, float bias, float* y, int32_t N, int32_t D,
                        int32_t H, int32_t W, int32_t mode, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Homography warp + group-wise correlation (module.py:181-218, 514-548, 575-667).
 * ------------------------------------------------------------------------------------------- */

/* proj [B][V][2][4][4] (extrinsic, intrinsic) -> hom [B][V-1][12] = rows of [R|t] of
 * P_src @ inverse(P_ref), P = [K E[:3,:4]; 0 0 0 1]   (module.py:188-190, 520-525).  fp64 inside. */
int dmvs_compose_homographies(const float* proj, float* hom, int32_t B, int32_t V, void* stream);

/* differentiable_warping (module.py:181-218): src [B][Hs][Ws][C] (pixel stride src_ps), hom [B][12],
 * depth [B][D][H][W] -> out [B][D][H][W][C] (channels-last volume).  Operator-surface entry point; the
 * fused kernels below never materialise this volume. */
int dmvs_warp_volume(const float* src, int32_t src_ps, const float* hom, const float* depth, float* out,
                     int32_t B, int32_t C, int32_t Hs, int32_t Ws, int32_t D, int32_t H, int32_t W, void* stream);

/* Stage-1 sweep (module.py:518-531): for every source view v and plane d, cor[b][v][d][y][x][g] =
 * mean_k ref[g*C/G+k] * warp(src_v)[g*C/G+k].  feats [V][B][H][W][C] (view 0 = reference, pixel
 * stride C), hom [B][V-1][12], plane_depth [B][D] metric depth per plane.  cor [B*(V-1)][D][H][W][G]. */
int dmvs_plane_sweep_corr(const float* feats, const float* hom, const float* plane_depth, float* cor,
                          int32_t B, int32_t V, int32_t C, int32_t G, int32_t D, int32_t H, int32_t W, void* stream);

/* PixelViewWeight tail (module.py:460-463): logit [N][D][H][W] -> w [N][H][W] = max_d sigmoid(logit). */
int dmvs_view_weight_max(const float* logit, float* w, int32_t N, int32_t D, int32_t HW, void* stream);

/* Stage-1 aggregation (module.py:539-548): vol[b] = sum_v w_v*cor_v / (1e-8 + sum_v w_v).
 * cor [B][V1][D][H][W][G], w [B][V1][H][W] -> vol [B][D][H][W][G]. */
int dmvs_aggregate_views(const float* cor, const float* w, float* vol, int32_t B, int32_t V1, int32_t D,
                         int32_t HW, int32_t G, void* stream);

/* softmax over D + expected index + window-4 confidence (module.py:554-571).
 * logits [B][D][H][W] -> norm_inv [B][H][W] (= idx/(D-1)), depth [B][H][W], conf [B][H][W],
 * floor_idx [B][H][W] (int32, optional).  depth_min/depth_max [B] = 1/depth_values[:,-1], 1/depth_values[:,0]
 * (diffusion.py:140-143). */
int dmvs_depth_regression(const float* logits, const float* depth_min, const float* depth_max, float* norm_inv,
                          float* depth, float* conf, int32_t* floor_idx, int32_t B, int32_t D, int32_t HW,
                          void* stream);

/* GetCost (module.py:250-277, 583-667) fused: hypothesis sampler + V-1 warps + group correlation +
 * view-weighted mean.  feats [V][B][H][W][C]; inv_depth [B][H][W]; conf [B][H][W] or NULL;
 * view_w [B][V-1][H>>wshift][W>>wshift] (stage-1 weights, nearest-upsampled on the fly,
 * diffusion.py:219-221); cost [B][H][W][G*D] channel g*D+d (pixel stride cost_ps);
 * samples [B][H][W][D] (pixel stride samp_ps). */
int dmvs_get_cost(const float* feats, const float* hom, const float* inv_depth, const float* conf, int32_t conf_ps,
                  const float* view_w, const float* depth_min, const float* depth_max, float* cost, int32_t cost_ps,
                  float* samples, int32_t samp_ps, int32_t B, int32_t V, int32_t C, int32_t G, int32_t D, int32_t H,
                  int32_t W, int32_t wshift, float interval, float min_radius, float max_radius, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Point-wise pieces
 * ------------------------------------------------------------------------------------------- */

/* Block tail of ResnetBlock (update.py:117-159): y = silu(GN(x)*g1+g0) + res.
 * x [N][HW][C] raw conv output, stats [N][4][2] fixed-point accumulators (see dmvs_conv_desc.out_stats); res (pixel
 * stride res_ps) or NULL. */
int dmvs_groupnorm_silu_add(const float* x, const int64_t* stats, const float* g1, const float* g0, const float* res,
                            int32_t res_ps, float* y, int32_t y_ps, int32_t N, int32_t HW, int32_t C, void* stream);

/* upsample_depth (module.py:237-248) + disp_to_depth (module.py:220-227) + depth_to_disp (:229-235).
 * n [B][H][W], mask [B][H][W][9*r*r] -> raw_up [B][rH][rW] (the convex combination itself), depth_up
 * (= disp_to_depth(raw_up)), norm_up (= depth_to_disp(depth_up), the value the next stage starts from,
 * diffusion.py:215-217).  Each output is optional; depth_min/max are needed for the last two. */
int dmvs_upsample_depth(const float* n, const float* mask, int32_t mask_ps, const float* depth_min,
                        const float* depth_max, float* raw_up, float* depth_up, float* norm_up, int32_t B, int32_t H,
                        int32_t W, int32_t ratio, void* stream);

/* Refinement state update (update.py:479-483, 496-502).
 * mode 0 (start of a DDIM step): inv = clamp(inv0 + img, 0, 1); delta = inv - inv0, img = scale*noise
 *        when noise != NULL else img read from `delta` (in place).
 * mode 1 (after an iteration):   delta += upd; inv = clamp(inv0 + delta, 0, 1); delta = inv - inv0.
 * inv is written twice: dense [B][HW] and into channel `inv_slot` of a [B][HW][slot_ps] buffer (the
 * U-Net input, update.py:297,493).  depth (optional) = disp_to_depth(inv). */
int dmvs_refine_update(int32_t mode, const float* inv0, const float* noise_or_upd, int32_t upd_ps, float scale,
                       float* delta, float* inv, float* inv_slot, int32_t slot_ps, const float* depth_min,
                       const float* depth_max, float* depth, int32_t B, int32_t HW, void* stream);

/* DDIM step (update.py:504-519): img = delta*sqrt(a_next) + c*pred_noise + sigma*(scale*noise) with
 * pred_noise = (k_recip*img - delta)/k_recipm1.  All scalars precomputed on the host. */
int dmvs_ddim_step(float* img, const float* delta, const float* noise, float k_recip, float k_recipm1,
                   float sqrt_a_next, float c, float sigma, float scale, int64_t count, void* stream);

/* nearest-neighbour upsampling of a [B][H][W] map (element stride x_ps) by an integer factor
 * (diffusion.py:205-207,274-278). */
int dmvs_upsample_nearest(const float* x, int32_t x_ps, float* y, int32_t B, int32_t H, int32_t W, int32_t factor,
                          void* stream);

/* y [N][H][W][y_ps >= C] += table[ry][rx][C] on the one-pixel frame of every image, (ry, rx) in {0 first, 1 interior,
 * 2 last} row / column (the interior is untouched): the position-dependent remainder of a per-channel bias that passed
 * through a zero-padded 3x3 convolution.  Used where FeatureNet's `inner2` bias is folded through `out3`
 * (module.py:395-396,415-417; pipeline.FeatureNetPlan).  C % 4 == 0, H, W >= 2. */
int dmvs_border_bias_add(float* y, int32_t y_ps, const float* table, int32_t N, int32_t H, int32_t W, int32_t C,
                         void* stream);

/* Input images (datasets/mvs.py:93-97: [N][3][H][W], RGB in [0,1]) -> channels-last [N][HW][4] with a zero
 * fourth channel, the layout the first convolutions (module.py:332,364) stage with 128-bit copies. */
int dmvs_image_to_nhwc4(const float* x, float* y, int32_t N, int32_t HW, void* stream);

/* The same staging for 8-bit images as decoded from disk: y = float(x) / 255 (one IEEE fp32 division, bit-identical to
 * the loader's `np.array(img, dtype=np.float32) / 255.`, datasets/data_io.py:166-170).  x is planar [N][3][HW] (c_stride =
 * HW, p_stride = 1) or interleaved [N][HW][3] (c_stride = 1, p_stride = 3); n_stride = bytes between images. */
int dmvs_image_u8_to_nhwc4(const uint8_t* x, int64_t n_stride, int64_t c_stride, int32_t p_stride, float* y, int32_t N,
                           int32_t HW, void* stream);

/* layout transposes between the reference's NCHW operator surface and channels-last */
int dmvs_nchw_to_nhwc(const float* x, float* y, int32_t y_ps, int32_t N, int32_t C, int32_t HW, void* stream);
int dmvs_nhwc_to_nchw(const float* x, int32_t x_ps, float* y, int32_t N, int32_t C, int32_t HW, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Depth-map filtering / fusion (filter.py:8-87,189-215).  Host-side matrix algebra (float32 inverses and
 * products, as numpy computes them in the reference) is passed in as doubles.
 * ------------------------------------------------------------------------------------------- */

/* check_geometric_consistency (filter.py:54-87) for one (reference, source) pair of depth maps [H][W] / [Hs][Ws]:
 * mats68 = inv(K_ref)[9], E_src@inv(E_ref)[16], K_src[9], inv(K_src)[9], E_ref@inv(E_src)[16], K_ref[9] (row major).
 * Writes mask (0/1), the reprojected depth (0 where inconsistent) and optionally the source pixel coordinates; when
 * sum_reproj / count are given they are updated in place (+= depth_reproj, += mask), which accumulates
 * `sum(all_srcview_depth_ests)` and `geo_mask_sum` of filter_depth over successive source views. */
int dmvs_geo_consistency(const float* depth_ref, const float* depth_src, const double* mats68, float depth_min,
                         float depth_max, double pix_thres, float depth_thres, uint8_t* mask, float* depth_reproj,
                         float* x_src, float* y_src, float* sum_reproj, int32_t* count, int32_t H, int32_t W,
                         int32_t Hs, int32_t Ws, void* stream);

/* Averaged depth (float64, as numpy's float32 / int32 quotient), geometric and final masks and the world-space point
 * of every pixel (filter.py:189-212).  mats25 = inv(K_ref)[9], inv(E_ref)[16]; photo_mask may be NULL. */
int dmvs_fuse_points(const float* depth_ref, const float* sum_reproj, const int32_t* count, const uint8_t* photo_mask,
                     int32_t geo_thres, const double* mats25, double* depth_avg, uint8_t* geo_mask, uint8_t* final_mask,
                     float* xyz, int32_t H, int32_t W, void* stream);

/* One reference view of filter_depth (filter.py:117-212) or, when dyn_dist > 0, of filter_depth_dynamic (:230-262,
 * 311-412) in ONE launch: photometric mask (conf[c] > photo_thres[c], float32), geometric consistency against all S
 * source views (S <= 16), averaged depth, geometric / final masks and the world point of every pixel.
 * depth_src: host array of S device pointers ([Hs][Ws] maps); mats_dev: DEVICE array [S][68] doubles, one mats68 block
 * (see dmvs_geo_consistency) per source view; conf / photo_thres: host arrays of n_conf (<= 3) device pointers / floats.
 * Static mode: pix_thres (float64 compare), depth_thres (float32), depth_min/max (float32 range test on depth_ref),
 * geo_thres.  Dynamic mode: tests i/dyn_dist (float64) and i/dyn_rel (rounded to float32) for i = dyn_view_num..10, a
 * pixel passes when >= i source views pass test i for some i; the final mask also requires avg_min <= averaged depth
 * <= avg_max (float64).  Outputs as dmvs_fuse_points plus the photometric mask. */
int dmvs_fuse_view(const float* depth_ref, const float* const* depth_src, const double* mats_dev, int32_t S, int32_t H,
                   int32_t W, int32_t Hs, int32_t Ws, const float* const* conf, const float* photo_thres, int32_t n_conf,
                   const double* mats25, float depth_min, float depth_max, double pix_thres, float depth_thres,
                   int32_t geo_thres, int32_t dyn_view_num, double dyn_dist, double dyn_rel, double avg_min,
                   double avg_max, uint8_t* photo_mask, uint8_t* geo_mask, uint8_t* final_mask, double* depth_avg,
                   float* xyz, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFMVS_B200_H */
