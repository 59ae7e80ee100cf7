// Point-wise / small-stencil kernels of the refinement loop and the layout transposes.
// All are pure streaming kernels (HBM/L2-bound): 128-bit accesses where the layout allows.
#include "common.cuh"

namespace dmvs {
namespace {

// y = silu(GN(x) * g1 + g0) + res            (update.py:117-159, Block + ResnetBlock tail)
__global__ void __launch_bounds__(256) groupnorm_silu_add_kernel(const float* __restrict__ x,
                                                                 const long long* __restrict__ stats,
                                                                 const float* __restrict__ g1,
                                                                 const float* __restrict__ g0,
                                                                 const float* __restrict__ res, int res_ps,
                                                                 float* __restrict__ y, int y_ps, int N, int HW, int C,
                                                                 float inv_count) {
  pdl_sync();
  extern __shared__ float ab[];  // [2][C] for sample n = blockIdx.y
  const int n = blockIdx.y;
  const int cpg = C / 4;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const double s = stat_value(stats[(n * 4 + g) * 2 + 0]), q = stat_value(stats[(n * 4 + g) * 2 + 1]);
    const double mean = s * (double)inv_count;
    double var = q * (double)inv_count - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    const float rstd = (float)(1.0 / sqrt(var + 1e-5));
    const float a = g1[c] * rstd;
    ab[c] = a;
    ab[C + c] = g0[c] - (float)mean * a;
  }
  __syncthreads();
  const int c4n = C / 4;
  const int64_t total = (int64_t)HW * c4n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const int64_t p = (int64_t)n * HW + i / c4n;
    const float4 v = ldg4(x + p * C + c);
    float4 r;
    r.x = siluf_(fmaf(v.x, ab[c + 0], ab[C + c + 0]));
    r.y = siluf_(fmaf(v.y, ab[c + 1], ab[C + c + 1]));
    r.z = siluf_(fmaf(v.z, ab[c + 2], ab[C + c + 2]));
    r.w = siluf_(fmaf(v.w, ab[c + 3], ab[C + c + 3]));
    if (res != nullptr) {
      const float4 s = ldg4(res + p * res_ps + c);
      r.x += s.x; r.y += s.y; r.z += s.z; r.w += s.w;
    }
    *reinterpret_cast<float4*>(y + p * y_ps + c) = r;
  }
}

// convex upsampling (module.py:237-248): one thread per output pixel
__global__ void __launch_bounds__(256) upsample_depth_kernel(const float* __restrict__ nrm,
                                                             const float* __restrict__ mask, int mask_ps,
                                                             const float* __restrict__ depth_min,
                                                             const float* __restrict__ depth_max,
                                                             float* __restrict__ raw_up, float* __restrict__ depth_up,
                                                             float* __restrict__ norm_up, int B, int H, int W, int r) {
  pdl_sync();
  const int Ho = H * r, Wo = W * r;
  const int64_t total = (int64_t)B * Ho * Wo;
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= total) return;
  const int ox = (int)(o % Wo);
  const int64_t t = o / Wo;
  const int oy = (int)(t % Ho);
  const int b = (int)(t / Ho);
  const int X = ox / r, j = ox - X * r;
  const int Y = oy / r, i = oy - Y * r;
  const float* mp = mask + (((int64_t)b * H + Y) * W + X) * mask_ps + i * r + j;
  const int rr = r * r;
  float m[9];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    m[k] = __ldg(mp + k * rr);
    mx = fmaxf(mx, m[k]);
  }
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    m[k] = expf(m[k] - mx);
    sum += m[k];
  }
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int yy = Y + k / 3 - 1, xx = X + k % 3 - 1;
    const float nv = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(nrm + ((int64_t)b * H + yy) * W + xx) : 0.f;
    acc = __fadd_rn(acc, __fmul_rn(__fdiv_rn(m[k], sum), nv));
  }
  if (raw_up) raw_up[o] = acc;
  if (depth_up || norm_up) {
    const DepthRange rng(__ldg(depth_min + b), __ldg(depth_max + b));
    const float dep = rng.to_depth(acc);
    if (depth_up) depth_up[o] = dep;
    if (norm_up) norm_up[o] = rng.to_norm(dep);
  }
}

__global__ void __launch_bounds__(256) refine_update_kernel(int mode, const float* __restrict__ inv0,
                                                            const float* __restrict__ src, int src_ps, float scale,
                                                            float* __restrict__ delta, float* __restrict__ inv,
                                                            float* __restrict__ inv_slot, int slot_ps,
                                                            const float* __restrict__ depth_min,
                                                            const float* __restrict__ depth_max,
                                                            float* __restrict__ depth, int B, int HW) {
  pdl_sync();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * HW) return;
  const float base = __ldg(inv0 + i);
  float dl;
  if (mode == 0) {
    dl = src != nullptr ? __fmul_rn(scale, __ldg(src + i)) : delta[i];
  } else {
    dl = __fadd_rn(delta[i], __ldg(src + i * src_ps));
  }
  const float v = fminf(fmaxf(__fadd_rn(base, dl), 0.0f), 1.0f);
  delta[i] = __fsub_rn(v, base);
  inv[i] = v;
  if (inv_slot) inv_slot[i * slot_ps] = v;
  if (depth) {
    const int b = (int)(i / HW);
    const DepthRange rng(__ldg(depth_min + b), __ldg(depth_max + b));
    depth[i] = rng.to_depth(v);
  }
}

__global__ void ddim_step_kernel(float* __restrict__ img, const float* __restrict__ delta,
                                 const float* __restrict__ noise, float k_recip, float k_recipm1, float sqrt_a_next,
                                 float c, float sigma, float scale, int64_t count) {
  pdl_sync();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float dl = __ldg(delta + i);
  const float pred = __fdiv_rn(__fsub_rn(__fmul_rn(k_recip, img[i]), dl), k_recipm1);
  float v = __fadd_rn(__fmul_rn(dl, sqrt_a_next), __fmul_rn(c, pred));
  v = __fadd_rn(v, __fmul_rn(sigma, __fmul_rn(scale, __ldg(noise + i))));
  img[i] = v;
}

__global__ void upsample_nearest_kernel(const float* __restrict__ x, int x_ps, float* __restrict__ y, int B, int H, int W,
                                        int f) {
  pdl_sync();
  const int Ho = H * f, Wo = W * f;
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= (int64_t)B * Ho * Wo) return;
  const int ox = (int)(o % Wo);
  const int64_t t = o / Wo;
  const int oy = (int)(t % Ho);
  const int b = (int)(t / Ho);
  y[o] = __ldg(x + (((int64_t)b * H + oy / f) * W + ox / f) * x_ps);
}

// y[n][oy][ox][:] += table[ry][rx][:] on the one-pixel frame of every image (ry, rx = 0 first / 1 interior / 2 last row
// or column): the position-dependent part of a bias that went through a zero-padded 3x3 convolution (see
// pipeline.FeatureNetPlan: inner2's bias seen through out3).  One thread per (frame pixel, channel quad).
__global__ void border_bias_add_kernel(float* __restrict__ y, int y_ps, const float* __restrict__ table, int N, int H, int W,
                                       int C4) {
  pdl_sync();
  const int frame = 2 * W + 2 * (H - 2);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * frame * C4) return;
  const int c4 = (int)(i % C4);
  int64_t t = i / C4;
  const int f = (int)(t % frame);
  const int n = (int)(t / frame);
  int oy, ox;
  if (f < W) { oy = 0; ox = f; }
  else if (f < 2 * W) { oy = H - 1; ox = f - W; }
  else { const int g = f - 2 * W; oy = 1 + (g >> 1); ox = (g & 1) ? W - 1 : 0; }
  const int ry = oy == 0 ? 0 : (oy == H - 1 ? 2 : 1), rx = ox == 0 ? 0 : (ox == W - 1 ? 2 : 1);
  const float4 b = ldg4(table + ((ry * 3 + rx) * C4 + c4) * 4);
  float4* p = reinterpret_cast<float4*>(y + (((int64_t)n * H + oy) * W + ox) * y_ps + c4 * 4);
  float4 v = *p;
  v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  *p = v;
}

// [N][C][HW] -> [N][HW][C] through a 32x32 shared tile (both sides coalesced)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int y_ps, int C, int HW) {
  pdl_sync();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < C && p < HW) ? __ldg(x + ((int64_t)n * C + c) * HW + p) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    if (p < HW && c < C) y[((int64_t)n * HW + p) * y_ps + c] = tile[threadIdx.x][r];
  }
}

// C <= 4 (images): one thread per pixel, C coalesced plane reads, one packed write
__global__ void nchw_to_nhwc_small_kernel(const float* __restrict__ x, float* __restrict__ y, int y_ps, int C, int HW,
                                          int64_t total) {
  pdl_sync();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t n = i / HW;
  const int p = (int)(i - n * HW);
  const float* xp = x + n * C * HW + p;
  float* yp = y + i * y_ps;
  for (int c = 0; c < C; ++c) yp[c] = __ldg(xp + (int64_t)c * HW);
}

// RGB image planes [N][3][HW] -> [N][HW][4] with a zero fourth channel: one aligned 16-byte store per pixel, so
// the first convolution of FeatureNet / ContextNet can stage its input with 128-bit asynchronous copies.
__global__ void image_to_nhwc4_kernel(const float* __restrict__ x, float4* __restrict__ y, int HW, int64_t total) {
  pdl_sync();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t n = i / HW;
  const int p = (int)(i - n * HW);
  const float* xp = x + n * 3 * HW + p;
  y[i] = make_float4(__ldg(xp), __ldg(xp + HW), __ldg(xp + 2 * (int64_t)HW), 0.0f);
}

// Same for 8-bit images as decoded from disk (datasets/data_io.py:166-170 computes np.float32(u8) / 255. on the host:
// one IEEE fp32 division per value, reproduced here bit for bit), either planar [N][3][HW] (c_stride = HW, p_stride = 1)
// or interleaved [N][HW][3] (c_stride = 1, p_stride = 3): 4x fewer bytes over PCIe than fp32 images.
__global__ void image_u8_to_nhwc4_kernel(const uint8_t* __restrict__ x, int64_t n_stride, int64_t c_stride, int p_stride,
                                         float4* __restrict__ y, int HW, int64_t total) {
  pdl_sync();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t n = i / HW;
  const int p = (int)(i - n * HW);
  const uint8_t* xp = x + n * n_stride + (int64_t)p * p_stride;
  y[i] = make_float4(__fdiv_rn((float)__ldg(xp), 255.0f), __fdiv_rn((float)__ldg(xp + c_stride), 255.0f),
                     __fdiv_rn((float)__ldg(xp + 2 * c_stride), 255.0f), 0.0f);
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, int x_ps, float* __restrict__ y, int C, int HW) {
  pdl_sync();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (p < HW && c < C) ? __ldg(x + ((int64_t)n * HW + p) * x_ps + c) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    if (c < C && p < HW) y[((int64_t)n * C + c) * HW + p] = tile[threadIdx.x][r];
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace dmvs

using namespace dmvs;

namespace dmvs {
unsigned long long g_launch_count = 0;
}

extern "C" int dmvs_abi_version(void) { return DMVS_ABI_VERSION; }

extern "C" uint64_t dmvs_launch_count(void) { return g_launch_count; }

extern "C" const char* dmvs_build_info(void) {
  return "diffmvs_b200 kernels: sm_100a, CUDA " DMVS_STR(__CUDACC_VER_MAJOR__) "." DMVS_STR(__CUDACC_VER_MINOR__)
         ", TMA-fed tcgen05/TMEM width-stacked + FFMA2 convolutions, fused warp/correlation"
#ifdef DMVS_LEGACY_BACKENDS
         ", legacy back ends (mma.sync, tap-offset tcgen05)"
#endif
      ;
}

extern "C" int dmvs_groupnorm_silu_add(const float* x, const int64_t* stats, const float* g1, const float* g0,
                                       const float* res, int32_t res_ps, float* y, int32_t y_ps, int32_t N, int32_t HW,
                                       int32_t C, void* stream) {
  if (!x || !stats || !g1 || !g0 || !y || N <= 0 || HW <= 0 || C <= 0) return DMVS_ERR_ARG;
  if ((C % 16) != 0 && (C % 4) != 0) return DMVS_ERR_ARG;
  if ((C % 4) != 0 || (y_ps % 4) != 0 || (res && (res_ps % 4) != 0)) return DMVS_ERR_ALIGN;
  if (!aligned16(x) || !aligned16(y) || (res && !aligned16(res))) return DMVS_ERR_ALIGN;
  const int64_t total = (int64_t)HW * (C / 4);
  const int64_t want = ceil_div64(total, 256);
  const int bx = (int)(want < (int64_t)kNumSMs * 8 ? want : (int64_t)kNumSMs * 8);
  dim3 grid(bx, N);
  const float inv_count = 1.0f / ((float)HW * (float)(C / 4));
  launch_pdl(groupnorm_silu_add_kernel, dim3(grid), dim3(256), 2 * C * sizeof(float), static_cast<cudaStream_t>(stream), 
      x, reinterpret_cast<const long long*>(stats), g1, g0, res, res_ps, y, y_ps, N, HW, C, inv_count);
  return launch_status();
}

extern "C" int dmvs_upsample_depth(const float* n, const float* mask, int32_t mask_ps, const float* depth_min,
                                   const float* depth_max, float* raw_up, float* depth_up, float* norm_up, int32_t B,
                                   int32_t H, int32_t W, int32_t ratio, void* stream) {
  if (!n || !mask || B <= 0 || H <= 0 || W <= 0 || ratio <= 0) return DMVS_ERR_ARG;
  if (!raw_up && !depth_up && !norm_up) return DMVS_ERR_ARG;
  if ((depth_up || norm_up) && (!depth_min || !depth_max)) return DMVS_ERR_ARG;
  if (mask_ps < 9 * ratio * ratio) return DMVS_ERR_ARG;
  const int64_t total = (int64_t)B * H * W * ratio * ratio;
  launch_pdl(upsample_depth_kernel, dim3((unsigned)ceil_div64(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      n, mask, mask_ps, depth_min, depth_max, raw_up, depth_up, norm_up, B, H, W, ratio);
  return launch_status();
}

extern "C" int dmvs_refine_update(int32_t mode, const float* inv0, const float* noise_or_upd, int32_t upd_ps,
                                  float scale, float* delta, float* inv, float* inv_slot, int32_t slot_ps,
                                  const float* depth_min, const float* depth_max, float* depth, int32_t B, int32_t HW,
                                  void* stream) {
  if (!inv0 || !delta || !inv || B <= 0 || HW <= 0) return DMVS_ERR_ARG;
  if (mode != 0 && mode != 1) return DMVS_ERR_ARG;
  if (mode == 1 && (!noise_or_upd || upd_ps <= 0)) return DMVS_ERR_ARG;
  if (depth && (!depth_min || !depth_max)) return DMVS_ERR_ARG;
  const int64_t total = (int64_t)B * HW;
  launch_pdl(refine_update_kernel, dim3((unsigned)ceil_div64(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      mode, inv0, noise_or_upd, upd_ps, scale, delta, inv, inv_slot, slot_ps, depth_min, depth_max, depth, B, HW);
  return launch_status();
}

extern "C" int dmvs_ddim_step(float* img, const float* delta, const float* noise, float k_recip, float k_recipm1,
                              float sqrt_a_next, float c, float sigma, float scale, int64_t count, void* stream) {
  if (!img || !delta || !noise || count <= 0) return DMVS_ERR_ARG;
  launch_pdl(ddim_step_kernel, dim3((unsigned)ceil_div64(count, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      img, delta, noise, k_recip, k_recipm1, sqrt_a_next, c, sigma, scale, count);
  return launch_status();
}

extern "C" int dmvs_upsample_nearest(const float* x, int32_t x_ps, float* y, int32_t B, int32_t H, int32_t W,
                                     int32_t factor, void* stream) {
  if (!x || !y || B <= 0 || H <= 0 || W <= 0 || factor <= 0 || x_ps <= 0) return DMVS_ERR_ARG;
  const int64_t total = (int64_t)B * H * W * factor * factor;
  launch_pdl(upsample_nearest_kernel, dim3((unsigned)ceil_div64(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, x_ps, y, B, H,
                                                                                                         W, factor);
  return launch_status();
}

extern "C" int dmvs_border_bias_add(float* y, int32_t y_ps, const float* table, int32_t N, int32_t H, int32_t W, int32_t C,
                                    void* stream) {
  if (!y || !table || N <= 0 || H < 2 || W < 2 || C <= 0 || (C % 4) != 0 || y_ps < C) return DMVS_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(y) & 15u) != 0 || (reinterpret_cast<uintptr_t>(table) & 15u) != 0 || (y_ps % 4) != 0)
    return DMVS_ERR_ALIGN;
  const int64_t total = (int64_t)N * (2 * W + 2 * (H - 2)) * (C / 4);
  launch_pdl(border_bias_add_kernel, dim3((unsigned)ceil_div64(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
             y, y_ps, table, N, H, W, C / 4);
  return launch_status();
}

extern "C" int dmvs_nchw_to_nhwc(const float* x, float* y, int32_t y_ps, int32_t N, int32_t C, int32_t HW,
                                 void* stream) {
  if (!x || !y || N <= 0 || C <= 0 || HW <= 0 || y_ps < C) return DMVS_ERR_ARG;
  if (C <= 4) {
    const int64_t total = (int64_t)N * HW;
    launch_pdl(nchw_to_nhwc_small_kernel, dim3((unsigned)ceil_div64(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, y, y_ps, C,
                                                                                                             HW, total);
    return launch_status();
  }
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), N), block(32, 8);
  if (grid.z > 65535 || grid.y > 65535) return DMVS_ERR_UNSUPPORTED;
  launch_pdl(nchw_to_nhwc_kernel, dim3(grid), dim3(block), 0, static_cast<cudaStream_t>(stream), x, y, y_ps, C, HW);
  return launch_status();
}

extern "C" int dmvs_image_to_nhwc4(const float* x, float* y, int32_t N, int32_t HW, void* stream) {
  if (!x || !y || N <= 0 || HW <= 0) return DMVS_ERR_ARG;
  if (!aligned16(y)) return DMVS_ERR_ALIGN;
  const int64_t total = (int64_t)N * HW;
  launch_pdl(image_to_nhwc4_kernel, dim3((unsigned)ceil_div64(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      x, reinterpret_cast<float4*>(y), HW, total);
  return launch_status();
}

extern "C" int dmvs_image_u8_to_nhwc4(const uint8_t* x, int64_t n_stride, int64_t c_stride, int32_t p_stride, float* y, int32_t N,
                                      int32_t HW, void* stream) {
  if (!x || !y || N <= 0 || HW <= 0 || c_stride <= 0 || p_stride <= 0) return DMVS_ERR_ARG;
  if (!aligned16(y)) return DMVS_ERR_ALIGN;
  const int64_t total = (int64_t)N * HW;
  launch_pdl(image_u8_to_nhwc4_kernel, dim3((unsigned)ceil_div64(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      x, n_stride, c_stride, p_stride, reinterpret_cast<float4*>(y), HW, total);
  return launch_status();
}

extern "C" int dmvs_nhwc_to_nchw(const float* x, int32_t x_ps, float* y, int32_t N, int32_t C, int32_t HW,
                                 void* stream) {
  if (!x || !y || N <= 0 || C <= 0 || HW <= 0 || x_ps < C) return DMVS_ERR_ARG;
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), N), block(32, 8);
  if (grid.z > 65535 || grid.y > 65535) return DMVS_ERR_UNSUPPORTED;
  launch_pdl(nhwc_to_nchw_kernel, dim3(grid), dim3(block), 0, static_cast<cudaStream_t>(stream), x, x_ps, y, C, HW);
  return launch_status();
}
