// Depth-map filtering and fusion on the GPU: the per-pixel part of the reference's
// `filter.py` - `reproject_with_depth` (:8-52), `check_geometric_consistency` (:54-87) and the averaging /
// back-projection of `filter_depth` (:189-215).  The reference evaluates these with numpy in float64 (its integer
// pixel grids and float32 maps promote to double) around one `cv2.remap(..., INTER_LINEAR)` in float32; the kernels
// below follow the same operation order and the same dtypes, so masks agree and values differ only by the last ulps
// of BLAS / OpenCV internals.  The small matrix algebra (inverses, extrinsic products in float32) stays on the
// host, computed with numpy exactly as the reference does (diffmvs_b200/fusion.py), and is passed in as doubles.
#include "common.cuh"

namespace dmvs {
namespace {

struct GeoMats {
  double Kref_inv[9];   // inv(intrinsics_ref)                       float32 inverse, widened
  double T_rs[16];      // extrinsics_src @ inv(extrinsics_ref)      float32 product, widened
  double Ksrc[9];
  double Ksrc_inv[9];
  double T_sr[16];      // extrinsics_ref @ inv(extrinsics_src)
  double Kref[9];
};

__device__ __forceinline__ void mat3(const double* M, double x, double y, double z, double& ox, double& oy, double& oz) {
  ox = M[0] * x + M[1] * y + M[2] * z;
  oy = M[3] * x + M[4] * y + M[5] * z;
  oz = M[6] * x + M[7] * y + M[8] * z;
}
__device__ __forceinline__ void mat4_rows3(const double* M, double x, double y, double z, double& ox, double& oy, double& oz) {
  ox = M[0] * x + M[1] * y + M[2] * z + M[3];
  oy = M[4] * x + M[5] * y + M[6] * z + M[7];
  oz = M[8] * x + M[9] * y + M[10] * z + M[11];
}

// cv2.remap(src, mapx, mapy, INTER_LINEAR) for a float32 image, BORDER_CONSTANT 0: the sampling position is rounded to
// 1/32 pixel (cvRound = nearest even), the integer part saturates to int16, the four weights are products of the
// float tables (1 - f/32, f/32), and the taps are accumulated left to right in float32.
__device__ __forceinline__ float remap_linear(const float* __restrict__ src, int Hs, int Ws, float mx, float my) {
  const float fx32 = __fmul_rn(mx, 32.0f), fy32 = __fmul_rn(my, 32.0f);
  // cvRound: round to nearest even; out-of-range / NaN behave like INT_MIN (-> far outside)
  int sx = (fx32 >= -2147483648.0f && fx32 < 2147483648.0f) ? __float2int_rn(fx32) : (int)0x80000000;
  int sy = (fy32 >= -2147483648.0f && fy32 < 2147483648.0f) ? __float2int_rn(fy32) : (int)0x80000000;
  int ix = sx >> 5, iy = sy >> 5;
  ix = ix < -32768 ? -32768 : (ix > 32767 ? 32767 : ix);
  iy = iy < -32768 ? -32768 : (iy > 32767 ? 32767 : iy);
  const float ax = (float)(sx & 31) * (1.0f / 32.0f), ay = (float)(sy & 31) * (1.0f / 32.0f);
  const float wx0 = 1.0f - ax, wy0 = 1.0f - ay;
  const float w0 = __fmul_rn(wy0, wx0), w1 = __fmul_rn(wy0, ax), w2 = __fmul_rn(ay, wx0), w3 = __fmul_rn(ay, ax);
  auto tap = [&](int y, int x) -> float {
    return (x >= 0 && x < Ws && y >= 0 && y < Hs) ? __ldg(src + (int64_t)y * Ws + x) : 0.0f;
  };
  const float a = tap(iy, ix), b = tap(iy, ix + 1), c = tap(iy + 1, ix), d = tap(iy + 1, ix + 1);
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, w0), __fmul_rn(b, w1)), __fmul_rn(c, w2)), __fmul_rn(d, w3));
}

// One (reference pixel, source view) pair of check_geometric_consistency (filter.py:8-87): reprojected depth, the
// float64 pixel distance and the float32 relative depth difference the thresholds are applied to, and the source
// pixel coordinates.  Shared by the per-pair kernel and the fused per-view kernel.
struct Reproj {
  float drep;      // depth_reproj (before masking)
  double dist;     // sqrt((x2d_reproj - x)^2 + (y2d_reproj - y)^2), float64 as numpy promotes float32 - int64
  float rel;       // |depth_reproj - depth_ref| / depth_ref, float32
  float xs, ys;    // x2d_src, y2d_src
};

__device__ __forceinline__ Reproj reproject_pixel(const GeoMats& m, const float* __restrict__ depth_src, int Hs, int Ws, int x, int y,
                                                  float dref_f) {
  const double dref = (double)dref_f;
  // reference pixel -> reference camera -> source camera -> source pixel            (filter.py:19-31)
  double rx, ry, rz;
  mat3(m.Kref_inv, (double)x * dref, (double)y * dref, dref, rx, ry, rz);
  double sx, sy, sz;
  mat4_rows3(m.T_rs, rx, ry, rz, sx, sy, sz);
  double kx, ky, kz;
  mat3(m.Ksrc, sx, sy, sz, kx, ky, kz);
  const double u = kx / kz, v = ky / kz;
  Reproj r;
  r.xs = (float)u;
  r.ys = (float)v;
  const float sampled = remap_linear(depth_src, Hs, Ws, r.xs, r.ys);                 // :32-33
  // source pixel with the sampled depth -> back to the reference                     (:35-47)
  const double sd = (double)sampled;
  double qx, qy, qz;
  mat3(m.Ksrc_inv, u * sd, v * sd, sd, qx, qy, qz);
  double px, py, pz;
  mat4_rows3(m.T_sr, qx, qy, qz, px, py, pz);
  r.drep = (float)pz;
  double ex, ey, ez;
  mat3(m.Kref, px, py, pz, ex, ey, ez);
  if (ex == 0.0) ex = 1e-5;
  if (ey == 0.0) ey = 1e-5;
  if (ez == 0.0) ez = 1e-5;
  double xr = ex / ez, yr = ey / ez;
  xr = fmin(fmax(xr, -1e8), 1e8);   // np.clip keeps NaN; fmin/fmax would drop it - NaN fails the threshold either way
  yr = fmin(fmax(yr, -1e8), 1e8);
  const float xr_f = (float)xr, yr_f = (float)yr;
  // consistency measures                                                             (:75-79)
  const double ddx = (double)xr_f - (double)x, ddy = (double)yr_f - (double)y;
  r.dist = sqrt(ddx * ddx + ddy * ddy);
  const float depth_diff = fabsf(__fsub_rn(r.drep, dref_f));
  r.rel = __fdiv_rn(depth_diff, dref_f);
  return r;
}

__global__ void geo_consistency_kernel(const float* __restrict__ depth_ref, const float* __restrict__ depth_src,
                                       const GeoMats m, float dmin, float dmax, double pix_thres, float depth_thres,
                                       uint8_t* __restrict__ mask, float* __restrict__ depth_reproj,
                                       float* __restrict__ x_src_out, float* __restrict__ y_src_out,
                                       float* __restrict__ sum_reproj, int32_t* __restrict__ count, int H, int W, int Hs,
                                       int Ws) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)H * W) return;
  const int x = (int)(i % W), y = (int)(i / W);
  const float dref_f = __ldg(depth_ref + i);
  const Reproj r = reproject_pixel(m, depth_src, Hs, Ws, x, y, dref_f);
  float drep = r.drep;
  const bool ok = r.dist < pix_thres && r.rel < depth_thres && dref_f > dmin && dref_f < dmax;   // :80-86
  if (!ok) drep = 0.0f;
  mask[i] = ok ? 1 : 0;
  depth_reproj[i] = drep;
  if (x_src_out) x_src_out[i] = r.xs;
  if (y_src_out) y_src_out[i] = r.ys;
  if (sum_reproj) sum_reproj[i] = __fadd_rn(sum_reproj[i], drep);   // sum(all_srcview_depth_ests), in view order
  if (count) count[i] += ok ? 1 : 0;
}

struct FuseMats {
  double Kref_inv[9];
  double Eref_inv[16];
};

__global__ void fuse_kernel(const float* __restrict__ depth_ref, const float* __restrict__ sum_reproj,
                            const int32_t* __restrict__ count, const uint8_t* __restrict__ photo_mask, int geo_thres,
                            const FuseMats m, double* __restrict__ depth_avg, uint8_t* __restrict__ geo_mask,
                            uint8_t* __restrict__ final_mask, float* __restrict__ xyz, int H, int W) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)H * W) return;
  const int x = (int)(i % W), y = (int)(i / W);
  const int n = count[i];
  // (sum(depth_reproj) + ref_depth) / (geo_mask_sum + 1): float32 sum, float64 quotient      (filter.py:189)
  const double avg = (double)__fadd_rn(sum_reproj[i], __ldg(depth_ref + i)) / (double)(n + 1);
  const bool geo = n >= geo_thres;
  const bool fin = geo && (photo_mask == nullptr || photo_mask[i] != 0);
  depth_avg[i] = avg;
  geo_mask[i] = geo ? 1 : 0;
  final_mask[i] = fin ? 1 : 0;
  // back-projection to world coordinates                                                     (:208-212)
  double cx, cy, cz;
  mat3(m.Kref_inv, (double)x * avg, (double)y * avg, avg, cx, cy, cz);
  double wx, wy, wz;
  mat4_rows3(m.Eref_inv, cx, cy, cz, wx, wy, wz);
  xyz[i * 3 + 0] = (float)wx;
  xyz[i * 3 + 1] = (float)wy;
  xyz[i * 3 + 2] = (float)wz;
}

// ------------------------------------------------------------------------------------------------------------------
// One launch per reference view: photometric mask, geometric consistency against EVERY source view (static thresholds,
// filter.py:117-191, or the nine dynamic threshold pairs of the Tanks & Temples variant, :230-262,311-392), averaged
// depth, geometric / final masks and the world-space point of every pixel.  The reference re-reads each source depth
// map from disk and makes ~40 numpy passes per pair; here a thread owns a pixel, the source maps are gathered once
// (4 taps each) and nothing intermediate is written.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kMaxSrc = 16;
constexpr int kDynLevels = 11;   // thresholds i / dh_dist, i / dh_rel_diff for i = dh_view_num .. 10

struct FuseViewArgs {
  const float* depth_ref;            // [H][W]
  const float* depth_src[kMaxSrc];   // [Hs][Ws] each
  const double* mats;                // [S][68] (device): GeoMats of every (reference, source) pair
  int S, H, W, Hs, Ws;
  const float* conf[3];
  float photo_thres[3];
  int n_conf;
  FuseMats fm;
  // static mode
  float dmin, dmax;                  // check_geometric_consistency's range test on depth_ref (float32 compare)
  double pix_thres;
  float depth_thres;
  int geo_thres;
  // dynamic mode
  int dyn_view_num;                  // 0: static mode
  double dyn_pix[kDynLevels];        // i / dh_dist        (float64 compare with the float64 distance)
  float dyn_rel[kDynLevels];         // i / dh_rel_diff    (float32 compare, NumPy's weak-scalar rule)
  double avg_min, avg_max;           // dynamic mode: final mask also needs depth_min <= averaged depth <= depth_max
  uint8_t* photo_mask;
  uint8_t* geo_mask;
  uint8_t* final_mask;
  double* depth_avg;
  float* xyz;
};

template <bool DYN>
__global__ void __launch_bounds__(256) fuse_view_kernel(const __grid_constant__ FuseViewArgs a) {
  __shared__ GeoMats mats_s[kMaxSrc];
  {
    double* dst = reinterpret_cast<double*>(mats_s);
    for (int k = threadIdx.x; k < a.S * 68; k += blockDim.x) dst[k] = a.mats[k];
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)a.H * a.W) return;
  const int x = (int)(i % a.W), y = (int)(i / a.W);
  const float dref_f = __ldg(a.depth_ref + i);
  bool photo = true;
  for (int c = 0; c < a.n_conf; ++c) photo = photo && (__ldg(a.conf[c] + i) > a.photo_thres[c]);
  float sum = 0.0f;                  // python's sum(list): ((0 + d_0) + d_1) + ... in float32
  int cnt = 0;                       // geo_mask_sum (static) / geo_mask_sum of the loosest test (dynamic)
  int lvl[kDynLevels];
#pragma unroll
  for (int k = 0; k < kDynLevels; ++k) lvl[k] = 0;
  for (int s = 0; s < a.S; ++s) {
    const Reproj r = reproject_pixel(mats_s[s], a.depth_src[s], a.Hs, a.Ws, x, y, dref_f);
    bool ok;
    if (DYN) {
      ok = false;
#pragma unroll
      for (int k = 0; k < kDynLevels; ++k) {
        if (k >= a.dyn_view_num) {
          const bool m = r.dist < a.dyn_pix[k] && r.rel < a.dyn_rel[k];
          lvl[k] += m ? 1 : 0;
          if (k == kDynLevels - 1) ok = m;     // depth_reproj is masked with the last (i = 10) test, filter.py:260
        }
      }
    } else {
      ok = r.dist < a.pix_thres && r.rel < a.depth_thres && dref_f > a.dmin && dref_f < a.dmax;
    }
    sum = __fadd_rn(sum, ok ? r.drep : 0.0f);
    cnt += ok ? 1 : 0;
  }
  // (sum(depth_reproj) + ref_depth) / (geo_mask_sum + 1): float32 sum, float64 quotient      (filter.py:189, :383)
  const double avg = (double)__fadd_rn(sum, dref_f) / (double)(cnt + 1);
  bool geo;
  if (DYN) {
    geo = cnt >= 10;
#pragma unroll
    for (int k = 0; k < kDynLevels; ++k)
      if (k >= a.dyn_view_num) geo = geo || lvl[k] >= k;
  } else {
    geo = cnt >= a.geo_thres;
  }
  bool fin = photo && geo;
  if (DYN) fin = fin && avg >= a.avg_min && avg <= a.avg_max;
  a.photo_mask[i] = photo ? 1 : 0;
  a.geo_mask[i] = geo ? 1 : 0;
  a.final_mask[i] = fin ? 1 : 0;
  a.depth_avg[i] = avg;
  double cx, cy, cz;
  mat3(a.fm.Kref_inv, (double)x * avg, (double)y * avg, avg, cx, cy, cz);
  double wx, wy, wz;
  mat4_rows3(a.fm.Eref_inv, cx, cy, cz, wx, wy, wz);
  a.xyz[i * 3 + 0] = (float)wx;
  a.xyz[i * 3 + 1] = (float)wy;
  a.xyz[i * 3 + 2] = (float)wz;
}

}  // namespace
}  // namespace dmvs

using namespace dmvs;

extern "C" int dmvs_geo_consistency(const float* depth_ref, const float* depth_src, const double* mats68, float depth_min,
                                    float depth_max, double pix_thres, float depth_thres, uint8_t* mask,
                                    float* depth_reproj, float* x_src, float* y_src, float* sum_reproj, int32_t* count,
                                    int32_t H, int32_t W, int32_t Hs, int32_t Ws, void* stream) {
  if (!depth_ref || !depth_src || !mats68 || !mask || !depth_reproj) return DMVS_ERR_ARG;
  if (H <= 0 || W <= 0 || Hs <= 0 || Ws <= 0) return DMVS_ERR_ARG;
  GeoMats m;
  const double* p = mats68;
  for (int k = 0; k < 9; ++k) m.Kref_inv[k] = *p++;
  for (int k = 0; k < 16; ++k) m.T_rs[k] = *p++;
  for (int k = 0; k < 9; ++k) m.Ksrc[k] = *p++;
  for (int k = 0; k < 9; ++k) m.Ksrc_inv[k] = *p++;
  for (int k = 0; k < 16; ++k) m.T_sr[k] = *p++;
  for (int k = 0; k < 9; ++k) m.Kref[k] = *p++;
  const int64_t total = (int64_t)H * W;
  geo_consistency_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      depth_ref, depth_src, m, depth_min, depth_max, pix_thres, depth_thres, mask, depth_reproj, x_src, y_src, sum_reproj,
      count, H, W, Hs, Ws);
  return launch_status();
}

extern "C" int dmvs_fuse_points(const float* depth_ref, const float* sum_reproj, const int32_t* count,
                                const uint8_t* photo_mask, int32_t geo_thres, const double* mats25, double* depth_avg,
                                uint8_t* geo_mask, uint8_t* final_mask, float* xyz, int32_t H, int32_t W, void* stream) {
  if (!depth_ref || !sum_reproj || !count || !mats25 || !depth_avg || !geo_mask || !final_mask || !xyz) return DMVS_ERR_ARG;
  if (H <= 0 || W <= 0) return DMVS_ERR_ARG;
  FuseMats m;
  const double* p = mats25;
  for (int k = 0; k < 9; ++k) m.Kref_inv[k] = *p++;
  for (int k = 0; k < 16; ++k) m.Eref_inv[k] = *p++;
  const int64_t total = (int64_t)H * W;
  fuse_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      depth_ref, sum_reproj, count, photo_mask, geo_thres, m, depth_avg, geo_mask, final_mask, xyz, H, W);
  return launch_status();
}

extern "C" int dmvs_fuse_view(const float* depth_ref, const float* const* depth_src, const double* mats_dev, int32_t S, int32_t H,
                              int32_t W, int32_t Hs, int32_t Ws, const float* const* conf, const float* photo_thres,
                              int32_t n_conf, const double* mats25, float depth_min, float depth_max, double pix_thres,
                              float depth_thres, int32_t geo_thres, int32_t dyn_view_num, double dyn_dist, double dyn_rel,
                              double avg_min, double avg_max, uint8_t* photo_mask, uint8_t* geo_mask, uint8_t* final_mask,
                              double* depth_avg, float* xyz, void* stream) {
  if (!depth_ref || !depth_src || !mats_dev || !mats25 || !photo_mask || !geo_mask || !final_mask || !depth_avg || !xyz)
    return DMVS_ERR_ARG;
  if (S < 0 || S > kMaxSrc || H <= 0 || W <= 0 || Hs <= 0 || Ws <= 0 || n_conf < 0 || n_conf > 3) return DMVS_ERR_ARG;
  if (dyn_view_num < 0 || dyn_view_num >= kDynLevels) return DMVS_ERR_ARG;
  if (n_conf > 0 && (!conf || !photo_thres)) return DMVS_ERR_ARG;
  FuseViewArgs a = {};
  a.depth_ref = depth_ref;
  for (int s = 0; s < S; ++s) {
    if (!depth_src[s]) return DMVS_ERR_ARG;
    a.depth_src[s] = depth_src[s];
  }
  a.mats = mats_dev;
  a.S = S; a.H = H; a.W = W; a.Hs = Hs; a.Ws = Ws;
  a.n_conf = n_conf;
  for (int c = 0; c < n_conf; ++c) {
    if (!conf[c]) return DMVS_ERR_ARG;
    a.conf[c] = conf[c];
    a.photo_thres[c] = photo_thres[c];
  }
  const double* p = mats25;
  for (int k = 0; k < 9; ++k) a.fm.Kref_inv[k] = *p++;
  for (int k = 0; k < 16; ++k) a.fm.Eref_inv[k] = *p++;
  a.dmin = depth_min; a.dmax = depth_max; a.pix_thres = pix_thres; a.depth_thres = depth_thres; a.geo_thres = geo_thres;
  a.dyn_view_num = dyn_view_num;
  for (int k = 0; k < kDynLevels; ++k) {
    a.dyn_pix[k] = dyn_dist > 0 ? (double)k / dyn_dist : 0.0;
    a.dyn_rel[k] = dyn_rel > 0 ? (float)((double)k / dyn_rel) : 0.0f;
  }
  a.avg_min = avg_min; a.avg_max = avg_max;
  a.photo_mask = photo_mask; a.geo_mask = geo_mask; a.final_mask = final_mask; a.depth_avg = depth_avg; a.xyz = xyz;
  const int64_t total = (int64_t)H * W;
  const unsigned grid = (unsigned)ceil_div64(total, 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dyn_dist > 0)
    fuse_view_kernel<true><<<grid, 256, 0, st>>>(a);
  else
    fuse_view_kernel<false><<<grid, 256, 0, st>>>(a);
  return launch_status();
}
