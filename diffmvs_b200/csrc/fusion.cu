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
  const double dref = (double)dref_f;
  // reference pixel -> reference camera -> source camera -> source pixel            (filter.py:19-31)
  double rx, ry, rz;
  mat3(m.Kref_inv, (double)x * dref, (double)y * dref, dref, rx, ry, rz);
  double sx, sy, sz;
  mat4_rows3(m.T_rs, rx, ry, rz, sx, sy, sz);
  double kx, ky, kz;
  mat3(m.Ksrc, sx, sy, sz, kx, ky, kz);
  const double u = kx / kz, v = ky / kz;
  const float xs = (float)u, ys = (float)v;
  const float sampled = remap_linear(depth_src, Hs, Ws, xs, ys);                     // :32-33
  // source pixel with the sampled depth -> back to the reference                     (:35-47)
  const double sd = (double)sampled;
  double qx, qy, qz;
  mat3(m.Ksrc_inv, u * sd, v * sd, sd, qx, qy, qz);
  double px, py, pz;
  mat4_rows3(m.T_sr, qx, qy, qz, px, py, pz);
  float drep = (float)pz;
  double ex, ey, ez;
  mat3(m.Kref, px, py, pz, ex, ey, ez);
  if (ex == 0.0) ex = 1e-5;
  if (ey == 0.0) ey = 1e-5;
  if (ez == 0.0) ez = 1e-5;
  double xr = ex / ez, yr = ey / ez;
  xr = fmin(fmax(xr, -1e8), 1e8);   // np.clip keeps NaN; fmin/fmax would drop it - NaN fails the threshold either way
  yr = fmin(fmax(yr, -1e8), 1e8);
  const float xr_f = (float)xr, yr_f = (float)yr;
  // consistency                                                                      (:75-86)
  const double ddx = (double)xr_f - (double)x, ddy = (double)yr_f - (double)y;
  const double dist = sqrt(ddx * ddx + ddy * ddy);
  const float depth_diff = fabsf(__fsub_rn(drep, dref_f));
  const float rel = __fdiv_rn(depth_diff, dref_f);
  const bool ok = dist < pix_thres && rel < depth_thres && dref_f > dmin && dref_f < dmax;
  if (!ok) drep = 0.0f;
  mask[i] = ok ? 1 : 0;
  depth_reproj[i] = drep;
  if (x_src_out) x_src_out[i] = xs;
  if (y_src_out) y_src_out[i] = ys;
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
