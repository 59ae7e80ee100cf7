// Homography warp + group-wise correlation kernels (module.py:181-218, 514-548, 575-667).
//
// The reference materialises warped volumes [B,C,D,H,W] (4.8 GB per ref-view at DTU size) and then
// multiplies/means them.  Here a group of LPP lanes owns one reference pixel: every lane keeps its
// slice of the reference feature in registers, gathers the four bilinear taps of its channel slice
// with 128-bit loads from the channels-last source map (the LPP lanes of a pixel read one contiguous
// segment), and reduces the per-group dot product with warp shuffles.  Nothing but the final
// correlation leaves the SM.
#include "common.cuh"

namespace dmvs {
namespace {

// ---------------------------------------------------------------------------------------------
// [R|t] of P_src * inverse(P_ref), P = [K*E[:3,:4]; 0 0 0 1]  (module.py:188-190, 520-525)
// ---------------------------------------------------------------------------------------------
__device__ void compose_projection(const float* pair, double P[4][4]) {
  // pair = [2][4][4]: extrinsic E, intrinsic K (top-left 3x3).  The reference forms K*E[:3,:4] in fp32.
  const float* E = pair;
  const float* K = pair + 16;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) {
      float s = 0.f;
      for (int k = 0; k < 3; ++k) s = fmaf(K[r * 4 + k], E[k * 4 + c], s);
      P[r][c] = (double)s;
    }
  for (int c = 0; c < 4; ++c) P[3][c] = (double)E[12 + c];
}

__device__ bool invert4(const double A[4][4], double inv[4][4]) {
  double m[4][8];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      m[r][c] = A[r][c];
      m[r][4 + c] = (r == c) ? 1.0 : 0.0;
    }
  for (int col = 0; col < 4; ++col) {
    int piv = col;
    double best = fabs(m[col][col]);
    for (int r = col + 1; r < 4; ++r)
      if (fabs(m[r][col]) > best) { best = fabs(m[r][col]); piv = r; }
    if (best == 0.0) return false;
    if (piv != col)
      for (int c = 0; c < 8; ++c) { double t = m[col][c]; m[col][c] = m[piv][c]; m[piv][c] = t; }
    const double ip = 1.0 / m[col][col];
    for (int c = 0; c < 8; ++c) m[col][c] *= ip;
    for (int r = 0; r < 4; ++r) {
      if (r == col) continue;
      const double f = m[r][col];
      if (f != 0.0)
        for (int c = 0; c < 8; ++c) m[r][c] -= f * m[col][c];
    }
  }
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) inv[r][c] = m[r][4 + c];
  return true;
}

__global__ void compose_homographies_kernel(const float* __restrict__ proj, float* __restrict__ hom, int B, int V) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * (V - 1)) return;
  const int b = i / (V - 1), v = 1 + i % (V - 1);
  double Pr[4][4], Ps[4][4], inv[4][4];
  compose_projection(proj + ((int64_t)b * V + 0) * 32, Pr);
  compose_projection(proj + ((int64_t)b * V + v) * 32, Ps);
  float* out = hom + (int64_t)i * 12;
  if (!invert4(Pr, inv)) {
    for (int k = 0; k < 12; ++k) out[k] = __int_as_float(0x7fc00000);  // NaN: singular reference camera
    return;
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += Ps[r][k] * inv[k][c];
      out[r * 4 + c] = (float)s;
    }
}

// ---------------------------------------------------------------------------------------------
// warp geometry shared by all kernels
// ---------------------------------------------------------------------------------------------
struct Hom {
  float r[9];
  float t[3];
};

__device__ __forceinline__ Hom load_hom(const float* h) {
  Hom H;
  H.r[0] = __ldg(h + 0); H.r[1] = __ldg(h + 1); H.r[2] = __ldg(h + 2);  H.t[0] = __ldg(h + 3);
  H.r[3] = __ldg(h + 4); H.r[4] = __ldg(h + 5); H.r[5] = __ldg(h + 6);  H.t[1] = __ldg(h + 7);
  H.r[6] = __ldg(h + 8); H.r[7] = __ldg(h + 9); H.r[8] = __ldg(h + 10); H.t[2] = __ldg(h + 11);
  return H;
}

// rot @ (x, y, 1): the reference evaluates this with a batched matmul (module.py:201); sum in the
// natural left-to-right order without contraction.
__device__ __forceinline__ void ray_of_pixel(const Hom& H, float x, float y, float ray[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k)
    ray[k] = __fadd_rn(__fadd_rn(__fmul_rn(H.r[k * 3 + 0], x), __fmul_rn(H.r[k * 3 + 1], y)), H.r[k * 3 + 2]);
}

// source pixel coordinates of ray*depth + t (module.py:202-207): z==0 -> +1e-8, negative z unmasked
__device__ __forceinline__ void project(const Hom& H, const float ray[3], float depth, float& u, float& v) {
  const float px = __fadd_rn(__fmul_rn(ray[0], depth), H.t[0]);
  const float py = __fadd_rn(__fmul_rn(ray[1], depth), H.t[1]);
  float pz = __fadd_rn(__fmul_rn(ray[2], depth), H.t[2]);
  if (pz == 0.0f) pz = __fadd_rn(pz, 1e-8f);
  u = __fdiv_rn(px, pz);
  v = __fdiv_rn(py, pz);
}

// grid_sample(align_corners=True, zeros) normalises to [-1,1] and back (module.py:208-215); the round
// trip u -> u/((W-1)/2) - 1 -> ((g+1)/2)*(W-1) is reproduced so tap selection matches at the borders.
__device__ __forceinline__ float grid_roundtrip(float u, int size) {
  const float half = (float)(size - 1) / 2.0f;
  const float g = __fsub_rn(__fdiv_rn(u, half), 1.0f);
  return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(size - 1));
}

struct Taps {
  int off[4];     // pixel offsets (y*Ws + x) of the 4 taps, -1 when outside
  float wgt[4];   // nw, ne, sw, se
};

__device__ __forceinline__ Taps make_taps(float u, float v, int Hs, int Ws) {
  Taps t;
  const float ix = grid_roundtrip(u, Ws);
  const float iy = grid_roundtrip(v, Hs);
  const float fx = floorf(ix), fy = floorf(iy);
  // ATen grid_sampler_2d: weights from the distances to the four integer corners
  const float wx1 = ix - fx, wy1 = iy - fy;
  const float wx0 = (fx + 1.0f) - ix, wy0 = (fy + 1.0f) - iy;
  t.wgt[0] = wx0 * wy0; t.wgt[1] = wx1 * wy0; t.wgt[2] = wx0 * wy1; t.wgt[3] = wx1 * wy1;
  // NaN / huge coordinates: comparisons below are false -> all taps masked
  const bool x0 = fx >= 0.0f && fx <= (float)(Ws - 1);
  const bool x1 = fx + 1.0f >= 0.0f && fx + 1.0f <= (float)(Ws - 1);
  const bool y0 = fy >= 0.0f && fy <= (float)(Hs - 1);
  const bool y1 = fy + 1.0f >= 0.0f && fy + 1.0f <= (float)(Hs - 1);
  const int xi = x0 || x1 ? (int)fx : 0;
  const int yi = y0 || y1 ? (int)fy : 0;
  t.off[0] = (x0 && y0) ? yi * Ws + xi : -1;
  t.off[1] = (x1 && y0) ? yi * Ws + xi + 1 : -1;
  t.off[2] = (x0 && y1) ? (yi + 1) * Ws + xi : -1;
  t.off[3] = (x1 && y1) ? (yi + 1) * Ws + xi + 1 : -1;
  return t;
}

// The LPP lanes of a pixel share every hypothesis: one lane computes the projection and the taps (four IEEE
// divisions and the grid_sample round trip are the expensive part of these kernels), the others receive them.
__device__ __forceinline__ Taps bcast_taps(const Taps& mine, int src_lane) {
  Taps t;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    t.off[i] = __shfl_sync(0xffffffffu, mine.off[i], src_lane);
    t.wgt[i] = __shfl_sync(0xffffffffu, mine.wgt[i], src_lane);
  }
  return t;
}

__device__ __forceinline__ Taps masked_taps() {
  Taps t;
#pragma unroll
  for (int i = 0; i < 4; ++i) { t.off[i] = -1; t.wgt[i] = 0.0f; }
  return t;
}

// bilinear sample of VEC float4s starting at channel `c0` of a channels-last map with pixel stride ps
template <int VEC>
__device__ __forceinline__ void gather(const float* __restrict__ src, int ps, int c0, const Taps& t, float4 out[VEC]) {
#pragma unroll
  for (int k = 0; k < VEC; ++k) out[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (t.off[i] < 0) continue;
    const float* p = src + (int64_t)t.off[i] * ps + c0;
    const float w = t.wgt[i];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float4 f = ldg4(p + k * 4);
      out[k].x = fmaf(f.x, w, out[k].x);
      out[k].y = fmaf(f.y, w, out[k].y);
      out[k].z = fmaf(f.z, w, out[k].z);
      out[k].w = fmaf(f.w, w, out[k].w);
    }
  }
}

template <int VEC>
__device__ __forceinline__ float dot_vec(const float4 a[VEC], const float4 b[VEC]) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    s = fmaf(a[k].x, b[k].x, s);
    s = fmaf(a[k].y, b[k].y, s);
    s = fmaf(a[k].z, b[k].z, s);
    s = fmaf(a[k].w, b[k].w, s);
  }
  return s;
}

// ---------------------------------------------------------------------------------------------
// differentiable_warping as a stand-alone operator (materialises [B][D][H][W][C])
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) warp_volume_kernel(const float* __restrict__ src, int src_ps,
                                                          const float* __restrict__ hom,
                                                          const float* __restrict__ depth, float* __restrict__ out,
                                                          int B, int C, int Hs, int Ws, int D, int H, int W) {
  pdl_sync();
  // one thread per (b, d, y, x, 4-channel group)
  const int c4n = (C + 3) / 4;
  const int64_t total = (int64_t)B * D * H * W * c4n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % c4n);
    int64_t p = i / c4n;
    const int x = (int)(p % W);
    int64_t q = p / W;
    const int y = (int)(q % H);
    q /= H;
    const int dd = (int)(q % D);
    const int b = (int)(q / D);
    const Hom Hm = load_hom(hom + (int64_t)b * 12);
    float ray[3];
    ray_of_pixel(Hm, (float)x, (float)y, ray);
    float u, v;
    project(Hm, ray, __ldg(depth + p), u, v);
    const Taps t = make_taps(u, v, Hs, Ws);
    const float* sb = src + (int64_t)b * Hs * Ws * src_ps;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < 4; ++k) {
      if (t.off[k] < 0) continue;
      for (int e = 0; e < 4; ++e) {
        const int c = c4 * 4 + e;
        if (c < C) acc[e] = fmaf(__ldg(sb + (int64_t)t.off[k] * src_ps + c), t.wgt[k], acc[e]);
      }
    }
    for (int e = 0; e < 4; ++e) {
      const int c = c4 * 4 + e;
      if (c < C) out[p * C + c] = acc[e];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Stage-1 plane sweep: cor[b][v][d][y][x][g]
// LPP lanes per pixel, each lane owns VEC float4 (= C/LPP channels), G groups -> LPP/G lanes per group.
// ---------------------------------------------------------------------------------------------
template <int C, int G, int LPP>
__global__ void __launch_bounds__(256) plane_sweep_kernel(const float* __restrict__ feats,
                                                          const float* __restrict__ hom,
                                                          const float* __restrict__ plane_depth,
                                                          float* __restrict__ cor, int B, int V, int D, int H, int W) {
  pdl_sync();
  constexpr int VEC = C / (4 * LPP);
  constexpr int LPG = LPP / G;  // lanes per group
  static_assert(VEC >= 1 && LPG >= 1 && C % (4 * LPP) == 0 && LPP % G == 0, "bad split");
  const int HW = H * W;
  const int pix_per_block = 256 / LPP;
  const int sub = threadIdx.x % LPP;
  const int pix = blockIdx.x * pix_per_block + threadIdx.x / LPP;
  const int v1 = blockIdx.y;                 // source view index - 1
  const int b = blockIdx.z;
  const bool active = pix < HW;
  const int p = active ? pix : HW - 1;
  const int x = p % W, y = p / W;
  const int c0 = sub * VEC * 4;
  const float cpg = (float)(C / G);  // torch.mean divides the sum by the count

  const float* ref = feats + ((int64_t)(0 * B + b) * HW + p) * C + c0;
  const float* src = feats + (int64_t)((v1 + 1) * B + b) * HW * C;
  float4 rf[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) rf[k] = ldg4(ref + k * 4);
  const Hom Hm = load_hom(hom + ((int64_t)b * (V - 1) + v1) * 12);
  float ray[3];
  ray_of_pixel(Hm, (float)x, (float)y, ray);

  float* out = cor + ((int64_t)(b * (V - 1) + v1) * D) * HW * G;
  const int lane_base = (threadIdx.x & 31) & ~(LPP - 1);   // first lane of this pixel's group
  for (int d0 = 0; d0 < D; d0 += LPP) {
    // lane `sub` of the group prepares plane d0 + sub, then the group walks the LPP planes together
    Taps mine = masked_taps();
    if (d0 + sub < D) {
      float u, vv;
      project(Hm, ray, __ldg(plane_depth + b * D + d0 + sub), u, vv);
      mine = make_taps(u, vv, H, W);
    }
#pragma unroll
    for (int j = 0; j < LPP; ++j) {
      const int d = d0 + j;
      if (d >= D) break;                                   // uniform across the warp
      const Taps t = bcast_taps(mine, lane_base + j);
      float4 wf[VEC];
      gather<VEC>(src, C, c0, t, wf);
      float s = dot_vec<VEC>(rf, wf);
#pragma unroll
      for (int o = 1; o < LPG; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (active && (sub % LPG) == 0) out[((int64_t)d * HW + p) * G + sub / LPG] = __fdiv_rn(s, cpg);
    }
  }
}

__global__ void view_weight_max_kernel(const float* __restrict__ logit, float* __restrict__ w, int N, int D, int HW) {
  pdl_sync();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * HW) return;
  const int n = (int)(i / HW), p = (int)(i % HW);
  const float* l = logit + (int64_t)n * D * HW + p;
  float m = -INFINITY;
  for (int d = 0; d < D; ++d) m = fmaxf(m, sigmoidf_(__ldg(l + (int64_t)d * HW)));
  w[i] = m;
}

__global__ void aggregate_views_kernel(const float* __restrict__ cor, const float* __restrict__ w,
                                       float* __restrict__ vol, int B, int V1, int D, int HW, int G4) {
  pdl_sync();
  // G == 4: one float4 per (b, d, p)
  const int64_t total = (int64_t)B * D * HW;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int p = (int)(i % HW);
  const int64_t bd = i / HW;
  const int d = (int)(bd % D);
  const int b = (int)(bd / D);
  float wsum = 1e-8f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int v = 0; v < V1; ++v) {
    const float wv = __ldg(w + ((int64_t)b * V1 + v) * HW + p);
    const float4 c = ldg4(cor + ((((int64_t)b * V1 + v) * D + d) * HW + p) * 4);
    wsum = __fadd_rn(wsum, wv);
    acc.x = __fadd_rn(acc.x, __fmul_rn(wv, c.x));
    acc.y = __fadd_rn(acc.y, __fmul_rn(wv, c.y));
    acc.z = __fadd_rn(acc.z, __fmul_rn(wv, c.z));
    acc.w = __fadd_rn(acc.w, __fmul_rn(wv, c.w));
  }
  float4 r;
  r.x = __fdiv_rn(acc.x, wsum); r.y = __fdiv_rn(acc.y, wsum);
  r.z = __fdiv_rn(acc.z, wsum); r.w = __fdiv_rn(acc.w, wsum);
  *reinterpret_cast<float4*>(vol + i * 4) = r;
}

// ---------------------------------------------------------------------------------------------
// softmax over D, expected index, window confidence (module.py:554-571)
// ---------------------------------------------------------------------------------------------
__global__ void depth_regression_kernel(const float* __restrict__ logits, const float* __restrict__ depth_min,
                                        const float* __restrict__ depth_max, float* __restrict__ norm_inv,
                                        float* __restrict__ depth, float* __restrict__ conf,
                                        int32_t* __restrict__ floor_idx, int B, int D, int HW) {
  pdl_sync();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * HW) return;
  const int b = (int)(i / HW), p = (int)(i % HW);
  const float* l = logits + (int64_t)b * D * HW + p;
  float m = -INFINITY;
  for (int d = 0; d < D; ++d) m = fmaxf(m, __ldg(l + (int64_t)d * HW));
  float sum = 0.f;
  for (int d = 0; d < D; ++d) sum += expf(__ldg(l + (int64_t)d * HW) - m);
  float idx = 0.f;
  for (int d = 0; d < D; ++d) {
    const float pd = __fdiv_rn(expf(__ldg(l + (int64_t)d * HW) - m), sum);
    idx = __fadd_rn(idx, __fmul_rn((float)d, pd));
  }
  int j = (int)idx;  // .long(): truncation
  j = j < 0 ? 0 : (j > D - 1 ? D - 1 : j);
  // 4*avg_pool3d over planes j-1..j+2 of the zero-padded probabilities
  float win = 0.f;
  for (int k = -1; k <= 2; ++k) {
    const int d = j + k;
    const float pd = (d >= 0 && d < D) ? __fdiv_rn(expf(__ldg(l + (int64_t)d * HW) - m), sum) : 0.f;
    win = __fadd_rn(win, pd);
  }
  const float n = __fdiv_rn(idx, (float)D - 1.0f);
  const DepthRange rng(__ldg(depth_min + b), __ldg(depth_max + b));
  norm_inv[i] = n;
  depth[i] = rng.to_depth(n);
  conf[i] = __fmul_rn(4.0f, __fmul_rn(win, 0.25f));
  if (floor_idx) floor_idx[i] = j;
}

// ---------------------------------------------------------------------------------------------
// GetCost: hypothesis sampler + all source views + weighted mean, one launch per iteration
// ---------------------------------------------------------------------------------------------
template <int C, int G, int LPP, int D>
__global__ void __launch_bounds__(256) get_cost_kernel(const float* __restrict__ feats, const float* __restrict__ hom,
                                                       const float* __restrict__ inv_depth,
                                                       const float* __restrict__ conf, int conf_ps,
                                                       const float* __restrict__ view_w,
                                                       const float* __restrict__ depth_min,
                                                       const float* __restrict__ depth_max, float* __restrict__ cost,
                                                       int cost_ps, float* __restrict__ samples, int samp_ps, int B,
                                                       int V, int H, int W, int wshift, float interval,
                                                       float min_radius, float max_radius) {
  pdl_sync();
  constexpr int VEC = C / (4 * LPP);
  constexpr int LPG = LPP / G;
  static_assert(VEC >= 1 && LPG >= 1 && C % (4 * LPP) == 0 && LPP % G == 0, "bad split");
  const int HW = H * W;
  const int pix_per_block = 256 / LPP;
  const int sub = threadIdx.x % LPP;
  const int pix = blockIdx.x * pix_per_block + threadIdx.x / LPP;
  const int b = blockIdx.y;
  const bool active = pix < HW;
  const int p = active ? pix : HW - 1;
  const int x = p % W, y = p / W;
  const int c0 = sub * VEC * 4;
  const float cpg = (float)(C / G);  // torch.mean divides the sum by the count
  const DepthRange rng(__ldg(depth_min + b), __ldg(depth_max + b));

  // hypotheses (module.py:250-277)
  const float cur = __ldg(inv_depth + (int64_t)b * HW + p);
  float lo, hi;
  if (conf == nullptr) {
    const float r = (float)(D / 2) * interval;
    lo = __fsub_rn(cur, r);
    hi = __fadd_rn(cur, r);
  } else {
    const float r0 = (float)(D / 2) * interval;
    const float rmin = __fmul_rn(min_radius, r0), rmax = __fmul_rn(max_radius, r0);
    const float cf = __ldg(conf + ((int64_t)b * HW + p) * conf_ps);
    const float r = __fadd_rn(rmin, __fmul_rn(__fsub_rn(1.0f, cf), __fsub_rn(rmax, rmin)));
    lo = __fsub_rn(cur, r);
    hi = __fadd_rn(cur, r);
  }
  // lane `sub` of the pixel's group owns hypotheses sub, sub + LPP, ...: it computes their sample, metric depth and,
  // per source view, their projection and taps; the other lanes of the group receive the taps by shuffle
  constexpr int NCHUNK = (D + LPP - 1) / LPP;
  float my_samp[NCHUNK], my_dep[NCHUNK];
  const float step = D > 1 ? __fdiv_rn(__fsub_rn(hi, lo), (float)(D - 1)) : 0.0f;
#pragma unroll
  for (int j = 0; j < NCHUNK; ++j) {
    const int d = j * LPP + sub;
    float sv = cur;
    if (D > 1) sv = fminf(fmaxf(__fadd_rn(__fmul_rn((float)d, step), lo), 0.0f), 1.0f);
    my_samp[j] = sv;
    my_dep[j] = rng.to_depth(sv);
  }
  const int lane_base = (threadIdx.x & 31) & ~(LPP - 1);   // first lane of this pixel's group

  const float* ref = feats + ((int64_t)b * HW + p) * C + c0;
  float4 rf[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) rf[k] = ldg4(ref + k * 4);

  float acc[D];
#pragma unroll
  for (int d = 0; d < D; ++d) acc[d] = 0.f;
  float wsum = 1e-8f;
  const int Hw = H >> wshift, Ww = W >> wshift;
  const int wp = (y >> wshift) * Ww + (x >> wshift);
  for (int v = 1; v < V; ++v) {
    const Hom Hm = load_hom(hom + ((int64_t)b * (V - 1) + (v - 1)) * 12);
    const float* src = feats + (int64_t)(v * B + b) * HW * C;
    const float wv = __ldg(view_w + ((int64_t)b * (V - 1) + (v - 1)) * Hw * Ww + wp);
    wsum = __fadd_rn(wsum, wv);
    float ray[3];
    ray_of_pixel(Hm, (float)x, (float)y, ray);
#pragma unroll
    for (int jc = 0; jc < NCHUNK; ++jc) {
      Taps mine = masked_taps();
      if (jc * LPP + sub < D) {
        float u, vv;
        project(Hm, ray, my_dep[jc], u, vv);
        mine = make_taps(u, vv, H, W);
      }
#pragma unroll
      for (int j = 0; j < LPP; ++j) {
        const int d = jc * LPP + j;
        if (d < D) {
          const Taps t = bcast_taps(mine, lane_base + j);
          float4 wf[VEC];
          gather<VEC>(src, C, c0, t, wf);
          float s = dot_vec<VEC>(rf, wf);
#pragma unroll
          for (int o = 1; o < LPG; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          acc[d] = __fadd_rn(acc[d], __fmul_rn(wv, __fdiv_rn(s, cpg)));
        }
      }
    }
  }
  if (!active) return;
  if ((sub % LPG) == 0) {
    const int g = sub / LPG;
    float* cp = cost + ((int64_t)b * HW + p) * cost_ps + g * D;
#pragma unroll
    for (int d = 0; d < D; ++d) cp[d] = __fdiv_rn(acc[d], wsum);
  }
  float* sp = samples + ((int64_t)b * HW + p) * samp_ps;
#pragma unroll
  for (int j = 0; j < NCHUNK; ++j)
    if (j * LPP + sub < D) sp[j * LPP + sub] = my_samp[j];   // every lane stores the samples it owns
}

}  // namespace
}  // namespace dmvs

using namespace dmvs;

extern "C" int dmvs_compose_homographies(const float* proj, float* hom, int32_t B, int32_t V, void* stream) {
  if (!proj || !hom || B <= 0 || V < 2) return DMVS_ERR_ARG;
  const int n = B * (V - 1);
  launch_pdl(compose_homographies_kernel, dim3(ceil_div(n, 64)), dim3(64), 0, static_cast<cudaStream_t>(stream), proj, hom, B, V);
  return launch_status();
}

extern "C" int dmvs_warp_volume(const float* src, int32_t src_ps, const float* hom, const float* depth, float* out,
                                int32_t B, int32_t C, int32_t Hs, int32_t Ws, int32_t D, int32_t H, int32_t W,
                                void* stream) {
  if (!src || !hom || !depth || !out) return DMVS_ERR_ARG;
  if (B <= 0 || C <= 0 || Hs <= 0 || Ws <= 0 || D <= 0 || H <= 0 || W <= 0 || src_ps < C) return DMVS_ERR_ARG;
  const int64_t total = (int64_t)B * D * H * W * ((C + 3) / 4);
  const int64_t want = ceil_div64(total, 256);
  const int blocks = (int)(want < (int64_t)kNumSMs * 32 ? want : (int64_t)kNumSMs * 32);
  launch_pdl(warp_volume_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), src, src_ps, hom, depth, out, B, C, Hs, Ws,
                                                                           D, H, W);
  return launch_status();
}

extern "C" int dmvs_plane_sweep_corr(const float* feats, const float* hom, const float* plane_depth, float* cor,
                                     int32_t B, int32_t V, int32_t C, int32_t G, int32_t D, int32_t H, int32_t W,
                                     void* stream) {
  if (!feats || !hom || !plane_depth || !cor) return DMVS_ERR_ARG;
  if (B <= 0 || V < 2 || D <= 0 || H <= 0 || W <= 0) return DMVS_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(feats) & 15u) != 0) return DMVS_ERR_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int HW = H * W;
  if (G != 4) return DMVS_ERR_UNSUPPORTED;
  if (C == 48) {
    dim3 grid(ceil_div(HW, 256 / 4), V - 1, B);
    launch_pdl(plane_sweep_kernel<48, 4, 4>, dim3(grid), dim3(256), 0, st, feats, hom, plane_depth, cor, B, V, D, H, W);
  } else if (C == 32) {
    dim3 grid(ceil_div(HW, 256 / 8), V - 1, B);
    launch_pdl(plane_sweep_kernel<32, 4, 8>, dim3(grid), dim3(256), 0, st, feats, hom, plane_depth, cor, B, V, D, H, W);
  } else if (C == 16) {
    dim3 grid(ceil_div(HW, 256 / 4), V - 1, B);
    launch_pdl(plane_sweep_kernel<16, 4, 4>, dim3(grid), dim3(256), 0, st, feats, hom, plane_depth, cor, B, V, D, H, W);
  } else {
    return DMVS_ERR_UNSUPPORTED;
  }
  return launch_status();
}

extern "C" int dmvs_view_weight_max(const float* logit, float* w, int32_t N, int32_t D, int32_t HW, void* stream) {
  if (!logit || !w || N <= 0 || D <= 0 || HW <= 0) return DMVS_ERR_ARG;
  const int64_t total = (int64_t)N * HW;
  launch_pdl(view_weight_max_kernel, dim3((unsigned)ceil_div64(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), logit, w, N, D, HW);
  return launch_status();
}

extern "C" int dmvs_aggregate_views(const float* cor, const float* w, float* vol, int32_t B, int32_t V1, int32_t D,
                                    int32_t HW, int32_t G, void* stream) {
  if (!cor || !w || !vol || B <= 0 || V1 <= 0 || D <= 0 || HW <= 0) return DMVS_ERR_ARG;
  if (G != 4) return DMVS_ERR_UNSUPPORTED;
  const int64_t total = (int64_t)B * D * HW;
  launch_pdl(aggregate_views_kernel, dim3((unsigned)ceil_div64(total, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), cor, w, vol, B, V1,
                                                                                                        D, HW, G);
  return launch_status();
}

extern "C" int dmvs_depth_regression(const float* logits, const float* depth_min, const float* depth_max,
                                     float* norm_inv, float* depth, float* conf, int32_t* floor_idx, int32_t B,
                                     int32_t D, int32_t HW, void* stream) {
  if (!logits || !depth_min || !depth_max || !norm_inv || !depth || !conf || B <= 0 || D <= 0 || HW <= 0)
    return DMVS_ERR_ARG;
  const int64_t total = (int64_t)B * HW;
  launch_pdl(depth_regression_kernel, dim3((unsigned)ceil_div64(total, 128)), dim3(128), 0, static_cast<cudaStream_t>(stream), 
      logits, depth_min, depth_max, norm_inv, depth, conf, floor_idx, B, D, HW);
  return launch_status();
}

namespace {
template <int C, int LPP>
int launch_get_cost(int D, const float* feats, const float* hom, const float* inv_depth, const float* conf,
                    int conf_ps, const float* view_w, const float* depth_min, const float* depth_max, float* cost, int cost_ps,
                    float* samples, int samp_ps, int B, int V, int H, int W, int wshift, float interval, float rmin,
                    float rmax, cudaStream_t st) {
  dim3 grid(ceil_div(H * W, 256 / LPP), B);
#define DMVS_GC(DD)                                                                                               \
  launch_pdl(get_cost_kernel<C, 4, LPP, DD>, dim3(grid), dim3(256), 0, st, feats, hom, inv_depth, conf, conf_ps, view_w, depth_min, depth_max, cost, \
                                                       cost_ps, samples, samp_ps, B, V, H, W, wshift, interval, rmin, rmax)
  switch (D) {
    case 1: DMVS_GC(1); break;
    case 2: DMVS_GC(2); break;
    case 4: DMVS_GC(4); break;
    case 6: DMVS_GC(6); break;
    case 8: DMVS_GC(8); break;
    default: return DMVS_ERR_UNSUPPORTED;
  }
#undef DMVS_GC
  return launch_status();
}
}  // namespace

extern "C" int dmvs_get_cost(const float* feats, const float* hom, const float* inv_depth, const float* conf,
                             int32_t conf_ps, const float* view_w, const float* depth_min, const float* depth_max, float* cost,
                             int32_t cost_ps, float* samples, int32_t samp_ps, int32_t B, int32_t V, int32_t C,
                             int32_t G, int32_t D, int32_t H, int32_t W, int32_t wshift, float interval,
                             float min_radius, float max_radius, void* stream) {
  if (!feats || !hom || !inv_depth || !view_w || !depth_min || !depth_max || !cost || !samples) return DMVS_ERR_ARG;
  if (B <= 0 || V < 2 || H <= 0 || W <= 0 || wshift < 0 || cost_ps < G * D || samp_ps < D) return DMVS_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(feats) & 15u) != 0) return DMVS_ERR_ALIGN;
  if (G != 4) return DMVS_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (C == 32)
    return launch_get_cost<32, 8>(D, feats, hom, inv_depth, conf, conf_ps, view_w, depth_min, depth_max, cost, cost_ps, samples,
                                  samp_ps, B, V, H, W, wshift, interval, min_radius, max_radius, st);
  if (C == 16)
    return launch_get_cost<16, 4>(D, feats, hom, inv_depth, conf, conf_ps, view_w, depth_min, depth_max, cost, cost_ps, samples,
                                  samp_ps, B, V, H, W, wshift, interval, min_radius, max_radius, st);
  if (C == 48)
    return launch_get_cost<48, 4>(D, feats, hom, inv_depth, conf, conf_ps, view_w, depth_min, depth_max, cost, cost_ps, samples,
                                  samp_ps, B, V, H, W, wshift, interval, min_radius, max_radius, st);
  return DMVS_ERR_UNSUPPORTED;
}
