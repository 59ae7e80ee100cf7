// Conv3d(8 -> 1, k=3, s=1, p=1) marching along depth: the last layer of PixelViewWeight (module.py:454-457, followed by
// sigmoid and the maximum over depth, module.py:459-463) and CostRegNet_small.prob (module.py:439,447).
//
// On the tensor-core back ends a single output channel wastes 7/8 of the narrowest MMA and every input slice is staged
// three times (once per depth tap).  Here a CTA owns a 32 x 8 pixel column of the volume and walks the D input slices
// once: a thread keeps the partial sums of the three output slices an input slice contributes to (od = z+1, z, z-1) in
// registers, the 216 weights are launch parameters (constant-bank operands of the FFMAs, no loads), and the slice tile
// is double buffered in shared memory with cp.async.  fp32 FFMA arithmetic (the DMVS_PREC_FP32 class).
//
// Bound: FFMA issue (216 per output) and the shared-memory reads of the slice (18 x 128-bit per thread and slice);
// HBM traffic is the input volume once plus the output.
#include "common.cuh"

namespace dmvs {
namespace {

constexpr int kTW = 32, kTH = 8;                         // output pixels per CTA (one warp per row)
constexpr int kIW = kTW + 2, kIH = kTH + 2;              // staged slice tile
constexpr int kTilePx = kIW * kIH;                        // 340
constexpr int kCin = 8;

struct To1Args {
  const float* x;      // [N][D][H][W][x_ps]
  float* y;            // mode 0: [N][D][H][W]   mode 1: [N][H][W]
  int x_ps;
  int N, D, H, W;
  int mode;            // 0: y = conv + bias   1: y = max_d sigmoid(conv + bias)
  int segs, seg_len;   // mode 0: the depth range is cut into `segs` segments of `seg_len` output slices (more CTAs for small N)
  float bias;
  float w[27 * kCin];  // [kd][kh][kw][ci]
};

__device__ __forceinline__ void cp_async16_zfill(float* smem_dst, const float* gsrc, bool ok) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int bytes = ok ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(kTW * kTH) conv3d_to1_kernel(const __grid_constant__ To1Args a) {
  // slice tile, planar by channel quad: [buffer][quad][pixel][4] - a warp's 128-bit reads are contiguous
  __shared__ __align__(16) float tile[2][2][kTilePx][4];
  const int tid = threadIdx.x;
  const int lx = tid & 31, ly = tid >> 5;
  const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
  const int n = blockIdx.z / a.segs, seg = blockIdx.z - n * a.segs;
  const int d0 = seg * a.seg_len, d1 = min(d0 + a.seg_len, a.D);          // output slices of this CTA
  const int zs = max(d0 - 1, 0), ze = min(d1 + 1, a.D);                    // input slices it walks
  const int ox = x0 + lx, oy = y0 + ly;
  const bool valid = ox < a.W && oy < a.H;
  pdl_sync();

  // the (quad, pixel) units this thread stages are the same for every slice: offsets and bounds once
  constexpr int kUnits = (2 * kTilePx + kTW * kTH - 1) / (kTW * kTH);   // 3
  int src_off[kUnits], dst_off[kUnits];
#pragma unroll
  for (int i = 0; i < kUnits; ++i) {
    const int u = tid + i * kTW * kTH;
    const int q = u / kTilePx, p = u - q * kTilePx;
    const int r = p / kIW, c = p - r * kIW;
    const int iy = y0 - 1 + r, ix = x0 - 1 + c;
    const bool ok = u < 2 * kTilePx && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
    src_off[i] = ok ? (iy * a.W + ix) * a.x_ps + q * 4 : -1;          // one slice is < 2^31 floats (host-checked)
    dst_off[i] = u < 2 * kTilePx ? (q * kTilePx + p) * 4 : -1;
  }
  const int64_t slice_in = (int64_t)a.H * a.W * a.x_ps;
  auto stage = [&](int z, int buf) {
    const float* slice = a.x + (int64_t)(n * a.D + z) * slice_in;
#pragma unroll
    for (int i = 0; i < kUnits; ++i) {
      if (dst_off[i] < 0) continue;
      const bool ok = src_off[i] >= 0;
      cp_async16_zfill(&tile[buf][0][0][0] + dst_off[i], ok ? slice + src_off[i] : a.x, ok);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };

  stage(zs, 0);
  float p0 = 0.f, p1 = 0.f;          // partial sums of output slices z-1 (taps kd = 0, 1 done) and z (tap kd = 0 done)
  float best = 0.f;                   // mode 1: sigmoid > 0, so 0 is below every candidate
  float* yout = a.y + ((int64_t)n * (a.mode == 0 ? a.D : 1) * a.H + oy) * a.W + ox;
  const int64_t slice_out = (int64_t)a.H * a.W;
  for (int z = zs; z < ze; ++z) {
    const int buf = (z - zs) & 1;
    if (z + 1 < ze) {
      stage(z + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    __syncthreads();
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int p = (ly + kh) * kIW + lx + kw;
        const float4 v0 = *reinterpret_cast<const float4*>(&tile[buf][0][p][0]);
        const float4 v1 = *reinterpret_cast<const float4*>(&tile[buf][1][p][0]);
        const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        const int wb = (kh * 3 + kw) * kCin;
#pragma unroll
        for (int c = 0; c < kCin; ++c) {
          t0 = fmaf(v[c], a.w[0 * 9 * kCin + wb + c], t0);
          t1 = fmaf(v[c], a.w[1 * 9 * kCin + wb + c], t1);
          t2 = fmaf(v[c], a.w[2 * 9 * kCin + wb + c], t2);
        }
      }
    }
    __syncthreads();                 // the buffer is refilled by the next iteration's prefetch
    if (z - 1 >= d0) {               // output slice z-1 is complete: taps kd = 0, 1 (earlier slices) + kd = 2 (this one)
      const float o = (p0 + t2) + a.bias;
      if (a.mode == 0) {
        if (valid) yout[(int64_t)(z - 1) * slice_out] = o;
      } else {
        best = fmaxf(best, sigmoidf_(o));
      }
    }
    p0 = p1 + t1;
    p1 = t0;
  }
  if (d1 == a.D) {                   // last output slice of the volume: its kd = 2 tap reads the zero padding
    const float o = p0 + a.bias;
    if (a.mode == 0) {
      if (valid) yout[(int64_t)(a.D - 1) * slice_out] = o;
    } else {
      best = fmaxf(best, sigmoidf_(o));
    }
  }
  if (a.mode == 1 && valid) *yout = best;
}

}  // namespace
}  // namespace dmvs

using namespace dmvs;

extern "C" int dmvs_conv3d_to1_f32(const float* x, int32_t x_ps, const float* w_host, float bias, float* y, int32_t N,
                                   int32_t D, int32_t H, int32_t W, int32_t mode, void* stream) {
  if (!x || !w_host || !y) return DMVS_ERR_ARG;
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0 || x_ps < kCin || (mode != 0 && mode != 1)) return DMVS_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(x) & 15u) != 0 || (x_ps % 4) != 0) return DMVS_ERR_ALIGN;
  if (N > 65535 || ceil_div(H, kTH) > 65535 || (int64_t)H * W * x_ps > 0x7fffffffLL) return DMVS_ERR_UNSUPPORTED;
  To1Args a;
  a.x = x;
  a.y = y;
  a.x_ps = x_ps;
  a.N = N; a.D = D; a.H = H; a.W = W;
  a.mode = mode;
  a.bias = bias;
  for (int i = 0; i < 27 * kCin; ++i) a.w[i] = w_host[i];
  // small batches: cut the depth range so that about four CTAs per SM exist (each segment re-reads two halo slices);
  // the fused maximum over depth keeps one CTA per pixel column
  const int ctas = ceil_div(W, kTW) * ceil_div(H, kTH) * N;
  int segs = 1;
  if (mode == 0) {
    segs = ceil_div(4 * kNumSMs, ctas);
    if (segs > D / 8) segs = D / 8;
    if (segs < 1) segs = 1;
  }
  a.seg_len = ceil_div(D, segs);
  a.segs = ceil_div(D, a.seg_len);
  if ((int64_t)N * a.segs > 65535) return DMVS_ERR_UNSUPPORTED;
  const dim3 grid(ceil_div(W, kTW), ceil_div(H, kTH), N * a.segs);
  launch_pdl(conv3d_to1_kernel, grid, dim3(kTW * kTH), 0, static_cast<cudaStream_t>(stream), a);
  return launch_status();
}
