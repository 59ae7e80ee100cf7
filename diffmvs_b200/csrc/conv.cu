// Direct fp32 convolution (2-D / 3-D, channels-last) with fused prologue/epilogue, and the 3-D
// transposed convolution of CostRegNet_small.  See include/diffmvs_b200.h for the contract.
//
// Design (B200, HBM/L2-bound layers with 3..64 channels):
//   * one CTA = 256 threads = 8 warps computes a 32-wide output tile; a warp owns PX output rows
//     (lane = x) and CO_T output channels, accumulating PX*CO_T values in registers;
//   * the input tile (with halo) and the weight slab of one input-channel chunk are staged in shared
//     memory; activations are read as 128-bit vectors along channels with a padded pixel pitch
//     (bank-conflict free), weights as warp-uniform broadcast vectors;
//   * GroupNorm+SiLU of the producer layer is applied while staging (no extra pass over HBM), and
//     the GroupNorm statistics of this layer's output are reduced in the epilogue.
#include <cstdlib>

#include "conv_common.cuh"

namespace dmvs {
namespace {

constexpr int kThreads = kConvThreads;

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// c = a * b + c on both fp32 halves (two independent round-to-nearest FMAs)
__device__ __forceinline__ void fma2(unsigned long long& c, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
}

template <int CO_T, int WC, int PX, int S>
__global__ void __launch_bounds__(kThreads, 2) conv_kernel(const __grid_constant__ ConvArgs a) {
  constexpr int WP = 8 / WC;           // warps along output rows
  constexpr int TH = WP * PX;          // output rows per CTA
  constexpr int COUT_S = CO_T * WC;    // output channels per CTA
  constexpr int N4 = COUT_S / 4;       // channel quads per CTA
  constexpr int OP = COUT_S + 4;       // pitch of the staged output tile (bank-conflict free)
  const dmvs_conv_desc& d = a.d;

  extern __shared__ __align__(16) float smem[];
  float* in_s = smem;
  float* w_s = in_s + a.in_rows * a.in_cols * a.CKP;
  float* gn_s = w_s + d.KH * d.KW * a.CK * COUT_S;   // [2][C1] when in_stats
  __shared__ unsigned long long stat_s[8];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int wc = warp % WC;
  const int wp = warp / WC;

  pdl_sync();
  const int n = blockIdx.z / d.Do;
  const int od = blockIdx.z - n * d.Do;
  const int ty0 = blockIdx.y * TH;
  const int tx0 = blockIdx.x * kTileW;
  const int iy0 = ty0 * S - d.pad_h;
  const int ix0 = tx0 * S - d.pad_w;

  if (tid < 8) stat_s[tid] = 0ull;
  if (d.in_stats != nullptr) {
    for (int c = tid; c < d.C1; c += kThreads) groupnorm_affine(d, n, c, gn_s);
  }

  // Accumulators are kept as packed fp32x2 pairs {c(2k), c(2k+1)}: Blackwell's FFMA2 (`fma.rn.f32x2`) retires two
  // IEEE fp32 FMAs per issue slot (same 74 TFLOP/s peak as FFMA, half the instructions -
  // tools/probes/ffma2_probe.cu), which is what this issue-bound loop needs.
  unsigned long long acc2[PX][CO_T / 2];
#pragma unroll
  for (int p = 0; p < PX; ++p)
#pragma unroll
    for (int j = 0; j < CO_T / 2; ++j) acc2[p][j] = 0ull;

  const int ck4 = a.CK >> 2;
  const int w_rows = d.KH * d.KW * a.CK;
  const int w_units = w_rows * N4;

  for (int kd = 0; kd < d.KD; ++kd) {
    const int id = od * S + kd - d.pad_d;
    if (id < 0 || id >= d.D) continue;  // zero padding along depth (uniform for the CTA)
    for (int c0 = 0; c0 < a.cin_pad; c0 += a.CK) {
      __syncthreads();  // previous chunk fully consumed (also orders gn_s / stat_s init)
      stage_input_tile(a, in_s, gn_s, n, id, iy0, ix0, c0);
      // ---- stage the weight slab [KH*KW][CK][COUT_S] ----------------------------------------
#pragma unroll 1
      for (int idx = tid; idx < w_units; idx += kThreads) {
        const int j4 = idx % N4;
        const int r = idx / N4;
        const int ci = r & (a.CK - 1);
        const int tap = r / a.CK;
        const bool ok = c0 + ci < a.cin_pad;
        const int64_t off = ((int64_t)((kd * d.KH * d.KW + tap) * a.cin_pad + c0 + ci)) * a.w_cstride + a.co_base + j4 * 4;
        cp_async16(w_s + r * COUT_S + j4 * 4, ok ? d.w + off : d.w, ok);
      }
      cp_async_wait_all();
      __syncthreads();
      // ---- accumulate -----------------------------------------------------------------------
      const float* in_base = in_s + ((wp * PX * S) * a.in_cols + lane * S) * a.CKP;
      const int row_pitch = S * a.in_cols * a.CKP;
#pragma unroll 1
      for (int kh = 0; kh < d.KH; ++kh) {
#pragma unroll 1
        for (int kw = 0; kw < d.KW; ++kw) {
          const float* ip = in_base + (kh * a.in_cols + kw) * a.CKP;
          const float* wt = w_s + ((kh * d.KW + kw) * a.CK) * COUT_S + wc * CO_T;
#pragma unroll 1
          for (int c4 = 0; c4 < ck4; ++c4) {
            float av[PX][4];
#pragma unroll
            for (int p = 0; p < PX; ++p) {
              const float4 t4 = *reinterpret_cast<const float4*>(ip + p * row_pitch + c4 * 4);
              av[p][0] = t4.x; av[p][1] = t4.y; av[p][2] = t4.z; av[p][3] = t4.w;
            }
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) {
              const float* wr = wt + (c4 * 4 + ci) * COUT_S;
              unsigned long long aa[PX];   // {a, a}
#pragma unroll
              for (int p = 0; p < PX; ++p) aa[p] = pack2(av[p][ci], av[p][ci]);
#pragma unroll
              for (int j4 = 0; j4 < CO_T / 4; ++j4) {
                const float4 wv = *reinterpret_cast<const float4*>(wr + j4 * 4);
                const unsigned long long w01 = pack2(wv.x, wv.y), w23 = pack2(wv.z, wv.w);
#pragma unroll
                for (int p = 0; p < PX; ++p) {
                  fma2(acc2[p][j4 * 2 + 0], aa[p], w01);
                  fma2(acc2[p][j4 * 2 + 1], aa[p], w23);
                }
              }
            }
          }
        }
      }
    }
  }

  // ---- epilogue: accumulators -> shared tile -> compact, fully coalesced write-out ---------------
  __syncthreads();  // every warp is done with in_s / w_s
  float* out_s = smem;  // [TH*32][OP]
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    float* op = out_s + ((wp * PX + p) * kTileW + lane) * OP + wc * CO_T;
#pragma unroll
    for (int j4 = 0; j4 < CO_T / 4; ++j4) {
      float c0, c1, c2, c3;
      unpack2(acc2[p][j4 * 2 + 0], c0, c1);
      unpack2(acc2[p][j4 * 2 + 1], c2, c3);
      *reinterpret_cast<float4*>(op + j4 * 4) = make_float4(c0, c1, c2, c3);
    }
  }
  __syncthreads();

  epilogue_tile<TH, COUT_S>(a, out_s, stat_s, n, od, ty0, tx0);
}

// ---------------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------------
using KernelFn = void (*)(const ConvArgs);

template <int CO_T, int WC, int PX, int S>
KernelFn get_kernel() {
  static SmemOptIn opt_in;
  KernelFn fn = conv_kernel<CO_T, WC, PX, S>;
  opt_in.ensure(fn, 200 * 1024);
  return fn;
}

template <int CO_T, int WC>
KernelFn pick_px(int px, int s) {
  constexpr int kMaxPx = (CO_T >= 32) ? 2 : 4;
  if (px > kMaxPx) px = kMaxPx;
  if (s == 1) {
    if (px >= 4) { if constexpr (kMaxPx >= 4) return get_kernel<CO_T, WC, 4, 1>(); }
    if (px >= 2) return get_kernel<CO_T, WC, 2, 1>();
    return get_kernel<CO_T, WC, 1, 1>();
  }
  if (px >= 4) { if constexpr (kMaxPx >= 4) return get_kernel<CO_T, WC, 4, 2>(); }
  if (px >= 2) return get_kernel<CO_T, WC, 2, 2>();
  return get_kernel<CO_T, WC, 1, 2>();
}

KernelFn pick_kernel(int chunk, int px, int s, int* co_t, int* wc) {
  switch (chunk) {
    case 4: *co_t = 4; *wc = 1; return pick_px<4, 1>(px, s);
    case 8: *co_t = 8; *wc = 1; return pick_px<8, 1>(px, s);
    case 16: *co_t = 16; *wc = 1; return pick_px<16, 1>(px, s);
    case 32: *co_t = 32; *wc = 1; return pick_px<32, 1>(px, s);
    case 64: *co_t = 32; *wc = 2; return pick_px<32, 2>(px, s);
    default: *co_t = 32; *wc = 4; return pick_px<32, 4>(px, s);
  }
}


}  // namespace

#ifndef DMVS_LEGACY_BACKENDS
// The round-1 back ends (mma.sync implicit GEMM, tap-offset tcgen05) live in csrc/legacy/ and are only built with
// DMVS_BUILD_LEGACY=1: no shipped configuration selects them (the autotuner never picked them at cfg3 / cfg4).
int dispatch_conv_mma(const dmvs_conv_desc&, cudaStream_t) { return DMVS_ERR_UNSUPPORTED; }
bool conv_tc_supported(const dmvs_conv_desc&) { return false; }
int dispatch_conv_tc(const dmvs_conv_desc&, cudaStream_t) { return DMVS_ERR_UNSUPPORTED; }
#endif

}  // namespace dmvs

using namespace dmvs;

extern "C" int dmvs_conv_f32(const dmvs_conv_desc* dp, void* stream) {
  if (dp == nullptr) return DMVS_ERR_ARG;
  const dmvs_conv_desc& d = *dp;
  if (!d.x || !d.w || !d.y) return DMVS_ERR_ARG;
  if (d.N <= 0 || d.D <= 0 || d.H <= 0 || d.W <= 0 || d.C1 <= 0 || d.C2 < 0 || d.Cout <= 0) return DMVS_ERR_ARG;
  if (d.C2 > 0 && !d.x2) return DMVS_ERR_ARG;
  if (d.KD <= 0 || d.KH <= 0 || d.KW <= 0 || (d.stride != 1 && d.stride != 2)) return DMVS_ERR_ARG;
  if (d.Do <= 0 || d.Ho <= 0 || d.Wo <= 0) return DMVS_ERR_ARG;
  if (d.res_mode != DMVS_RES_NONE && !d.res) return DMVS_ERR_ARG;
  if (d.epi != DMVS_EPI_STD && (!d.aux1 || (d.epi == DMVS_EPI_GRU_Q && !d.aux2))) return DMVS_ERR_ARG;
  if (d.in_stats && (d.C2 != 0 || (d.C1 % 4) != 0 || !d.in_g1 || !d.in_g0)) return DMVS_ERR_ARG;
  if (d.out_stats && (d.Cout % 4) != 0) return DMVS_ERR_ARG;
  if ((d.in_up2 || d.res_up2) && (d.D != 1 || d.KD != 1)) return DMVS_ERR_UNSUPPORTED;
  if (d.in_up2 && ((d.H | d.W) & 1)) return DMVS_ERR_ARG;
  if (d.res_up2 && ((d.Ho | d.Wo) & 1)) return DMVS_ERR_ARG;
  const bool phase_launch = d.explicit_extent != 0 || d.y_row_stride != 0 || d.res_row_stride != 0;
  // the output size must be what the geometry implies (phase launches state it themselves)
  if (!d.explicit_extent &&
      ((d.H + 2 * d.pad_h - d.KH) / d.stride + 1 != d.Ho || (d.W + 2 * d.pad_w - d.KW) / d.stride + 1 != d.Wo ||
       (d.D + 2 * d.pad_d - d.KD) / d.stride + 1 != d.Do))
    return DMVS_ERR_ARG;
  if (!aligned16(d.w)) return DMVS_ERR_ALIGN;
  if (phase_launch) {   // one-sided padding / strided output rows: the TMA-fed tcgen05 kernel only
    if (d.precision != DMVS_PREC_AUTO && d.precision != DMVS_PREC_WS2_TF32X3 && d.precision != DMVS_PREC_WS2_TF32_F16C)
      return DMVS_ERR_UNSUPPORTED;
    if (d.res_up2 || d.pad_d < 0 || d.pad_h < 0 || d.pad_w < 0) return DMVS_ERR_ARG;
    if (!conv_ws2_supported(d)) return DMVS_ERR_UNSUPPORTED;
    return dispatch_conv_ws2(d, static_cast<cudaStream_t>(stream));
  }
  if (d.precision == DMVS_PREC_AUTO) {
    // fp32-class arithmetic, back end chosen per layer without measuring (the host-side autotuner measures instead,
    // ops._tune): the TMA-fed tcgen05 kernel for every layer it supports with at least 0.25 GMAC of work and 8 input
    // channels (measured: profiles/r2_*), the first-generation kernel for nearest-upsampled inputs, FFMA elsewhere
    const double macs = (double)d.KD * d.KH * d.KW * (d.C1 + d.C2) * d.Cout * d.N * d.Do * d.Ho * d.Wo;
    if (macs >= 2.5e8 && d.C1 + d.C2 >= 8) {
      if (conv_ws2_supported(d)) return dispatch_conv_ws2(d, static_cast<cudaStream_t>(stream));
      if (conv_ws_supported(d)) {
        dmvs_conv_desc alt = d;
        alt.precision = DMVS_PREC_WS_TF32X3;
        return dispatch_conv_ws(alt, static_cast<cudaStream_t>(stream));
      }
    }
  } else if (d.precision == DMVS_PREC_WS2_TF32X3 || d.precision == DMVS_PREC_WS2_TF32_F16C) {
    // TMA-fed width-stacked kernel where it applies, the first-generation one for nearest-upsampled inputs, FFMA elsewhere
    if (conv_ws2_supported(d)) return dispatch_conv_ws2(d, static_cast<cudaStream_t>(stream));
    if (conv_ws_supported(d)) {
      dmvs_conv_desc alt = d;
      alt.precision = DMVS_PREC_WS_TF32X3;
      return dispatch_conv_ws(alt, static_cast<cudaStream_t>(stream));
    }
    // fall through to the FFMA kernel
  } else if (d.precision == DMVS_PREC_WS_TF32X3 || d.precision == DMVS_PREC_WS_TF32) {
    // width-stacked tcgen05 kernel wherever its layout preconditions hold (stride 1, 16-byte channel groups);
    // DMVS_WS_MIN_K / DMVS_WS_MIN_MACS bound the layers it takes (tuning aids, see dispatch rule below)
    static const long ws_min_k = getenv("DMVS_WS_MIN_K") ? atol(getenv("DMVS_WS_MIN_K")) : 0;
    static const double ws_min_macs = getenv("DMVS_WS_MIN_MACS") ? atof(getenv("DMVS_WS_MIN_MACS")) : 0.0;
    const long kred = (long)d.KD * d.KH * d.KW * (d.C1 + d.C2);
    const double macs = (double)kred * d.Cout * d.N * d.Do * d.Ho * d.Wo;
    if (conv_ws_supported(d) && kred >= ws_min_k && macs >= ws_min_macs)
      return dispatch_conv_ws(d, static_cast<cudaStream_t>(stream));
#ifdef DMVS_LEGACY_BACKENDS
    if (d.precision == DMVS_PREC_WS_TF32 && d.w_t) {
      dmvs_conv_desc alt = d;
      alt.precision = DMVS_PREC_TF32;
      return dispatch_conv_mma(alt, static_cast<cudaStream_t>(stream));
    }
#endif
    // fall through to the FFMA kernel
  } else if (d.precision != DMVS_PREC_FP32) {
    if (d.precision < DMVS_PREC_FP32 || d.precision > DMVS_PREC_TC_TF32) return DMVS_ERR_ARG;
    if (d.precision >= DMVS_PREC_TC_TF32X3) {
      if (conv_tc_supported(d)) return dispatch_conv_tc(d, static_cast<cudaStream_t>(stream));
      dmvs_conv_desc alt = d;   // strided layers: legacy tensor-core path with the same arithmetic
      alt.precision = d.precision == DMVS_PREC_TC_TF32X3 ? DMVS_PREC_TF32X3 : DMVS_PREC_TF32;
      if (!alt.w_t) return DMVS_ERR_ARG;
      return dispatch_conv_mma(alt, static_cast<cudaStream_t>(stream));
    }
    if (!d.w_t) return DMVS_ERR_ARG;
    return dispatch_conv_mma(d, static_cast<cudaStream_t>(stream));
  }

  ConvArgs a;
  a.d = d;
  a.cin_pad = (d.C1 + d.C2 + 3) & ~3;
  a.w_cstride = (d.Cout + 3) & ~3;
  const bool vec_x = aligned16(d.x) && (d.x_ps % 4 == 0) && (d.C1 % 4 == 0);
  const bool vec_x2 = d.C2 == 0 || (aligned16(d.x2) && (d.x2_ps % 4 == 0) && (d.C2 % 4 == 0));
  a.fast_in = vec_x && vec_x2 && d.in_stats == nullptr;
  a.vec_y = aligned16(d.y) && (d.y_ps % 4 == 0);
  a.vec_res = d.res != nullptr && aligned16(d.res) && (d.res_ps % 4 == 0);
  a.Hs = d.in_up2 ? d.H / 2 : d.H;
  a.Ws = d.in_up2 ? d.W / 2 : d.W;

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int S = d.stride;
  int remaining = a.w_cstride;
  int co_base = 0;
  while (remaining > 0) {
    int chunk = 128;
    while (chunk > remaining) chunk >>= 1;  // largest power of two (>=4) not above what is left
    int co_t, wc;
    // choose rows-per-thread and the channel chunk that fit the shared-memory budget
    const int max_px = chunk >= 32 ? 2 : 4;
    int px = max_px, ck = 0;
    size_t smem = 0;
    const int wcount = chunk == 64 ? 2 : (chunk == 128 ? 4 : 1);
    for (; px >= 1; px >>= 1) {
      const int th = (8 / wcount) * px;
      const int in_rows = (th - 1) * S + d.KH;
      const int in_cols = (kTileW - 1) * S + d.KW;
      ck = 0;
      for (int c = 16; c >= 4; c >>= 1) {
        if (c > a.cin_pad && c > 4) continue;
        const int ckp = c == 4 ? 4 : c + 4;
        size_t need = ((size_t)in_rows * in_cols * ckp + (size_t)d.KH * d.KW * c * chunk + 2 * (size_t)d.C1) * 4;
        const size_t out_tile = (size_t)th * kTileW * (chunk + 4) * 4;  // epilogue staging reuses the buffer
        if (out_tile > need) need = out_tile;
        if (need <= (size_t)kSmemBudget) { ck = c; smem = need; break; }
      }
      if (ck) {
        // keep at least ~2 waves of CTAs when a smaller tile is possible
        const int th2 = (8 / wcount) * px;
        const long blocks = (long)ceil_div(d.Wo, kTileW) * ceil_div(d.Ho, th2) * d.N * d.Do;
        static const long min_blocks = getenv("DMVS_CONV_MINBLOCKS") ? atol(getenv("DMVS_CONV_MINBLOCKS")) : kNumSMs;
        if (blocks >= min_blocks || px == 1) break;   // one full wave is enough: taller per-thread tiles amortise LDS better
      }
    }
    if (!ck) return DMVS_ERR_UNSUPPORTED;
    if (px < 1) px = 1;
    KernelFn fn = pick_kernel(chunk, px, S, &co_t, &wc);
    const int th = (8 / wc) * px;
    a.co_base = co_base;
    a.CK = ck;
    a.CKP = ck == 4 ? 4 : ck + 4;
    a.ck4_shift = ck == 4 ? 0 : (ck == 8 ? 1 : 2);
    a.in_rows = (th - 1) * S + d.KH;
    a.in_cols = (kTileW - 1) * S + d.KW;
    dim3 grid(ceil_div(d.Wo, kTileW), ceil_div(d.Ho, th), d.N * d.Do);
    if (grid.y > 65535 || grid.z > 65535) return DMVS_ERR_UNSUPPORTED;
    launch_pdl(fn, grid, dim3(kThreads), smem, st, a);
    int rc = launch_status();
    if (rc) return rc;
    co_base += chunk;
    remaining -= chunk;
  }
  return 0;
}

extern "C" int dmvs_conv_ws_plan(const dmvs_conv_desc* dp, int32_t* out, int32_t cap) {
  if (dp == nullptr || out == nullptr || cap <= 0) return DMVS_ERR_ARG;
  return plan_conv_ws(*dp, out, cap);
}

extern "C" int dmvs_conv_ws2_plan(const dmvs_conv_desc* dp, int32_t* out, int32_t cap) {
  if (dp == nullptr || out == nullptr || cap <= 0) return DMVS_ERR_ARG;
  return plan_conv_ws2(*dp, out, cap);
}

extern "C" int dmvs_conv_ws2_timeline(int64_t* out, int32_t count) {
  if (out == nullptr || count <= 0) return DMVS_ERR_ARG;
  return read_ws2_debug(reinterpret_cast<long long*>(out), count);
}

extern "C" int dmvs_conv_backends(const dmvs_conv_desc* dp) {
  if (dp == nullptr) return DMVS_ERR_ARG;
  const dmvs_conv_desc& d = *dp;
  if (d.explicit_extent != 0 || d.y_row_stride != 0 || d.res_row_stride != 0)   // phase launches (ops.conv_up2)
    return conv_ws2_supported(d) ? 16 : 0;
  int mask = 1;                                        // bit 0: FFMA kernel (always)
#ifdef DMVS_LEGACY_BACKENDS
  if (d.w_t) mask |= 2;                                // bit 1: legacy mma.sync kernel
  if (d.w_tc && conv_tc_supported(d)) mask |= 4;       // bit 2: tcgen05 kernel, taps as descriptor offsets
#endif
  if (conv_ws_supported(d)) mask |= 8;       // bit 3: tcgen05 kernel, kernel-row taps stacked along N
  if (conv_ws2_supported(d)) mask |= 16;     // bit 4: the same arithmetic behind the TMA-fed pipeline (conv_ws2.cu)
  if ((mask & 16) && d.w_ws16 != nullptr && !(d.w_ws_pair != nullptr && d.C1 <= 4 && d.C2 == 0 && d.stride == 1 && d.KH >= 2))
    mask |= 32;                              // bit 5: ... with the fp16 correction MMA (DMVS_PREC_WS2_TF32_F16C)
  return mask;
}

// ---------------------------------------------------------------------------------------------
// ConvTranspose3d(k3, s2, p1, op1) + bias + ReLU + skip    (module.py:110-144, 436-437, 445-446)
// out[o] = sum_k in[(o + 1 - k)/2] * w[k] over taps with (o + 1 - k) even and in range.
// One thread = one output voxel x all Cout (<= 16).  Tiny FLOP share (< 2 GFLOP), so kept simple.
// ---------------------------------------------------------------------------------------------
namespace dmvs {
namespace {

template <int COUT>
__global__ void __launch_bounds__(128) deconv3d_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, const float* __restrict__ skip,
                                                       float* __restrict__ y, int N, int D, int H, int W, int Cin) {
  pdl_sync();
  extern __shared__ __align__(16) float ws[];  // [27][Cin][COUT]
  const int wtotal = 27 * Cin * COUT;
  for (int i = threadIdx.x; i < wtotal; i += blockDim.x) ws[i] = __ldg(w + i);
  __syncthreads();
  const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
  const int64_t total = (int64_t)N * Do * Ho * Wo;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(o % Wo);
    int64_t t = o / Wo;
    const int oy = (int)(t % Ho);
    t /= Ho;
    const int oz = (int)(t % Do);
    const int n = (int)(t / Do);
    float acc[COUT];
#pragma unroll
    for (int j = 0; j < COUT; ++j) acc[j] = 0.f;
    for (int kd = 0; kd < 3; ++kd) {
      const int tz = oz + 1 - kd;
      if (tz < 0 || (tz & 1) || (tz >> 1) >= D) continue;
      for (int kh = 0; kh < 3; ++kh) {
        const int ty = oy + 1 - kh;
        if (ty < 0 || (ty & 1) || (ty >> 1) >= H) continue;
        for (int kw = 0; kw < 3; ++kw) {
          const int tx = ox + 1 - kw;
          if (tx < 0 || (tx & 1) || (tx >> 1) >= W) continue;
          const float* xp = x + ((((int64_t)n * D + (tz >> 1)) * H + (ty >> 1)) * W + (tx >> 1)) * Cin;
          const float* wp = ws + ((kd * 3 + kh) * 3 + kw) * Cin * COUT;
          for (int c4 = 0; c4 < Cin; c4 += 4) {
            const float4 v = ldg4(xp + c4);
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
              for (int j = 0; j < COUT; ++j) acc[j] = fmaf(e[k], wp[(c4 + k) * COUT + j], acc[j]);
            }
          }
        }
      }
    }
    float* yp = y + o * COUT;
    const float* sp = skip + o * COUT;
#pragma unroll
    for (int j4 = 0; j4 < COUT / 4; ++j4) {
      const float4 s = ldg4(sp + j4 * 4);
      float4 r;
      r.x = fmaxf(acc[j4 * 4 + 0] + __ldg(bias + j4 * 4 + 0), 0.f) + s.x;
      r.y = fmaxf(acc[j4 * 4 + 1] + __ldg(bias + j4 * 4 + 1), 0.f) + s.y;
      r.z = fmaxf(acc[j4 * 4 + 2] + __ldg(bias + j4 * 4 + 2), 0.f) + s.z;
      r.w = fmaxf(acc[j4 * 4 + 3] + __ldg(bias + j4 * 4 + 3), 0.f) + s.w;
      *reinterpret_cast<float4*>(yp + j4 * 4) = r;
    }
  }
}

// Parity form of the same operator for the two shapes CostRegNet_small uses (32 -> 16 and 16 -> 8 channels).
// An output voxel (2z+pz, 2y+py, 2x+px) reads kernel tap 1 of input index i along a dimension of even parity and taps
// 0 (input i+1) and 2 (input i) along a dimension of odd parity: 1, 2, 4 or 8 taps depending on the parity class.  A
// warp owns one parity class of 32 consecutive input columns, so its tap list is warp-uniform (no divergence), the
// weights are 128-bit shared-memory broadcasts feeding four FFMAs each, and the four warps of a CTA take class lists of
// 8, 7, 6 and 6 taps: together they write the complete 2 x 2 x 2 output octets of their 32 input voxels.  Persistent
// CTAs load the 27 x Cin x Cout weights once.  Summation order = the gather kernel above (results are bit-identical).
template <int CIN, int COUT>
__global__ void __launch_bounds__(128) deconv3d_parity_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ bias,
                                                              const float* __restrict__ skip, float* __restrict__ y,
                                                              int N, int D, int H, int W) {
  extern __shared__ __align__(16) float ws[];  // [27][CIN][COUT]
  // the weights never change after packing: staged before the dependency wait, overlapping the producer's tail
  for (int i = threadIdx.x * 4; i < 27 * CIN * COUT; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(ws + i) = ldg4(w + i);
  pdl_sync();
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // parity classes (bit 2: depth, bit 1: row, bit 0: column) per warp, 0xF terminated: 8 | 4+2+1 | 4+2 | 4+2 taps
  const unsigned my_list = warp == 0 ? 0xFFF7u : warp == 1 ? 0xF046u : warp == 2 ? 0xFF25u : 0xFF13u;
  const int xchunks = ceil_div(W, 32);
  const int64_t items = (int64_t)N * D * H * xchunks;
  const int Ho = 2 * H, Wo = 2 * W;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int xc = (int)(item % xchunks);
    int64_t t = item / xchunks;
    const int iy = (int)(t % H);
    t /= H;
    const int iz = (int)(t % D);
    const int n = (int)(t / D);
    const int ix = xc * 32 + lane;
    for (unsigned list = my_list; (list & 0xFu) != 0xFu; list >>= 4) {
      const int pz = (list >> 2) & 1, py = (list >> 1) & 1, px = list & 1;
      float acc[COUT];
#pragma unroll
      for (int j = 0; j < COUT; ++j) acc[j] = 0.f;
      for (int a = 0; a <= pz; ++a) {                       // odd parity: tap 0 reads input i+1, then tap 2 reads input i
        const int kd = pz ? 2 * a : 1, sz = iz + (pz && a == 0 ? 1 : 0);
        if (sz >= D) continue;
        for (int b = 0; b <= py; ++b) {
          const int kh = py ? 2 * b : 1, sy = iy + (py && b == 0 ? 1 : 0);
          if (sy >= H) continue;
          for (int c = 0; c <= px; ++c) {
            const int kw = px ? 2 * c : 1, sx = ix + (px && c == 0 ? 1 : 0);
            const bool ok = sx < W;
            const float* xp = x + ((((int64_t)n * D + sz) * H + sy) * W + (ok ? sx : 0)) * CIN;
            const float* wp = ws + ((kd * 3 + kh) * 3 + kw) * CIN * COUT;
#pragma unroll 2
            for (int c4 = 0; c4 < CIN; c4 += 4) {
              float4 v = ldg4(xp + c4);
              if (!ok) v = make_float4(0.f, 0.f, 0.f, 0.f);
              const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int j4 = 0; j4 < COUT; j4 += 4) {
                  const float4 w4 = *reinterpret_cast<const float4*>(wp + (c4 + k) * COUT + j4);
                  acc[j4 + 0] = fmaf(e[k], w4.x, acc[j4 + 0]);
                  acc[j4 + 1] = fmaf(e[k], w4.y, acc[j4 + 1]);
                  acc[j4 + 2] = fmaf(e[k], w4.z, acc[j4 + 2]);
                  acc[j4 + 3] = fmaf(e[k], w4.w, acc[j4 + 3]);
                }
              }
            }
          }
        }
      }
      if (ix < W) {
        const int64_t o = (((int64_t)n * 2 * D + 2 * iz + pz) * Ho + 2 * iy + py) * Wo + 2 * ix + px;
        float* yp = y + o * COUT;
        const float* sp = skip + o * COUT;
#pragma unroll
        for (int j4 = 0; j4 < COUT; j4 += 4) {
          const float4 s = ldg4(sp + j4), b4 = ldg4(bias + j4);
          float4 r;
          r.x = fmaxf(acc[j4 + 0] + b4.x, 0.f) + s.x;
          r.y = fmaxf(acc[j4 + 1] + b4.y, 0.f) + s.y;
          r.z = fmaxf(acc[j4 + 2] + b4.z, 0.f) + s.z;
          r.w = fmaxf(acc[j4 + 3] + b4.w, 0.f) + s.w;
          *reinterpret_cast<float4*>(yp + j4) = r;
        }
      }
    }
  }
}

template <int CIN, int COUT>
int launch_deconv3d_parity(const float* x, const float* w, const float* bias, const float* skip, float* y, int N, int D,
                           int H, int W, cudaStream_t st) {
  static SmemOptIn opt_in;
  const size_t smem = (size_t)27 * CIN * COUT * 4;
  opt_in.ensure(deconv3d_parity_kernel<CIN, COUT>, 100 * 1024);
  const int64_t items = (int64_t)N * D * H * ceil_div(W, 32);
  const int per_sm = smem > 40 * 1024 ? 3 : 8;
  const int blocks = (int)(items < (int64_t)kNumSMs * per_sm ? items : (int64_t)kNumSMs * per_sm);
  launch_pdl(deconv3d_parity_kernel<CIN, COUT>, dim3(blocks), dim3(128), smem, st, x, w, bias, skip, y, N, D, H, W);
  return launch_status();
}

}  // namespace
}  // namespace dmvs

extern "C" int dmvs_deconv3d_f32(const float* x, const float* w, const float* bias, const float* skip, float* y,
                                 int32_t N, int32_t D, int32_t H, int32_t W, int32_t Cin, int32_t Cout, void* stream) {
  if (!x || !w || !bias || !skip || !y) return DMVS_ERR_ARG;
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0 || Cin <= 0 || (Cin % 4) != 0) return DMVS_ERR_ARG;
  if (Cout != 8 && Cout != 16) return DMVS_ERR_UNSUPPORTED;
  if (!aligned16(x) || !aligned16(skip) || !aligned16(y)) return DMVS_ERR_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (aligned16(w) && aligned16(bias)) {
    if (Cin == 32 && Cout == 16) return launch_deconv3d_parity<32, 16>(x, w, bias, skip, y, N, D, H, W, st);
    if (Cin == 16 && Cout == 8) return launch_deconv3d_parity<16, 8>(x, w, bias, skip, y, N, D, H, W, st);
  }
  const size_t smem = (size_t)27 * Cin * Cout * 4;
  const int64_t total = (int64_t)N * 8 * D * H * W;
  int blocks = (int)(ceil_div64(total, 128) < (int64_t)kNumSMs * 16 ? ceil_div64(total, 128) : (int64_t)kNumSMs * 16);
  if (Cout == 8) {
    static SmemOptIn opt_in;
    opt_in.ensure(deconv3d_kernel<8>, 100 * 1024);
    launch_pdl(deconv3d_kernel<8>, dim3(blocks), dim3(128), smem, st, x, w, bias, skip, y, N, D, H, W, Cin);
  } else {
    static SmemOptIn opt_in;
    opt_in.ensure(deconv3d_kernel<16>, 100 * 1024);
    launch_pdl(deconv3d_kernel<16>, dim3(blocks), dim3(128), smem, st, x, w, bias, skip, y, N, D, H, W, Cin);
  }
  return launch_status();
}
