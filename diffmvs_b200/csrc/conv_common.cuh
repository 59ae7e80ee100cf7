// Pieces shared by the two convolution back ends (conv.cu: fp32 FFMA; conv_mma.cu: tensor-core
// implicit GEMM): the staged input tile, the weight-independent prologue (GroupNorm+SiLU on load,
// nearest x2 upsampling, virtual concat) and the fused epilogue.
#pragma once
#include "common.cuh"

namespace dmvs {

constexpr int kConvThreads = 256;
constexpr int kTileW = 32;               // output tile width (one lane / MMA row per pixel)
constexpr int kSmemBudget = 100 * 1024;  // two CTAs per SM

struct ConvArgs {
  dmvs_conv_desc d;
  int cin_pad;     // (C1 + C2) rounded up to 4 (FFMA) or 8 (MMA)
  int w_cstride;   // FFMA: total padded Cout of the packed weights; MMA: padded Cin of the transposed weights
  int co_base;     // first output channel of this launch
  int CK;          // input channels staged per chunk (4, 8, 16)
  int CKP;         // padded pixel pitch of the staged tile in floats
  int ck4_shift;   // log2(CK / 4)
  int in_rows, in_cols;
  int fast_in;     // every staged 4-channel unit is one aligned 16-byte segment and needs no transform
  int vec_y;       // 128-bit stores allowed on y
  int vec_res;     // 128-bit loads allowed on the residual
  int Hs, Ws;      // stored size of x (H/2, W/2 when in_up2)
  int passes;      // MMA only: 1 = plain TF32, 3 = 3xTF32 split (fp32-class accuracy)
};

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool valid) {
  const unsigned saddr = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  const int bytes = valid ? 16 : 0;  // src-size 0 -> the 16 bytes are zero-filled (conv zero padding)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(saddr), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// One output value through the fused epilogue (kept out of line: it is cold relative to the math loop and
// inlining it four times per quad blows the kernel past the instruction cache).
static __device__ __noinline__ float epilogue_value(const dmvs_conv_desc& d, float x, int c, int64_t opix, int64_t rpix) {
  if (d.epi == DMVS_EPI_STD) {
    if (d.res_mode == DMVS_RES_PRE_ACT) x += __ldg(d.res + rpix * d.res_ps + c);
    if (c >= d.act_c0) x = apply_act(x, d.act);
    if (d.res_mode == DMVS_RES_POST_ACT) x += __ldg(d.res + rpix * d.res_ps + c);
  } else if (d.epi == DMVS_EPI_GRU_ZR) {
    x = sigmoidf_(x);
    if (c >= d.gru_hidden) x *= __ldg(d.aux1 + opix * d.aux1_ps + (c - d.gru_hidden));
  } else {  // DMVS_EPI_GRU_Q
    const float z = __ldg(d.aux1 + opix * d.aux1_ps + c);
    const float h = __ldg(d.aux2 + opix * d.aux2_ps + c);
    x = (1.0f - z) * h + z * tanhf(x);
  }
  return x;
}

// GroupNorm(4)+affine of the producer folded to a per-channel (scale, shift) pair for sample n.
static __device__ __noinline__ void groupnorm_affine(const dmvs_conv_desc& d, int n, int c, float* gn_s) {
  const int g = c / (d.C1 / 4);
  const double s = stat_value(d.in_stats[(n * 4 + g) * 2 + 0]);
  const double q = stat_value(d.in_stats[(n * 4 + g) * 2 + 1]);
  const double mean = s * (double)d.in_inv_count;
  double var = q * (double)d.in_inv_count - mean * mean;
  var = var < 0.0 ? 0.0 : var;
  const float rstd = (float)(1.0 / sqrt(var + 1e-5));
  const float g1 = d.in_g1[c] * rstd;
  gn_s[c] = g1;
  gn_s[d.C1 + c] = d.in_g0[c] - (float)mean * g1;
}

static __device__ __noinline__ float staged_silu(float v, float scale, float shift) { return siluf_(fmaf(v, scale, shift)); }

// Stage input channels [c0, c0+CK) of depth slice `id` for the tile whose top-left input pixel is (iy0, ix0).
// Warps take rows, lanes take (column, channel-quad) units.  Issues cp.async on the fast path; the caller
// waits (cp_async_wait_all) and synchronises.
__device__ __forceinline__ void stage_input_tile(const ConvArgs& a, float* in_s, const float* gn_s, int n, int id, int iy0,
                                                 int ix0, int c0) {
  const dmvs_conv_desc& d = a.d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ck4 = a.CK >> 2;
  const int units_per_row = a.in_cols << a.ck4_shift;
  const int Ctot = d.C1 + d.C2;
#pragma unroll 1
  for (int row = warp; row < a.in_rows; row += kConvThreads / 32) {
    const int iy = iy0 + row;
    const bool row_ok = iy >= 0 && iy < d.H;
    const int sy = d.in_up2 ? (iy >> 1) : iy;
    const int64_t row_pix = ((int64_t)(n * d.D + id) * a.Hs + sy) * a.Ws;
    float* row_dst = in_s + row * a.in_cols * a.CKP;
#pragma unroll 1
    for (int u = lane; u < units_per_row; u += 32) {
      const int c4 = u & (ck4 - 1);
      const int col = u >> a.ck4_shift;
      const int ix = ix0 + col;
      const int ch = c0 + c4 * 4;
      const bool ok = row_ok && ix >= 0 && ix < d.W && ch < Ctot;
      const int sx = d.in_up2 ? (ix >> 1) : ix;
      const int64_t pix = row_pix + sx;
      float* dst = row_dst + col * a.CKP + c4 * 4;
      if (a.fast_in) {
        const float* src = d.x;
        if (ok) src = ch < d.C1 ? d.x + pix * d.x_ps + ch : d.x2 + pix * d.x2_ps + (ch - d.C1);
        cp_async16(dst, src, ok);
      } else {
        float e[4] = {0.f, 0.f, 0.f, 0.f};
        if (ok) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int c = ch + k;
            if (c < d.C1) {
              float v = __ldg(d.x + pix * d.x_ps + c);
              if (d.in_stats != nullptr) v = staged_silu(v, gn_s[c], gn_s[d.C1 + c]);
              e[k] = v;
            } else if (c < Ctot) {
              e[k] = __ldg(d.x2 + pix * d.x2_ps + (c - d.C1));
            }
          }
        }
        *reinterpret_cast<float4*>(dst) = make_float4(e[0], e[1], e[2], e[3]);
      }
    }
  }
}

// Fused epilogue over an output tile staged in shared memory as out_s[TH*32][COUT_S+4]: bias, residual,
// activation / GRU blends, GroupNorm statistics, fully coalesced 128-bit stores.  All threads must call it.
template <int TH, int COUT_S>
__device__ __forceinline__ void epilogue_tile(const ConvArgs& a, const float* out_s, unsigned long long* stat_s, int n, int od, int ty0,
                                              int tx0) {
  constexpr int N4 = COUT_S / 4;
  constexpr int OP = COUT_S + 4;
  const dmvs_conv_desc& d = a.d;
  const int tid = threadIdx.x, lane = tid & 31;
  const int q4 = tid % N4;                 // this thread always handles the same channel quad
  const int cq = a.co_base + q4 * 4;       // its first absolute output channel
  float gs[4] = {0.f, 0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f};
  if (cq < d.Cout) {
    float bias[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (d.bias != nullptr && cq + k < d.Cout) bias[k] = __ldg(d.bias + cq + k);
    const bool full_quad = cq + 4 <= d.Cout;
    // Plain epilogues (bias + optional ReLU: > 80 % of all launches) take a branch-free inline path; residuals,
    // sigmoid/tanh/SiLU and the GRU blends go through the out-of-line generic routine.
    const bool plain = d.epi == DMVS_EPI_STD && (d.act == DMVS_ACT_NONE || d.act == DMVS_ACT_RELU);
    const int relu_from = d.act == DMVS_ACT_RELU ? d.act_c0 : 0x7fffffff;
    const int64_t img_base = (int64_t)(n * d.Do + od) * d.Ho;
#pragma unroll 1
    for (int pix = tid / N4; pix < TH * kTileW; pix += kConvThreads / N4) {
      const int oy = ty0 + (pix >> 5), ox = tx0 + (pix & 31);
      if (oy >= d.Ho || ox >= d.Wo) continue;
      const float4 t4 = *reinterpret_cast<const float4*>(out_s + pix * OP + q4 * 4);
      float v[4] = {t4.x + bias[0], t4.y + bias[1], t4.z + bias[2], t4.w + bias[3]};
      const int64_t opix = (img_base + oy) * d.Wo + ox;
      int64_t rpix = opix;
      if (d.res_up2) rpix = ((int64_t)n * (d.Ho >> 1) + (oy >> 1)) * (d.Wo >> 1) + (ox >> 1);
      if (plain) {   // bias (+ residual before / after) + optional ReLU, inline
        float r[4] = {0.f, 0.f, 0.f, 0.f};
        if (d.res_mode != DMVS_RES_NONE) {
          const float* rp = d.res + rpix * d.res_ps + cq;
          if (a.vec_res && full_quad) {
            const float4 r4 = ldg4(rp);
            r[0] = r4.x; r[1] = r4.y; r[2] = r4.z; r[3] = r4.w;
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (cq + k < d.Cout) r[k] = __ldg(rp + k);
          }
        }
        const bool pre = d.res_mode == DMVS_RES_PRE_ACT;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float x = pre ? v[k] + r[k] : v[k];
          if (cq + k >= relu_from) x = fmaxf(x, 0.0f);
          v[k] = pre ? x : x + r[k];
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (cq + k < d.Cout) v[k] = epilogue_value(d, v[k], cq + k, opix, rpix);
      }
      if (d.out_stats != nullptr) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          gs[k] += v[k];
          gq[k] += v[k] * v[k];
        }
      }
      float* yp = d.y + opix * d.y_ps + cq;
      if (a.vec_y && full_quad) {
        *reinterpret_cast<float4*>(yp) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (cq + k < d.Cout) yp[k] = v[k];
      }
    }
  }
  if (d.out_stats != nullptr) {
    // lanes l, l+N4, l+2*N4, ... of a warp share a channel quad: fold them, then one shared integer atomic per
    // (quad, element), then one global integer atomic per (group, moment) per CTA (fixed point: order independent).
    const int cpg = d.Cout / 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float s = gs[k], q = gq[k];
#pragma unroll
      for (int o = 16; o >= (N4 < 32 ? N4 : 32); o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      const int c = cq + k;
      if (lane < N4 && c < d.Cout) {
        const int g = c / cpg;
        atomicAdd(&stat_s[g * 2 + 0], stat_fixed(s));
        atomicAdd(&stat_s[g * 2 + 1], stat_fixed(q));
      }
    }
    __syncthreads();
    if (tid < 8) atomicAdd(reinterpret_cast<unsigned long long*>(d.out_stats) + n * 8 + tid, stat_s[tid]);
  }
}

// tensor-core back end (conv_mma.cu); returns 0 / DMVS_ERR_* / cudaError
int dispatch_conv_mma(const dmvs_conv_desc& d, cudaStream_t stream);
// tcgen05 / TMEM back end (conv_tc.cu)
bool conv_tc_supported(const dmvs_conv_desc& d);
int dispatch_conv_tc(const dmvs_conv_desc& d, cudaStream_t stream);
// width-stacked tcgen05 back end (conv_ws.cu): the KW taps of a kernel row share one MMA (N = KW * Cout)
bool conv_ws_supported(const dmvs_conv_desc& d);
int dispatch_conv_ws(const dmvs_conv_desc& d, cudaStream_t stream);
int plan_conv_ws(const dmvs_conv_desc& d, int32_t* out, int cap);   // tile plan only: {CC,N,TH,TW,n_blk,R,ctas,smem} per launch

// second-generation width-stacked back end (conv_ws2.cu): TMA-fed ring, dedicated split / MMA / epilogue warps, two
// TMEM accumulator sets
bool conv_ws2_supported(const dmvs_conv_desc& d);
int dispatch_conv_ws2(const dmvs_conv_desc& d, cudaStream_t stream);
int plan_conv_ws2(const dmvs_conv_desc& d, int32_t* out, int cap);
int read_ws2_debug(long long* host_out, int count);   // DMVS_WS2_DBG=1: clock stamps of CTA 0 of the last ws2 launch

}  // namespace dmvs
