// tcgen05 / TMEM implicit-GEMM convolution (stride 1, 2-D / 3-D, channels-last) for sm_100a.
//
// GEMM view per CTA: M = the output pixels of a TH x TW tile, N = a chunk of 16/32/64 output channels,
// K = taps x input channels, D accumulates in TMEM (fp32).
//
//   * The input halo tile is staged *planar by channel quad*: A_s[q][p][4 floats], p = row * in_cols + col the
//     flattened pixel index of the (TH+KH-1) x (TW+KW-1) tile.  8 consecutive pixels x 16 bytes are then one
//     K-major, un-swizzled UMMA core matrix (SBO = 128 B between 8-pixel groups, LBO = plane pitch between the
//     two channel quads of a K=8 step), so an M=128 operand is simply 128 consecutive flattened pixels.
//   * A convolution tap (kh,kw) is a *descriptor offset* of (kh*in_cols + kw)*16 bytes - no im2col, no
//     re-staging, the tile is read by the tensor core once per tap straight from shared memory
//     (validated stand-alone in tools/probes/tc_probe.cu).  Flattened positions that fall into halo columns
//     produce values that are never stored.
//   * 3xTF32: each staged element is split ONCE per stage into hi = rna_tf32(x), lo = rna_tf32(x - hi) planes
//     (the legacy mma.sync path re-split per tap and spent >90 % of its instructions there); weights are
//     pre-split on the host.  D += Alo*Bhi + Ahi*Blo + Ahi*Bhi, all issued by one thread.
//   * One CTA = 128 threads = 4 warps = the 128 TMEM lanes; two CTAs per SM overlap one CTA's loads / split /
//     epilogue with the other's MMAs.  Accumulators leave TMEM through tcgen05.ld into a small shared staging
//     buffer and the common fused epilogue (bias, residual, activation, GRU blends, GroupNorm statistics,
//     coalesced 128-bit stores).
#include <cstdlib>

#include "conv_common.cuh"

namespace dmvs {
namespace {

constexpr int kTcThreads = 256;   // warps 0-3 also own the 128 TMEM lanes in the epilogue

struct TcArgs {
  dmvs_conv_desc d;
  int cin_pad;      // (C1+C2) rounded up to 8
  int cout_pad;     // Cout rounded up to 16 (pitch of the packed weights)
  int co_base;      // first output channel of this launch
  int CK;           // channels per stage (8 or 16)
  int TH, TW;       // output tile
  int in_rows, in_cols;
  int plane;        // pixels per channel-quad plane (incl. slack for the last M block)
  int n_blk;        // number of M=128 blocks
  int tmem_cols;    // allocated TMEM columns (power of two >= 32)
  int tiles_x, tiles_y;
  int fast_in, vec_y;
  int Hs, Ws;
  int64_t w_lo_off; // offset (floats) of the lo weights inside w_tc
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t v = 0;
  v |= (uint64_t)((saddr >> 4) & 0x3fff);
  v |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  v |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  v |= 1ull << 46;  // descriptor version (Blackwell); layout_type 0 = no swizzle, K-major
  return v;
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    if (!done && ++spins > (1u << 24)) __trap();   // watchdog: a lost commit must not hang the GPU
  }
}

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// N = output channels per CTA (16, 32, 64), PASSES = 1 (TF32) or 3 (3xTF32)
template <int N, int PASSES>
__global__ void __launch_bounds__(kTcThreads, 1) conv_tc_kernel(const __grid_constant__ TcArgs a) {
  constexpr int OP = N + 4;   // pitch of the epilogue staging rows
  const dmvs_conv_desc& d = a.d;
  extern __shared__ __align__(128) float smem[];
  const int quads = a.CK >> 2;
  const int taps = d.KH * d.KW;
  const int plane_f = quads * a.plane * 4;                    // floats per operand plane set
  const int wslab_f = taps * quads * N * 4;                   // floats per weight slab
  float* a_raw = smem;                                        // [quads][plane][4]  cp.async landing buffer
  float* a_hi = a_raw + plane_f;                              // split operands read by the tensor core
  float* a_lo = a_hi + plane_f;
  float* w_hi0 = a_lo + (PASSES == 3 ? plane_f : 0);          // [2 buffers][taps][quads][N][4]
  float* w_lo0 = w_hi0 + 2 * wslab_f;
  float* gn_s = w_lo0 + (PASSES == 3 ? 2 * wslab_f : 0);      // [2][C1] when in_stats
  float* out_s = a_raw;                                       // [128][OP] epilogue staging (operands are dead by then)
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t mbar;
  __shared__ float stat_s[8];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.z / d.Do;
  const int od = blockIdx.z - n * d.Do;
  const int ty0 = blockIdx.y * a.TH, tx0 = blockIdx.x * a.TW;
  const int iy0 = ty0 - d.pad_h, ix0 = tx0 - d.pad_w;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(a.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
  }
  if (tid < 8) stat_s[tid] = 0.0f;
  if (d.in_stats != nullptr)
    for (int c = tid; c < d.C1; c += kTcThreads) groupnorm_affine(d, n, c, gn_s);
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  const int Ctot = d.C1 + d.C2;
  const int units_per_row = a.in_cols * quads;   // 16-byte units per tile row

  // stage list: every valid depth tap kd x every channel chunk c0
  const int kd_lo = (d.pad_d - od) > 0 ? (d.pad_d - od) : 0;
  const int kd_hi = (d.D - 1 + d.pad_d - od) < (d.KD - 1) ? (d.D - 1 + d.pad_d - od) : (d.KD - 1);
  const int nchunks = ceil_div(a.cin_pad, a.CK);
  const int nstages = (kd_hi - kd_lo + 1) * nchunks;

  // loads of one stage: halo tile (planar by channel quad) -> a_raw, weights -> w_*[buf]
  auto issue_loads = [&](int s) {
    const int kd = kd_lo + s / nchunks;
    const int c0 = (s % nchunks) * a.CK;
    const int id = od + kd - d.pad_d;
#pragma unroll 1
    for (int row = warp; row < a.in_rows; row += kTcThreads / 32) {
      const int iy = iy0 + row;
      const bool row_ok = iy >= 0 && iy < d.H;
      const int sy = d.in_up2 ? (iy >> 1) : iy;
      const int64_t row_pix = ((int64_t)(n * d.D + id) * a.Hs + sy) * a.Ws;
#pragma unroll 1
      for (int u = lane; u < units_per_row; u += 32) {
        const int q = u % quads;
        const int col = u / quads;
        const int ix = ix0 + col;
        const int ch = c0 + q * 4;
        const bool ok = row_ok && ix >= 0 && ix < d.W && ch < Ctot;
        const int sx = d.in_up2 ? (ix >> 1) : ix;
        const int64_t pix = row_pix + sx;
        const int off = (q * a.plane + row * a.in_cols + col) * 4;
        if (a.fast_in) {
          const float* src = d.x;
          if (ok) src = ch < d.C1 ? d.x + pix * d.x_ps + ch : d.x2 + pix * d.x2_ps + (ch - d.C1);
          cp_async16(a_raw + off, src, ok);
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
          *reinterpret_cast<float4*>(a_raw + off) = make_float4(e[0], e[1], e[2], e[3]);
        }
      }
    }
    // weights of this (kd, channel chunk): global [kd][tap][quad][cout_pad][4] -> [tap][quad][N][4]
    const int q0 = c0 >> 2;
    const int qtot = a.cin_pad >> 2;
    float* wh = w_hi0 + (s & 1) * wslab_f;
    float* wl = w_lo0 + (s & 1) * wslab_f;
#pragma unroll 1
    for (int idx = tid; idx < taps * quads * N; idx += kTcThreads) {
      const int nn = idx % N;
      const int r = idx / N;
      const int q = r % quads;
      const int tap = r / quads;
      const bool ok = q0 + q < qtot;
      const int64_t off = ((((int64_t)kd * taps + tap) * qtot + q0 + q) * a.cout_pad + a.co_base + nn) * 4;
      cp_async16(wh + idx * 4, ok ? d.w_tc + off : d.w_tc, ok);
      if (PASSES == 3) cp_async16(wl + idx * 4, ok ? d.w_tc + a.w_lo_off + off : d.w_tc, ok);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };

  uint32_t parity = 0;
  issue_loads(0);
  for (int s = 0; s < nstages; ++s) {
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();                                   // a_raw and w[s&1] of stage s are visible to every thread
    if (s > 0) {                                       // MMAs of stage s-1 retired: a_hi / a_lo / w[(s+1)&1] are free
      mbar_wait(&mbar, parity);
      parity ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    }
    // ---- split once per stage: raw -> (hi, lo) operand planes (slack included: harmless) ----------------
    {
      const int total = quads * a.plane;
#pragma unroll 1
      for (int u = tid; u < total; u += kTcThreads) {
        const float4 v = *reinterpret_cast<const float4*>(a_raw + u * 4);
        if (PASSES == 3) {
          float4 h, l;
          h.x = rna_tf32(v.x); l.x = rna_tf32(v.x - h.x);
          h.y = rna_tf32(v.y); l.y = rna_tf32(v.y - h.y);
          h.z = rna_tf32(v.z); l.z = rna_tf32(v.z - h.z);
          h.w = rna_tf32(v.w); l.w = rna_tf32(v.w - h.w);
          *reinterpret_cast<float4*>(a_hi + u * 4) = h;
          *reinterpret_cast<float4*>(a_lo + u * 4) = l;
        } else {
          *reinterpret_cast<float4*>(a_hi + u * 4) = v;
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();                                   // operands complete; a_raw may be refilled
    // ---- one thread issues every MMA of the stage; the next stage's loads go out meanwhile ----------------
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      // Descriptors differ only in their 14-bit start-address field (units of 16 bytes), so they are formed
      // once per stage and advanced with integer adds; the loop below is ~5 instructions per MMA.
      const uint32_t lbo_a = (uint32_t)a.plane * 16u, lbo_b = (uint32_t)N * 16u;
      const uint64_t dah0 = umma_desc(smem_u32(a_hi), lbo_a, 128), dal0 = umma_desc(smem_u32(a_lo), lbo_a, 128);
      const uint64_t dbh0 = umma_desc(smem_u32(w_hi0 + (s & 1) * wslab_f), lbo_b, 128);
      const uint64_t dbl0 = umma_desc(smem_u32(w_lo0 + (s & 1) * wslab_f), lbo_b, 128);
      const int ksteps = a.CK >> 3;
      const uint32_t a_kstep = 2u * (uint32_t)a.plane;        // two channel-quad planes per K=8 step (16-byte units)
      const uint32_t b_kstep = 2u * (uint32_t)N;
      for (int blk = 0; blk < a.n_blk; ++blk) {
        const uint32_t d_tmem = tmem_base + (uint32_t)(blk * N);
        uint32_t acc = s == 0 ? 0u : 1u;
        uint32_t b_off = 0;                                     // advances by one (tap, kstep) at a time
        for (int kh = 0; kh < d.KH; ++kh) {
          uint32_t a_off = (uint32_t)(blk * 128 + kh * a.in_cols);
          for (int kw = 0; kw < d.KW; ++kw, ++a_off) {
            uint32_t a_k = a_off;
            for (int ks = 0; ks < ksteps; ++ks, a_k += a_kstep, b_off += b_kstep) {
              if (PASSES == 3) {
                umma_tf32(d_tmem, dal0 + a_k, dbh0 + b_off, idesc, acc);
                umma_tf32(d_tmem, dah0 + a_k, dbl0 + b_off, idesc, 1u);
                umma_tf32(d_tmem, dah0 + a_k, dbh0 + b_off, idesc, 1u);
              } else {
                umma_tf32(d_tmem, dah0 + a_k, dbh0 + b_off, idesc, acc);
              }
              acc = 1u;
            }
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar))
                   : "memory");
    }
    if (s + 1 < nstages) issue_loads(s + 1);
  }
  // ---- all MMAs retired -> accumulators out of TMEM, block by block -----------------------------------------
  mbar_wait(&mbar, parity);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

  constexpr int N4 = N / 4;
  const int q4 = tid % N4;                 // fixed channel quad per thread in the write-out loop
  const int cq = a.co_base + q4 * 4;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (d.bias != nullptr && cq + k < d.Cout) bias[k] = __ldg(d.bias + cq + k);
  const bool full_quad = cq + 4 <= d.Cout;
  const bool plain = d.epi == DMVS_EPI_STD && d.res_mode == DMVS_RES_NONE && (d.act == DMVS_ACT_NONE || d.act == DMVS_ACT_RELU);
  const int relu_from = d.act == DMVS_ACT_RELU ? d.act_c0 : 0x7fffffff;
  const int64_t img_base = (int64_t)(n * d.Do + od) * d.Ho;
  float gs[4] = {0.f, 0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f};

  for (int blk = 0; blk < a.n_blk; ++blk) {
    // TMEM lane = flattened position within the block; warp w < 4 owns lanes [32w, 32w+32)
    float* orow = out_s + tid * OP;
    if (warp < 4) {
#pragma unroll
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(blk * N + c0);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4)
        *reinterpret_cast<float4*>(orow + c0 + j4 * 4) =
            make_float4(__uint_as_float(r[j4 * 4]), __uint_as_float(r[j4 * 4 + 1]), __uint_as_float(r[j4 * 4 + 2]),
                        __uint_as_float(r[j4 * 4 + 3]));
    }
    }
    __syncthreads();
    // cooperative, coalesced write-out of the 128 positions of this block
    if (cq < d.Cout) {
#pragma unroll 1
      for (int m = tid / N4; m < 128; m += kTcThreads / N4) {
        const int p = blk * 128 + m;
        const int py = p / a.in_cols, px = p - py * a.in_cols;
        const int oy = ty0 + py, ox = tx0 + px;
        if (px >= a.TW || py >= a.TH || oy >= d.Ho || ox >= d.Wo) continue;
        const float4 t4 = *reinterpret_cast<const float4*>(out_s + m * OP + q4 * 4);
        float v[4] = {t4.x + bias[0], t4.y + bias[1], t4.z + bias[2], t4.w + bias[3]};
        const int64_t opix = (img_base + oy) * d.Wo + ox;
        if (plain) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (cq + k >= relu_from) v[k] = fmaxf(v[k], 0.0f);
        } else {
          int64_t rpix = opix;
          if (d.res_up2) rpix = ((int64_t)n * (d.Ho >> 1) + (oy >> 1)) * (d.Wo >> 1) + (ox >> 1);
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
    __syncthreads();   // staging rows are rewritten by the next block
  }
  if (d.out_stats != nullptr) {
    const int cpg = d.Cout / 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float s = gs[k], q = gq[k];
#pragma unroll
      for (int o = 16; o >= N4; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      const int c = cq + k;
      if (lane < N4 && c < d.Cout) {
        const int g = c / cpg;
        atomicAdd(&stat_s[g * 2 + 0], s);
        atomicAdd(&stat_s[g * 2 + 1], q);
      }
    }
    __syncthreads();
    if (tid < 8) atomicAdd(d.out_stats + n * 8 + tid, (double)stat_s[tid]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(a.tmem_cols));
}

using KernelFn = void (*)(const TcArgs);

template <int N, int PASSES>
KernelFn get_kernel() {
  static bool configured = false;
  KernelFn fn = conv_tc_kernel<N, PASSES>;
  if (!configured) {
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    configured = true;
  }
  return fn;
}

KernelFn pick(int n, int passes) {
  if (passes == 3) return n == 16 ? get_kernel<16, 3>() : (n == 32 ? get_kernel<32, 3>() : get_kernel<64, 3>());
  return n == 16 ? get_kernel<16, 1>() : (n == 32 ? get_kernel<32, 1>() : get_kernel<64, 1>());
}

}  // namespace

bool conv_tc_supported(const dmvs_conv_desc& d) {
  return d.w_tc != nullptr && d.stride == 1 && d.KW <= 16 && d.KH <= 16;
}

int dispatch_conv_tc(const dmvs_conv_desc& d, cudaStream_t st) {
  if (!aligned16(d.w_tc)) return DMVS_ERR_ALIGN;
  const int passes = d.precision == DMVS_PREC_TC_TF32 ? 1 : 3;
  TcArgs a;
  a.d = d;
  a.cin_pad = (d.C1 + d.C2 + 7) & ~7;
  a.cout_pad = (d.Cout + 15) & ~15;
  const bool vec_x = aligned16(d.x) && (d.x_ps % 4 == 0) && (d.C1 % 4 == 0);
  const bool vec_x2 = d.C2 == 0 || (aligned16(d.x2) && (d.x2_ps % 4 == 0) && (d.C2 % 4 == 0));
  a.fast_in = vec_x && vec_x2 && d.in_stats == nullptr;
  a.vec_y = aligned16(d.y) && (d.y_ps % 4 == 0);
  a.Hs = d.in_up2 ? d.H / 2 : d.H;
  a.Ws = d.in_up2 ? d.W / 2 : d.W;
  a.w_lo_off = (int64_t)d.KD * d.KH * d.KW * (a.cin_pad / 4) * a.cout_pad * 4;

  int remaining = a.cout_pad, co_base = 0;
  while (remaining > 0) {
    int N = 64;
    while (N > remaining) N >>= 1;   // 64, 32 or 16
    // tile: TW balances the columns so that in_cols <= 128; TH limited by TMEM (n_blk*N <= 256) and shared memory
    const int tw_max = 128 - (d.KW - 1);
    const int ntx = ceil_div(d.Wo, tw_max);
    const int TW = ceil_div(d.Wo, ntx);
    const int in_cols = TW + d.KW - 1;
    int TH = 0, CK = 0, n_blk = 0, plane = 0;
    size_t smem = 0;
    // prefer two CTAs per SM (100 KB each); large-kernel layers (7x7) may take one CTA with up to 200 KB
    // (tall tiles first: a 1-row tile of a 7x7 layer would re-read its input seven times)
    static const int force_th = getenv("DMVS_TC_TH") ? atoi(getenv("DMVS_TC_TH")) : 0;   // tuning aid
    for (int pass = 0; pass < 4 && !CK; ++pass)
    for (int th = (pass < 2 ? 8 : 2); th >= (pass < 2 ? 4 : 1) && !CK; th >>= 1) {
      const size_t budget = (pass & 1) ? 216 * (size_t)1024 : (size_t)kSmemBudget;
      if (th > d.Ho && th > 1) continue;
      if (force_th && th != force_th && pass < 3) continue;
      const int m_total = (th - 1) * in_cols + TW;
      const int nb = ceil_div(m_total, 128);
      if (nb * N > 256) continue;
      const int pl = (nb * 128 + (d.KH - 1) * in_cols + d.KW + 7) & ~7;
      for (int ck = 16; ck >= 8; ck >>= 1) {
        if (ck > a.cin_pad) continue;
        const int quads = ck / 4;
        // raw landing buffer + hi (+ lo) operand planes, double-buffered hi (+ lo) weight slabs, GroupNorm affine
        size_t need = ((size_t)(passes == 3 ? 3 : 2) * quads * pl * 4 +
                       (size_t)(passes == 3 ? 4 : 2) * d.KH * d.KW * quads * N * 4 + 2 * (size_t)d.C1) * 4;
        const size_t stage = (size_t)128 * (N + 4) * 4;
        if (stage > need) need = stage;
        if (need <= budget) { TH = th; CK = ck; n_blk = nb; plane = pl; smem = need; break; }
      }
    }
    if (!CK) return DMVS_ERR_UNSUPPORTED;
    a.co_base = co_base;
    a.CK = CK;
    a.TH = TH;
    a.TW = TW;
    a.in_rows = TH + d.KH - 1;
    a.in_cols = in_cols;
    a.plane = plane;
    a.n_blk = n_blk;
    int cols = 32;
    while (cols < n_blk * N) cols <<= 1;
    a.tmem_cols = cols;
    a.tiles_x = ntx;
    a.tiles_y = ceil_div(d.Ho, TH);
    dim3 grid(a.tiles_x, a.tiles_y, d.N * d.Do);
    if (grid.y > 65535 || grid.z > 65535) return DMVS_ERR_UNSUPPORTED;
    pick(N, passes)<<<grid, kTcThreads, smem, st>>>(a);
    const int rc = launch_status();
    if (rc) return rc;
    co_base += N;
    remaining -= N;
  }
  return 0;
}

}  // namespace dmvs
