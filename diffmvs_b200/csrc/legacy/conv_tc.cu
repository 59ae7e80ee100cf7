// tcgen05 / TMEM implicit-GEMM convolution (stride 1, 2-D / 3-D, channels-last) for sm_100a.
//
// GEMM view per output tile: M = the TH x TW output pixels, N = a chunk of 16/32/64 output channels,
// K = taps x input channels, D accumulates in TMEM (fp32).
//
//   * The input halo tile is staged *planar by channel quad*: A[q][p][4 floats], p = row * in_cols + col the
//     flattened pixel index of the (TH+KH-1) x (TW+KW-1) tile.  8 consecutive pixels x 16 bytes are then one
//     K-major, un-swizzled UMMA core matrix (SBO = 128 B between 8-pixel groups, LBO = plane pitch between the
//     two channel quads of a K=8 step), so an M=128 operand is simply 128 consecutive flattened pixels.
//   * A convolution tap (kh,kw) is a *descriptor offset* of (kh*in_cols + kw)*16 bytes - no im2col, no
//     re-staging: the tensor core reads the tile once per tap straight from shared memory (validated
//     stand-alone in tools/probes/tc_probe.cu).  Flattened positions that fall into halo columns produce values
//     that are never stored.
//   * 3xTF32: each staged element is split ONCE per stage into hi = rna_tf32(x), lo = rna_tf32(x - hi) planes
//     (the legacy mma.sync path re-split per tap and spent >90 % of its instructions there); weights are
//     pre-split on the host.  D += Alo*Bhi + Ahi*Blo + Ahi*Bhi, all issued by one thread.
//   * Persistent CTAs (one per SM, 256 threads) walk a flat stream of stages = (tile, depth tap, channel chunk).
//     Raw fp32 tiles and weight slabs land in an R-deep cp.async ring, R-1 stages ahead of the math, so global
//     latency is paid once per CTA, not once per stage; a single (hi, lo) working buffer is refilled by the split
//     pass while nothing but the previous stage's MMAs has to retire.
//   * Accumulators leave TMEM through tcgen05.ld (warps 0-3 = the 128 lanes) into a small shared staging buffer
//     and the common fused epilogue (bias, residual, activation, GRU blends, GroupNorm statistics, coalesced
//     128-bit stores) while the ring keeps prefetching the next tile.
#include <cstdlib>

#include "conv_common.cuh"

namespace dmvs {
namespace {

constexpr int kTcThreads = 256;
constexpr int kIssuers = 4;   // lane 0 of warps 0-3 each issue the MMAs of every 4th M block (and commit)

struct TcArgs {
  dmvs_conv_desc d;
  int cin_pad;      // (C1+C2) rounded up to 8
  int cout_pad;     // Cout rounded up to 16 (pitch of the packed weights)
  int co_base;      // first output channel of this launch
  int CK;           // channels per stage (8)
  int TH, TW;       // output tile
  int in_rows, in_cols;
  int plane;        // pixels per channel-quad plane (incl. slack for the last M block)
  int n_blk;        // number of M=128 blocks per tile
  int tmem_cols;    // allocated TMEM columns (power of two >= 32)
  int tiles_x, tiles_y, total_tiles;
  int R;            // depth of the raw / weight ring
  int stage_f;      // floats reserved for the (hi, lo) working planes / epilogue staging
  int fast_in, vec_y, vec_res;
  int Hs, Ws;
  int64_t w_lo_off; // offset (floats) of the lo weights inside w_tc
};

struct Stage {
  int tile, kd, chunk;
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

template <int R>
__device__ __forceinline__ void cp_async_wait_ring() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(R - 2) : "memory");
}

// N = output channels per CTA (16, 32, 64), PASSES = 1 (TF32) or 3 (3xTF32), R = ring depth (2..4)
template <int N, int PASSES, int R>
__global__ void __launch_bounds__(kTcThreads, 1) conv_tc_kernel(const __grid_constant__ TcArgs a) {
  constexpr int OP = N + 4;   // pitch of the epilogue staging rows
  constexpr int N4 = N / 4;
  const dmvs_conv_desc& d = a.d;
  extern __shared__ __align__(128) float smem[];
  constexpr int quads = 2;                                    // CK == 8: one K=8 step per tap and stage
  const int taps = d.KH * d.KW;
  const int plane_f = quads * a.plane * 4;                    // floats per operand plane set
  const int wslab_f = taps * quads * N * 4;                   // floats per weight slab
  // ring of R operand pairs: pair p = [hi | lo]; the cp.async data lands in `hi` and is split in place
  float* pair0 = smem;                                        // [R][stage_f]
  float* w_hi0 = pair0 + R * a.stage_f;                       // [R][taps][quads][N][4]
  float* w_lo0 = w_hi0 + R * wslab_f;
  float* gn_s = w_lo0 + (PASSES == 3 ? R * wslab_f : 0);      // [2][C1] when in_stats
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t mbar[2];   // stages alternate barriers: a parity wait may lag by one phase only
  __shared__ unsigned long long stat_s[8];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(a.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&mbar[0])), "r"(kIssuers));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&mbar[1])), "r"(kIssuers));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  const int Ctot = d.C1 + d.C2;
  const int units_per_row = a.in_cols * quads;   // 16-byte units per tile row
  const int nchunks = ceil_div(a.cin_pad, a.CK);

  auto decode = [&](int tile, int& n, int& od, int& ty0, int& tx0) {
    const int tx = tile % a.tiles_x;
    const int r = tile / a.tiles_x;
    const int ty = r % a.tiles_y;
    const int z = r / a.tiles_y;
    n = z / d.Do;
    od = z - n * d.Do;
    ty0 = ty * a.TH;
    tx0 = tx * a.TW;
  };
  auto kd_first = [&](int od) { const int v = d.pad_d - od; return v > 0 ? v : 0; };
  auto kd_last = [&](int od) { const int v = d.D - 1 + d.pad_d - od; return v < d.KD - 1 ? v : d.KD - 1; };
  auto first_stage_of = [&](int tile) {
    Stage s{tile, 0, 0};
    if (tile < a.total_tiles) {
      int n, od, ty0, tx0;
      decode(tile, n, od, ty0, tx0);
      s.kd = kd_first(od);
    }
    return s;
  };
  auto advance = [&](const Stage& c) {
    Stage s = c;
    if (++s.chunk < nchunks) return s;
    s.chunk = 0;
    int n, od, ty0, tx0;
    decode(c.tile, n, od, ty0, tx0);
    if (++s.kd <= kd_last(od)) return s;
    return first_stage_of(c.tile + (int)gridDim.x);
  };

  int gn_n = -1;
  // loads of one stage into ring slot `slot`: halo tile (planar by channel quad) and its weight slab
  auto issue_loads = [&](const Stage& s, int slot) {
    if (s.tile < a.total_tiles) {
      int n, od, ty0, tx0;
      decode(s.tile, n, od, ty0, tx0);
      if (d.in_stats != nullptr && n != gn_n) {   // GroupNorm affine of the producer is per sample
        __syncthreads();
        for (int c = tid; c < d.C1; c += kTcThreads) groupnorm_affine(d, n, c, gn_s);
        __syncthreads();
        gn_n = n;
      }
      const int c0 = s.chunk * a.CK;
      const int id = od + s.kd - d.pad_d;
      const int iy0 = ty0 - d.pad_h, ix0 = tx0 - d.pad_w;
      float* a_raw = pair0 + slot * a.stage_f;   // lands in the hi plane of the pair, split in place later
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
      float* wh = w_hi0 + slot * wslab_f;
      float* wl = w_lo0 + slot * wslab_f;
#pragma unroll 1
      for (int idx = tid; idx < taps * quads * N; idx += kTcThreads) {
        const int nn = idx % N;
        const int r = idx / N;
        const int q = r % quads;
        const int tap = r / quads;
        const bool ok = q0 + q < qtot;
        const int64_t off = ((((int64_t)s.kd * taps + tap) * qtot + q0 + q) * a.cout_pad + a.co_base + nn) * 4;
        cp_async16(wh + idx * 4, ok ? d.w_tc + off : d.w_tc, ok);
        if (PASSES == 3) cp_async16(wl + idx * 4, ok ? d.w_tc + a.w_lo_off + off : d.w_tc, ok);
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");   // always one group per stage slot (possibly empty)
  };

  Stage cur = first_stage_of((int)blockIdx.x);
  if (cur.tile < a.total_tiles) {
    // prologue: R-1 stages in flight
    Stage pre = cur;
#pragma unroll 1
    for (int i = 0; i < R - 1; ++i) {
      issue_loads(pre, i);
      if (pre.tile < a.total_tiles) pre = advance(pre);
    }
    int issued = 0, waited = 0;  // stages whose MMAs were committed / whose completion was consumed (in order)
    bool tile_start = true;
    int slot = 0;
    auto wait_one = [&]() {      // stage t commits to mbar[t & 1]; its phase there has parity (t >> 1) & 1
      mbar_wait(&mbar[waited & 1], (uint32_t)((waited >> 1) & 1));
      ++waited;
    };
    for (;;) {
      cp_async_wait_ring<R>();                           // everything but the newest R-2 groups has landed
      __syncthreads();                                   // pair `slot` (raw data) is visible to every thread
      float* a_hi = pair0 + slot * a.stage_f;
      float* a_lo = a_hi + plane_f;
      // ---- split once per stage, in place: raw -> (hi, lo) operand planes (slack included: harmless) --------
      if (PASSES == 3) {
        const int total = quads * a.plane;
#pragma unroll 1
        for (int u = tid; u < total; u += kTcThreads) {
          const float4 v = *reinterpret_cast<const float4*>(a_hi + u * 4);
          float4 h, l;
          h.x = rna_tf32(v.x); l.x = rna_tf32(v.x - h.x);
          h.y = rna_tf32(v.y); l.y = rna_tf32(v.y - h.y);
          h.z = rna_tf32(v.z); l.z = rna_tf32(v.z - h.z);
          h.w = rna_tf32(v.w); l.w = rna_tf32(v.w - h.w);
          *reinterpret_cast<float4*>(a_hi + u * 4) = h;
          *reinterpret_cast<float4*>(a_lo + u * 4) = l;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncthreads();                                   // operands of this stage complete
      // ---- lane 0 of warps 0..3 issue the MMAs of M blocks warp, warp+4, ... (independent TMEM accumulators) ----
      if (lane == 0 && warp < kIssuers) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        // Descriptors differ only in their 14-bit start-address field (units of 16 bytes), so they are formed
        // once per stage and advanced with integer adds.
        const uint32_t lbo_a = (uint32_t)a.plane * 16u, lbo_b = (uint32_t)N * 16u;
        const uint64_t dah0 = umma_desc(smem_u32(a_hi), lbo_a, 128), dal0 = umma_desc(smem_u32(a_lo), lbo_a, 128);
        const uint64_t dbh0 = umma_desc(smem_u32(w_hi0 + slot * wslab_f), lbo_b, 128);
        const uint64_t dbl0 = umma_desc(smem_u32(w_lo0 + slot * wslab_f), lbo_b, 128);
        const uint32_t b_step = 2u * (uint32_t)N;              // one (tap, K=8 step) of weights, in 16-byte units
        for (int blk = warp; blk < a.n_blk; blk += kIssuers) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(blk * N);
          uint32_t acc = tile_start ? 0u : 1u;
          uint32_t b_off = 0;
          for (int kh = 0; kh < d.KH; ++kh) {
            uint32_t a_off = (uint32_t)(blk * 128 + kh * a.in_cols);
            for (int kw = 0; kw < d.KW; ++kw, ++a_off, b_off += b_step) {
              if (PASSES == 3) {
                umma_tf32(d_tmem, dal0 + a_off, dbh0 + b_off, idesc, acc);
                umma_tf32(d_tmem, dah0 + a_off, dbl0 + b_off, idesc, 1u);
                umma_tf32(d_tmem, dah0 + a_off, dbh0 + b_off, idesc, 1u);
              } else {
                umma_tf32(d_tmem, dah0 + a_off, dbh0 + b_off, idesc, acc);
              }
              acc = 1u;
            }
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                         smem_u32(&mbar[issued & 1]))
                     : "memory");
      }
      ++issued;
      // ---- keep the ring full: stage s+R-1 reuses the pair of stage s-1, whose MMAs must have retired ----------
      if (issued - waited > 1) {
        wait_one();
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      }
      issue_loads(pre, (slot + R - 1) % R);
      if (pre.tile < a.total_tiles) pre = advance(pre);

      const Stage nxt = advance(cur);
      const bool tile_done = nxt.tile != cur.tile;
      tile_start = tile_done;
      if (tile_done) {
        // ---- all MMAs of the tile retired -> accumulators out of TMEM, block by block -----------------------
        while (waited < issued) wait_one();
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        float* out_s = pair0 + slot * a.stage_f;   // this stage's pair is dead now: [128][OP] staging
        int n, od, ty0, tx0;
        decode(cur.tile, n, od, ty0, tx0);
        if (tid < 8) stat_s[tid] = 0ull;
        const int q4 = tid % N4;                 // fixed channel quad per thread in the write-out loop
        const int cq = a.co_base + q4 * 4;
        float bias[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (d.bias != nullptr && cq + k < d.Cout) bias[k] = __ldg(d.bias + cq + k);
        const bool full_quad = cq + 4 <= d.Cout;
        const bool plain = d.epi == DMVS_EPI_STD && (d.act == DMVS_ACT_NONE || d.act == DMVS_ACT_RELU);
        const int relu_from = d.act == DMVS_ACT_RELU ? d.act_c0 : 0x7fffffff;
        const int64_t img_base = (int64_t)(n * d.Do + od) * d.Ho;
        float gs[4] = {0.f, 0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int blk = 0; blk < a.n_blk; ++blk) {
          if (warp < 4) {   // TMEM lane = flattened position within the block; warp w owns lanes [32w, 32w+32)
            float* orow = out_s + tid * OP;
#pragma unroll
            for (int c0 = 0; c0 < N; c0 += 16) {
              uint32_t r[16];
              const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(blk * N + c0);
              asm volatile(
                  "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                  : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                    "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                  : "r"(taddr));
              asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4)
                *reinterpret_cast<float4*>(orow + c0 + j4 * 4) =
                    make_float4(__uint_as_float(r[j4 * 4]), __uint_as_float(r[j4 * 4 + 1]),
                                __uint_as_float(r[j4 * 4 + 2]), __uint_as_float(r[j4 * 4 + 3]));
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
          __syncthreads();   // staging rows are rewritten by the next block / the next split pass
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
              atomicAdd(&stat_s[g * 2 + 0], stat_fixed(s));
              atomicAdd(&stat_s[g * 2 + 1], stat_fixed(q));
            }
          }
          __syncthreads();
          if (tid < 8) atomicAdd(reinterpret_cast<unsigned long long*>(d.out_stats) + n * 8 + tid, stat_s[tid]);
          __syncthreads();
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");   // TMEM reads done before the next tile's MMAs
      }
      if (nxt.tile >= a.total_tiles) break;
      cur = nxt;
      slot = (slot + 1) % R;
    }
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(a.tmem_cols));
}

using KernelFn = void (*)(const TcArgs);

template <int N, int PASSES, int R>
KernelFn get_kernel() {
  static SmemOptIn opt_in;
  KernelFn fn = conv_tc_kernel<N, PASSES, R>;
  opt_in.ensure(fn, 220 * 1024);
  return fn;
}

template <int N, int PASSES>
KernelFn pick_r(int r) {
  return r >= 4 ? get_kernel<N, PASSES, 4>() : (r == 3 ? get_kernel<N, PASSES, 3>() : get_kernel<N, PASSES, 2>());
}

KernelFn pick(int n, int passes, int r) {
  if (passes == 3) return n == 16 ? pick_r<16, 3>(r) : (n == 32 ? pick_r<32, 3>(r) : pick_r<64, 3>(r));
  return n == 16 ? pick_r<16, 1>(r) : (n == 32 ? pick_r<32, 1>(r) : pick_r<64, 1>(r));
}

constexpr size_t kTcSmemMax = 216 * 1024;

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
  a.vec_res = d.res != nullptr && aligned16(d.res) && (d.res_ps % 4 == 0);
  a.Hs = d.in_up2 ? d.H / 2 : d.H;
  a.Ws = d.in_up2 ? d.W / 2 : d.W;
  a.w_lo_off = (int64_t)d.KD * d.KH * d.KW * (a.cin_pad / 4) * a.cout_pad * 4;
  static const int force_th = getenv("DMVS_TC_TH") ? atoi(getenv("DMVS_TC_TH")) : 0;   // tuning aids
  static const int force_r = getenv("DMVS_TC_R") ? atoi(getenv("DMVS_TC_R")) : 0;

  int remaining = a.cout_pad, co_base = 0;
  while (remaining > 0) {
    int N = 64;
    while (N > remaining) N >>= 1;   // 64, 32 or 16
    // tile: TW balances the columns so that in_cols <= 128; TH limited by TMEM (n_blk*N <= 512) and shared memory
    const int tw_max = 128 - (d.KW - 1);
    const int ntx = ceil_div(d.Wo, tw_max);
    const int TW = ceil_div(d.Wo, ntx);
    const int in_cols = TW + d.KW - 1;
    const int CK = 8;                 // one K=8 step per stage and tap keeps the (hi, lo) working set small
    const int quads = CK / 4;
    int TH = 0, R = 0, n_blk = 0, plane = 0, stage_f = 0;
    size_t smem = 0;
    // tallest tile first (halo amortisation, more MMAs per stage: measured faster than a deeper ring), then the
    // deepest ring that still fits
    for (int th = 8; th >= 1 && !TH; th >>= 1) {
      if (th > d.Ho && th > 1) continue;
      if (force_th && th != force_th) continue;
      for (int r = 4; r >= 2 && !TH; --r) {
        if (force_r && r != force_r) continue;
        const int m_total = (th - 1) * in_cols + TW;
        const int nb = ceil_div(m_total, 128);
        if (nb * N > 512) continue;
        const int pl = (nb * 128 + (d.KH - 1) * in_cols + d.KW + 7) & ~7;
        const size_t plane_f = (size_t)quads * pl * 4;
        const size_t wslab_f = (size_t)d.KH * d.KW * quads * N * 4;
        size_t work_f = (passes == 3 ? 2 : 1) * plane_f;                       // one operand pair [hi | lo]
        if (work_f < (size_t)128 * (N + 4)) work_f = (size_t)128 * (N + 4);   // ... doubles as epilogue staging
        const size_t need = (r * work_f + (passes == 3 ? 2 : 1) * r * wslab_f + 2 * (size_t)d.C1) * 4;
        if (need <= kTcSmemMax) { TH = th; R = r; n_blk = nb; plane = pl; stage_f = (int)work_f; smem = need; }
      }
    }
    if (!TH) return DMVS_ERR_UNSUPPORTED;
    a.co_base = co_base;
    a.CK = CK;
    a.TH = TH;
    a.TW = TW;
    a.in_rows = TH + d.KH - 1;
    a.in_cols = in_cols;
    a.plane = plane;
    a.n_blk = n_blk;
    a.R = R;
    a.stage_f = stage_f;
    int cols = 32;
    while (cols < n_blk * N) cols <<= 1;
    a.tmem_cols = cols;
    a.tiles_x = ntx;
    a.tiles_y = ceil_div(d.Ho, TH);
    const long tiles = (long)a.tiles_x * a.tiles_y * d.N * d.Do;
    if (tiles > 0x7fffffffL) return DMVS_ERR_UNSUPPORTED;
    a.total_tiles = (int)tiles;
    const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
    pick(N, passes, R)<<<grid, kTcThreads, smem, st>>>(a);
    const int rc = launch_status();
    if (rc) return rc;
    co_base += N;
    remaining -= N;
  }
  return 0;
}

}  // namespace dmvs
