// tcgen05 / TMEM implicit-GEMM convolution with the kernel-row taps stacked along N ("ws" = width-stacked),
// stride 1, 2-D / 3-D, channels-last, for sm_100a.
//
// Why a second tcgen05 kernel: with the taps as descriptor offsets (conv_tc.cu) every tap is its own MMA with
// N = Cout.  The layers of this network have Cout = 8..32, and an SS-mode M=128 x K=8 TF32 MMA has to read its
// 4 KB A operand from shared memory (128 B/clk => 32 clk) no matter how small N is, while the math needs only
// N/2 clk.  So at Cout = 16 the tensor pipe idles 75 % of the time and the kernel is no faster than FFMA2.
//
// Here the KW taps of one kernel row share ONE MMA:
//     E[p][kw*CC + c] = sum_{kd,kh,k} A[p + kh*in_cols][k] * W[kd][kh][kw][k][c]          (N = KW*CC columns)
//     out[p][c]       = sum_{kw} E[p + kw][kw*CC + c]                                      (shift-add epilogue)
// p = flattened position of the staged (TH+KH-1) x (TW+KW-1) halo tile (pitch in_cols), exactly the planar-by-
// channel-quad operand layout of conv_tc.cu: a kh tap is still a descriptor offset of kh*in_cols positions, the
// kw taps become columns.  One A read now feeds KW*CC >= 24..224 columns, MMA count drops KW-fold, and the
// shift-add costs KW*CC FADDs per pixel (1-10 % of the FMAs it replaces): warp shuffles plus a small shared halo.
//
//   * 3xTF32 (fp32-class): hi = x rounded to TF32, lo = (x - hi) truncated to TF32, D += Alo*Bhi + Ahi*Blo + Ahi*Bhi;
//     activations are split once per stage in shared memory (in place), weights are pre-split on the host (w_ws slabs).
//   * GroupNorm(4)+affine+SiLU of the producer (update.py:117-133) is applied inside the split pass, so those
//     layers keep the asynchronous cp.async landing path.
//   * Persistent CTAs, two per SM (<= 112 KB shared, <= 256 TMEM columns each), warp specialised: warps 0-7 copy and
//     split (each thread its own elements) and run the epilogue, warp 8 issues the MMAs; full/empty mbarriers only.
//   * Stages = (tile, depth tap, stride phase, 8 input channels) stream through an R-deep cp.async ring.
//   * Stride 2 runs as four stride-1 phases over decimated input planes accumulating into the same TMEM columns.
#include <cstdlib>
#include <type_traits>

#include "conv_common.cuh"

namespace dmvs {
namespace {

constexpr int kWorkers = 256;     // warps 0-7: copies, split pass, epilogue
constexpr int kWsThreads = 288;   // + warp 8: issues the MMAs while the workers refill the ring

struct WsArgs {
  dmvs_conv_desc d;
  int cin_pad;      // (C1+C2) rounded up to 8
  int cout_pad;     // pitch of the packed weights (Cout rounded up to 16)
  int co_base;      // first output channel of this launch
  int CC;           // output channels of this launch (multiple of 8)
  int N;            // MMA N = KW*CC rounded up to 16
  int S;            // stride (1 or 2); a stride-2 convolution runs as S*S stride-1 phases over decimated input planes
  int KHe, KWe;     // kernel extent in phase-plane shifts (== KH, KW for stride 1)
  int smin_h, smin_w;   // smallest shift (== -pad for stride 1)
  int TH, TW, in_rows, in_cols;
  int m_total;      // TH * in_cols flattened positions carry results
  int plane;        // positions per channel-quad plane (incl. slack read by the last M block)
  int n_blk;        // M=128 blocks per tile
  int tmem_cols;    // allocated TMEM columns (power of two >= 32)
  int tiles_x, tiles_y, total_tiles;
  int stage_f;      // floats per ring slot: operand pair [hi | lo]
  int halo_f;       // floats of the epilogue's halo exchange buffer
  int vec_y, vec_res, vec_bias;
  int raw_hi;       // 1: keep raw fp32 in the hi plane (hardware truncation), compute only the lo plane
  int Hs, Ws;
  float inv_in_cols;
  int lanes_row;    // threads that share one tile row in the loader (32..256, power of two >= 2*in_cols when possible)
  int64_t w_off;    // offset (floats) of this launch's hi slabs inside w_ws
  int64_t w_plane;  // distance (floats) from a hi slab to its lo twin
};

struct Stage {
  int tile, kd, phase, chunk;
  int n, od, ty0, tx0;   // decoded once per tile (the divisions are not free in a loop that runs per 8 channels)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// K-major, un-swizzled UMMA shared-memory descriptor (8 rows x 16 bytes core matrices)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t v = 0;
  v |= (uint64_t)((saddr >> 4) & 0x3fff);
  v |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  v |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  v |= 1ull << 46;  // descriptor version (Blackwell); layout_type 0 = no swizzle
  return v;
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

// One lane of a converged warp (the compiler keeps the guarded code on the uniform datapath)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
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

// x = hi + lo + r with hi, lo exactly representable in TF32 and |r| < 2^-21 |x|.  hi is x rounded to nearest
// (integer add of half an ulp, then mask - the same result as cvt.rna.tf32.f32 for finite values, which ptxas expands
// to four instructions on sm_100a); lo = x - hi is exact in fp32 and is truncated to TF32 explicitly, so the
// result does not depend on what the tensor core does with the 13 low mantissa bits.  4 instructions per value.
__device__ __forceinline__ void lo_trunc(float x, float& lo) {
  const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = __uint_as_float(__float_as_uint(x - hi) & 0xffffe000u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = __uint_as_float(__float_as_uint(x - hi) & 0xffffe000u);
}

// TMEM -> registers: NCH consecutive fp32 columns of this thread's lane (32x32b shape); no wait inside
template <int NCH>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[NCH]);
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, float (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "r"(taddr));
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, float (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
      : "r"(taddr));
}
// every register written by earlier tcgen05.ld of this thread is valid after this
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
// The compiler does not know that tcgen05.ld results only exist after the wait: route the registers through an empty
// volatile asm placed after it, so that no use of them can be scheduled above the wait (costs no instruction).
template <int NCH>
__device__ __forceinline__ void tmem_pin(float (&v)[NCH]) {
#pragma unroll
  for (int j = 0; j < NCH; ++j) asm volatile("" : "+f"(v[j])::"memory");
}

// 16-byte asynchronous copy without the src-size operand (the zero-fill form costs ~10 extra instructions of
// pointer fix-ups per copy in SASS); padding is written with a plain shared store instead.
__device__ __forceinline__ void cp_async16_full(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

// PASSES = 1 (TF32) or 3 (3xTF32), R = ring depth (2, 3), GN = GroupNorm+SiLU prologue in the split pass
template <int PASSES, int R, bool GN>
__global__ void __launch_bounds__(kWsThreads, 2) conv_ws_kernel(const __grid_constant__ WsArgs a) {
  const dmvs_conv_desc& d = a.d;
  extern __shared__ __align__(128) float smem[];
  const int N = a.N;
  const int plane_f = 2 * a.plane * 4;                        // floats per operand plane set (two channel quads)
  const int wslab_f = a.KHe * 2 * N * 4;                      // floats per weight slab: [kh'][quad][N][4]
  float* pair0 = smem;                                        // [R][stage_f]: pair = [hi | lo], raw data lands in hi
  float* w_hi0 = pair0 + R * a.stage_f;                       // [R][wslab_f]
  float* w_lo0 = w_hi0 + R * wslab_f;
  float* halo_s = w_lo0 + (PASSES == 3 ? R * wslab_f : 0);    // [halo_f]: epilogue exchange of rows across lane quadrants
  float* gn_s = halo_s + a.halo_f;                            // [2][C1] when GN
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t full_bar[R], empty_bar[R];
  __shared__ unsigned long long stat_s[8];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform in the compiler's eyes
  // Programmatic dependent launch: let the next kernel of the stream start its prologue (TMEM allocation, barrier
  // initialisation) while this grid drains; our own prologue likewise overlaps the tail of the previous kernel.
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(a.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    for (int i = 0; i < R; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&full_bar[i])), "r"(kWorkers / 32));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&empty_bar[i])), "r"(1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  asm volatile("griddepcontrol.wait;\n" ::: "memory");   // everything the previous kernels wrote is visible from here on
  const uint32_t tmem_base = tmem_base_s;
  // instruction descriptor: D = f32, A = B = tf32, both K-major, N >> 3, M >> 4
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  const int Ctot = d.C1 + d.C2;
  const int units_per_row = a.in_cols * 2;       // 16-byte units per tile row (two channel quads per stage)
  const int nchunks = a.cin_pad >> 3;
  const int nphase = a.S * a.S;

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
  auto kd_first = [&](int od) { const int v = d.pad_d - od * a.S; return v > 0 ? v : 0; };
  auto kd_last = [&](int od) { const int v = d.D - 1 + d.pad_d - od * a.S; return v < d.KD - 1 ? v : d.KD - 1; };
  auto first_stage_of = [&](int tile) {
    Stage s{tile, 0, 0, 0, 0, 0, 0, 0};
    if (tile < a.total_tiles) {
      decode(tile, s.n, s.od, s.ty0, s.tx0);
      s.kd = kd_first(s.od);
    }
    return s;
  };
  auto advance = [&](const Stage& c) {
    Stage s = c;
    if (++s.chunk < nchunks) return s;
    s.chunk = 0;
    if (++s.phase < nphase) return s;
    s.phase = 0;
    if (++s.kd <= kd_last(c.od)) return s;
    return first_stage_of(c.tile + (int)gridDim.x);
  };

  // ------------------------------------------------------------------------------------------------------------
  // Warp-specialised pipeline.  Workers (warps 0-7) copy, split and - at the end of a tile - run the epilogue; warp 8
  // issues the MMAs.  They meet only through mbarriers: full[slot] (one arrival per worker warp: "our part of this stage is
  // staged, split and fenced") and empty[slot] (one tcgen05.commit arrival: "every MMA issued so far has retired",
  // which frees the slot and, after the last stage of a tile, publishes the accumulators).  A worker splits exactly
  // the elements it copied itself, so workers never wait for each other inside the main loop.
  // ------------------------------------------------------------------------------------------------------------
  const int ld_u0 = tid & (a.lanes_row - 1);
  const int ld_r0 = tid / a.lanes_row;
  const int ld_rstep = kWorkers / a.lanes_row;
  const int up = d.in_up2 ? 1 : 0;

  // copies of one stage into ring slot `slot`: raw halo tile (planar by channel quad) and its weight slab.  A thread
  // owns fixed (column, channel quad) units and walks down the rows: per copy one predicate, one multiply-add, one
  // cp.async.  Zero padding is written with plain stores.
  auto copy_stage = [&](const Stage& s, int slot) {
    if (s.tile < a.total_tiles) {
      const int c0 = s.chunk * 8;
      const int id = s.od * a.S + s.kd - d.pad_d;
      // phase (pa, pb): tile row r / column c hold input pixel (S*(ty0 + r + smin_h) + pa, S*(tx0 + c + smin_w) + pb)
      const int pa = a.S == 2 ? (s.phase >> 1) : 0, pb = a.S == 2 ? (s.phase & 1) : 0;
      const int iy0 = a.S * (s.ty0 + a.smin_h) + pa, ix0 = a.S * (s.tx0 + a.smin_w) + pb;
      float* a_raw = pair0 + slot * a.stage_f;
      const int64_t img = (int64_t)(s.n * d.D + id) * a.Hs;
#pragma unroll 1
      for (int u = ld_u0; u < units_per_row; u += a.lanes_row) {
        const int q = u & 1;
        const int col = u >> 1;
        const int ix = ix0 + a.S * col;
        const int ch = c0 + q * 4;
        const bool col_ok = ix >= 0 && ix < d.W && ch < Ctot;
        const int sx = ix >> up;
        const bool from_x = ch < d.C1;
        const int ps = from_x ? d.x_ps : d.x2_ps;
        const float* base = from_x ? d.x + ch : d.x2 + (ch - d.C1);
        if (!col_ok) base = d.x;
        const float* colp = base + (col_ok ? (img * a.Ws + sx) * ps : 0);
        const int row_stride = a.Ws * ps;            // one image plane stays below 2^31 floats (checked on the host)
        float* dst = a_raw + (q * a.plane + ld_r0 * a.in_cols + col) * 4;
        const int dst_step = ld_rstep * a.in_cols * 4;
#pragma unroll 2
        for (int row = ld_r0; row < a.in_rows; row += ld_rstep, dst += dst_step) {
          const int iy = iy0 + a.S * row;
          if (col_ok && iy >= 0 && iy < d.H) {
            cp_async16_full(dst, colp + (iy >> up) * row_stride);
          } else {
            *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);   // zero padding
          }
        }
      }
      // weights of this (kd, phase, channel chunk): the host packed every slab [kh'][quad][N][4] contiguously (w_ws)
      float* wh = w_hi0 + slot * wslab_f;
      float* wl = w_lo0 + slot * wslab_f;
      const float* src = d.w_ws + a.w_off + (int64_t)((s.kd * nphase + s.phase) * nchunks + s.chunk) * wslab_f;
      const int wunits = wslab_f >> 2;
#pragma unroll 2
      for (int idx = tid; idx < wunits; idx += kWorkers) {
        cp_async16_full(wh + idx * 4, src + idx * 4);
        if (PASSES == 3) cp_async16_full(wl + idx * 4, src + a.w_plane + idx * 4);
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");   // always one group per stage (possibly empty)
  };

  // in-place (hi, lo) split (and GroupNorm+SiLU prologue) of exactly the elements this thread copied
  auto split_stage = [&](const Stage& s, int slot) {
    if (!(PASSES == 3 || GN)) return;
    const int c0 = s.chunk * 8;
    const int iy0 = s.ty0 + a.smin_h, ix0 = s.tx0 + a.smin_w;   // GN layers are stride 1
    float* a_hi = pair0 + slot * a.stage_f;
    const int lo_off = plane_f;
#pragma unroll 1
    for (int u = ld_u0; u < units_per_row; u += a.lanes_row) {
      const int q = u & 1;
      const int col = u >> 1;
      const int ch = c0 + q * 4;
      float4 g1 = make_float4(0.f, 0.f, 0.f, 0.f), g0 = g1;
      bool col_ok = false;
      if (GN) {
        const int ix = ix0 + col;
        col_ok = ix >= 0 && ix < d.W && ch < d.C1;
        if (col_ok) {
          g1 = *reinterpret_cast<const float4*>(gn_s + ch);
          g0 = *reinterpret_cast<const float4*>(gn_s + d.C1 + ch);
        }
      }
      float* ptr = a_hi + (q * a.plane + ld_r0 * a.in_cols + col) * 4;
      const int step = ld_rstep * a.in_cols * 4;
#pragma unroll 2
      for (int row = ld_r0; row < a.in_rows; row += ld_rstep, ptr += step) {
        float4 v = *reinterpret_cast<const float4*>(ptr);
        if (GN) {
          const int iy = iy0 + row;
          if (col_ok && iy >= 0 && iy < d.H) {                 // padding stays zero
            v.x = siluf_(fmaf(v.x, g1.x, g0.x));
            v.y = siluf_(fmaf(v.y, g1.y, g0.y));
            v.z = siluf_(fmaf(v.z, g1.z, g0.z));
            v.w = siluf_(fmaf(v.w, g1.w, g0.w));
          } else {
            v = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if (PASSES == 3) {
          float4 h, l;
          if (a.raw_hi) {
            // the tensor core reads only the 10 upper mantissa bits of a TF32 operand: the raw value can stay in the hi
            // plane (hi = x truncated), only lo = x - trunc(x) is computed (|lo| < 2^-10 |x|, truncated to TF32)
            lo_trunc(v.x, l.x); lo_trunc(v.y, l.y); lo_trunc(v.z, l.z); lo_trunc(v.w, l.w);
            if (GN) *reinterpret_cast<float4*>(ptr) = v;
          } else {
            split_tf32(v.x, h.x, l.x);
            split_tf32(v.y, h.y, l.y);
            split_tf32(v.z, h.z, l.z);
            split_tf32(v.w, h.w, l.w);
            *reinterpret_cast<float4*>(ptr) = h;
          }
          *reinterpret_cast<float4*>(ptr + lo_off) = l;
        } else {
          *reinterpret_cast<float4*>(ptr) = v;
        }
      }
    }
  };

  // Shift-add epilogue of one tile (workers only; see the call site).  All TMEM loads of an item are issued before
  // the single wait when the kernel row has at most three taps (every 3x3 layer), so their latency overlaps.
  auto epilogue = [&](auto nch_tag, int n, int od, int ty0, int tx0, int quadrant, int half) {
    constexpr int NCH = decltype(nch_tag)::value;
    float* halo = halo_s;                     // [item][quadrant][kw-1][lane][NCH]
    const int ncg = a.CC / NCH;
    const int n_items = a.n_blk * ncg;
    const int KWe = a.KWe, KWm1 = KWe - 1;
    const int halo_q = KWm1 * KWm1 * NCH;     // floats per (item, quadrant)
    const bool plain = d.epi == DMVS_EPI_STD && (d.act == DMVS_ACT_NONE || d.act == DMVS_ACT_RELU);
    const int relu_from = d.act == DMVS_ACT_RELU ? d.act_c0 : 0x7fffffff;
    const int64_t img_base = (int64_t)(n * d.Do + od) * d.Ho;
    if (d.out_stats != nullptr && tid < 8) stat_s[tid] = 0ull;
    if (KWm1 > 0) {
      int blk = 0, cg = half;                                // items half, half + 2, ... as (block, channel group)
      while (cg >= ncg) { cg -= ncg; ++blk; }
#pragma unroll 1
      for (int it = half; it < n_items; it += 2) {          // pass 1: rows other quadrants will need
        const uint32_t trow = tmem_base + ((uint32_t)(quadrant * 32) << 16) + (uint32_t)(blk * N + cg * NCH);
        float* hq = halo + (it * 4 + quadrant) * halo_q;
#pragma unroll 1
        for (int kw = 1; kw < KWe; ++kw) {
          float v[NCH];
          tmem_ld<NCH>(trow + (uint32_t)(kw * a.CC), v);
          tmem_wait_ld();
          tmem_pin(v);
          if (lane < kw) {
            float* dst = hq + ((kw - 1) * KWm1 + lane) * NCH;
#pragma unroll
            for (int j = 0; j < NCH; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
        cg += 2;
        while (cg >= ncg) { cg -= ncg; ++blk; }
      }
    }
    asm volatile("bar.sync 1, 256;\n" ::: "memory");   // workers only: halo (and the zeroed statistics) visible
    int blk = 0, cg = half;
    while (cg >= ncg) { cg -= ncg; ++blk; }
#pragma unroll 1
    for (int it = half; it < n_items; it += 2) {            // pass 2: shift-add, fused epilogue, store
      const uint32_t trow = tmem_base + ((uint32_t)(quadrant * 32) << 16) + (uint32_t)(blk * N + cg * NCH);
      const bool have_next = quadrant < 3 || blk + 1 < a.n_blk;
      const float* hn = halo + ((quadrant < 3 ? it : it + ncg) * 4 + ((quadrant + 1) & 3)) * halo_q;
      // value of tap kw for this lane's output: row of lane + kw (shuffle) or the halo of the next quadrant / block
      auto shift_in = [&](float (&v)[NCH], int kw) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) v[j] = __shfl_down_sync(0xffffffffu, v[j], kw);
        if (lane + kw >= 32) {
          if (have_next) {
            const float* src = hn + ((kw - 1) * KWm1 + (lane + kw - 32)) * NCH;
#pragma unroll
            for (int j = 0; j < NCH; j += 4) {
              const float4 h4 = *reinterpret_cast<const float4*>(src + j);
              v[j] = h4.x; v[j + 1] = h4.y; v[j + 2] = h4.z; v[j + 3] = h4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < NCH; ++j) v[j] = 0.0f;      // past the last block: never a valid output
          }
        }
      };
      float acc[NCH];
      if (KWe == 3) {
        float v1[NCH], v2[NCH];
        tmem_ld<NCH>(trow, acc);
        tmem_ld<NCH>(trow + (uint32_t)a.CC, v1);
        tmem_ld<NCH>(trow + (uint32_t)(2 * a.CC), v2);
        tmem_wait_ld();
        tmem_pin(acc);
        tmem_pin(v1);
        tmem_pin(v2);
        shift_in(v1, 1);
        shift_in(v2, 2);
#pragma unroll
        for (int j = 0; j < NCH; ++j) acc[j] = (acc[j] + v1[j]) + v2[j];
      } else {
        tmem_ld<NCH>(trow, acc);
        tmem_wait_ld();
        tmem_pin(acc);
#pragma unroll 1
        for (int kw = 1; kw < KWe; ++kw) {
          float v[NCH];
          tmem_ld<NCH>(trow + (uint32_t)(kw * a.CC), v);
          tmem_wait_ld();
          tmem_pin(v);
          shift_in(v, kw);
#pragma unroll
          for (int j = 0; j < NCH; ++j) acc[j] += v[j];
        }
      }
      const int c0 = a.co_base + cg * NCH;                  // first absolute output channel of this lane
      const int p = blk * 128 + quadrant * 32 + lane;
      const int py = (int)(((float)p + 0.5f) * a.inv_in_cols);
      const int px = p - py * a.in_cols;
      const int oy = ty0 + py, ox = tx0 + px;
      const bool valid = px < a.TW && py < a.TH && oy < d.Ho && ox < d.Wo && c0 < d.Cout;
      float ps[NCH / 2], pq[NCH / 2];                       // per channel pair: sum, sum of squares
#pragma unroll
      for (int j = 0; j < NCH / 2; ++j) ps[j] = pq[j] = 0.0f;
      if (valid) {
        const bool full = c0 + NCH <= d.Cout;
        if (d.bias != nullptr) {
          if (a.vec_bias && full) {
#pragma unroll
            for (int j = 0; j < NCH; j += 4) {
              const float4 b4 = ldg4(d.bias + c0 + j);
              acc[j] += b4.x; acc[j + 1] += b4.y; acc[j + 2] += b4.z; acc[j + 3] += b4.w;
            }
          } else {
#pragma unroll
            for (int k = 0; k < NCH; ++k)
              if (c0 + k < d.Cout) acc[k] += __ldg(d.bias + c0 + k);
          }
        }
        const int64_t opix = (img_base + oy) * d.Wo + ox;
        int64_t rpix = opix;
        if (d.res_up2) rpix = ((int64_t)n * (d.Ho >> 1) + (oy >> 1)) * (d.Wo >> 1) + (ox >> 1);
        if (plain) {   // bias (+ residual before / after) + optional ReLU, inline
          const bool pre_act = d.res_mode == DMVS_RES_PRE_ACT;
          if (d.res_mode != DMVS_RES_NONE) {
            float r[NCH];
            const float* rp = d.res + rpix * d.res_ps + c0;
            if (a.vec_res && full) {
#pragma unroll
              for (int j = 0; j < NCH; j += 4) {
                const float4 r4 = ldg4(rp + j);
                r[j] = r4.x; r[j + 1] = r4.y; r[j + 2] = r4.z; r[j + 3] = r4.w;
              }
            } else {
#pragma unroll
              for (int k = 0; k < NCH; ++k) r[k] = c0 + k < d.Cout ? __ldg(rp + k) : 0.0f;
            }
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
              float x = pre_act ? acc[k] + r[k] : acc[k];
              if (c0 + k >= relu_from) x = fmaxf(x, 0.0f);
              acc[k] = pre_act ? x : x + r[k];
            }
          } else if (relu_from <= c0) {                     // the common case: ReLU on every channel
#pragma unroll
            for (int k = 0; k < NCH; ++k) acc[k] = fmaxf(acc[k], 0.0f);
          } else if (relu_from < c0 + NCH) {
#pragma unroll
            for (int k = 0; k < NCH; ++k)
              if (c0 + k >= relu_from) acc[k] = fmaxf(acc[k], 0.0f);
          }
        } else {
#pragma unroll
          for (int k = 0; k < NCH; ++k)
            if (c0 + k < d.Cout) acc[k] = epilogue_value(d, acc[k], c0 + k, opix, rpix);
        }
        float* yp = d.y + opix * d.y_ps + c0;
        if (a.vec_y && full) {
#pragma unroll
          for (int j = 0; j < NCH; j += 4) *reinterpret_cast<float4*>(yp + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        } else {
#pragma unroll
          for (int k = 0; k < NCH; ++k)
            if (c0 + k < d.Cout) yp[k] = acc[k];
        }
        if (d.out_stats != nullptr) {
#pragma unroll
          for (int k = 0; k < NCH; ++k) {
            const float x = c0 + k < d.Cout ? acc[k] : 0.0f;
            ps[k >> 1] += x;
            pq[k >> 1] += x * x;
          }
        }
      }
      if (d.out_stats != nullptr) {
        // GroupNorm statistics: Cout/4 channels per group is 2, 4 or a multiple of 8 (checked on the host), so a
        // channel pair never straddles two groups
        const int cpg = d.Cout >> 2;
#pragma unroll
        for (int j = 0; j < NCH / 2; ++j) {
          const float s = warp_sum(ps[j]), q = warp_sum(pq[j]);
          const int c = c0 + 2 * j;
          if (lane == 0 && c < d.Cout) {
            const int g = c / cpg;
            atomicAdd(&stat_s[g * 2 + 0], stat_fixed(s));
            atomicAdd(&stat_s[g * 2 + 1], stat_fixed(q));
          }
        }
      }
      cg += 2;
      while (cg >= ncg) { cg -= ncg; ++blk; }
    }
    if (d.out_stats != nullptr) {
      asm volatile("bar.sync 1, 256;\n" ::: "memory");   // every worker's shared atomics are in
      if (tid < 8) atomicAdd(reinterpret_cast<unsigned long long*>(d.out_stats) + n * 8 + tid, stat_s[tid]);
    }
  };

  Stage cur = first_stage_of((int)blockIdx.x);
  if (cur.tile < a.total_tiles) {
    if (warp == 8) {
      // =============================== MMA warp ===============================================================
      const bool leader = elect_one();
      const uint32_t lbo_a = (uint32_t)a.plane * 16u, lbo_b = (uint32_t)N * 16u;
      const uint32_t b_step = 2u * (uint32_t)N;                // one kernel row of weights, in 16-byte units
      int slot = 0, use = 0;                                   // ring slot of `cur` and how often it was used before
      bool tile_start = true;
      for (;;) {
        mbar_wait(&full_bar[slot], (uint32_t)(use & 1));       // operands of this stage staged, split and fenced
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        float* a_hi = pair0 + slot * a.stage_f;
        float* a_lo = a_hi + plane_f;
        const uint64_t dah0 = umma_desc(smem_u32(a_hi), lbo_a, 128), dal0 = umma_desc(smem_u32(a_lo), lbo_a, 128);
        const uint64_t dbh0 = umma_desc(smem_u32(w_hi0 + slot * wslab_f), lbo_b, 128);
        const uint64_t dbl0 = umma_desc(smem_u32(w_lo0 + slot * wslab_f), lbo_b, 128);
        const int pa = a.S == 2 ? (cur.phase >> 1) : 0;
        uint32_t rows = 0;   // kernel rows present in this phase: shift khe <-> kernel row S*(khe + smin_h) + pa + pad_h
        for (int khe = 0; khe < a.KHe; ++khe) {
          const int kh = a.S * (khe + a.smin_h) + pa + d.pad_h;
          if (kh >= 0 && kh < d.KH) rows |= 1u << khe;
        }
        for (int blk = 0; blk < a.n_blk; ++blk) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(blk * N);
          uint32_t acc = tile_start ? 0u : 1u;
          uint32_t a_off = (uint32_t)(blk * 128), b_off = 0;
          for (int khe = 0; khe < a.KHe; ++khe, a_off += (uint32_t)a.in_cols, b_off += b_step) {
            if (!((rows >> khe) & 1u)) continue;
            if (leader) {
              if (PASSES == 3) {
                umma_tf32(d_tmem, dal0 + a_off, dbh0 + b_off, idesc, acc);
                umma_tf32(d_tmem, dah0 + a_off, dbl0 + b_off, idesc, 1u);
                umma_tf32(d_tmem, dah0 + a_off, dbh0 + b_off, idesc, 1u);
              } else {
                umma_tf32(d_tmem, dah0 + a_off, dbh0 + b_off, idesc, acc);
              }
            }
            acc = 1u;
          }
        }
        if (leader)   // arrives on empty[slot] once every MMA issued so far has retired
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                           smem_u32(&empty_bar[slot]))
                       : "memory");
        __syncwarp();
        const Stage nxt = advance(cur);
        tile_start = nxt.tile != cur.tile;
        if (nxt.tile >= a.total_tiles) break;
        cur = nxt;
        if (++slot == R) { slot = 0; ++use; }
      }
    } else {
      // =============================== workers ================================================================
      Stage pre = cur;
#pragma unroll 1
      for (int i = 0; i < R - 1; ++i) {   // prologue: R-1 stages of copies in flight
        copy_stage(pre, i);
        if (pre.tile < a.total_tiles) pre = advance(pre);
      }
      int slot = 0, use = 0;              // ring slot of `cur` / number of earlier uses of that slot
      int pslot = R - 1, puse = 0;        // same for `pre`, the stage whose copies are issued next
      int gn_n = -1;
      const int quadrant = warp & 3, half = warp >> 2;
      for (;;) {
        // ---- this thread's copies of stage `cur` have landed -> split them in place, publish ------------------------
        asm volatile("cp.async.wait_group %0;\n" ::"n"(R - 2) : "memory");
        if (GN && cur.n != gn_n) {                        // GroupNorm affine of the producer is per sample
          // every worker needs the whole table: recompute behind a worker barrier (once per sample)
          asm volatile("bar.sync 1, 256;\n" ::: "memory");
          for (int c = tid; c < d.C1; c += kWorkers) groupnorm_affine(d, cur.n, c, gn_s);
          asm volatile("bar.sync 1, 256;\n" ::: "memory");
          gn_n = cur.n;
        }
        split_stage(cur, slot);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();                                      // one arrival per warp: the lanes' writes are ordered before it
        if (lane == 0)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&full_bar[slot])) : "memory");
        // ---- refill: stage cur+R-1 reuses the slot of stage cur-1, whose MMAs must have retired (the MMAs of `cur`
        // run meanwhile; the split above did not have to wait for them) -----------------------------------------------
        if (puse > 0 && pre.tile < a.total_tiles) {
          mbar_wait(&empty_bar[pslot], (uint32_t)((puse - 1) & 1));
        }
        copy_stage(pre, pslot);
        if (pre.tile < a.total_tiles) pre = advance(pre);
        if (++pslot == R) { pslot = 0; ++puse; }

        const Stage nxt = advance(cur);
        if (nxt.tile != cur.tile) {
          // ---- last stage of the tile: wait for its MMAs, then the shift-add epilogue ----------------------------
          const int n = cur.n, od = cur.od, ty0 = cur.ty0, tx0 = cur.tx0;
          mbar_wait(&empty_bar[slot], (uint32_t)(use & 1));
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          // A lane owns one accumulator row (position p) and NCH = 8 or 16 output channels: tap kw of its output lives in
          // row p + kw, i.e. in lane + kw of the same warp (a shuffle) or, for the last KW-1 lanes, in the first rows
          // of the next lane quadrant / next M block (a small shared "halo" written in a first pass).  Work items are
          // (M block, channel group) pairs, split between the two warp halves (warp % 4 = TMEM lane quadrant).
          if ((a.CC & 15) == 0)
            epilogue(std::integral_constant<int, 16>{}, n, od, ty0, tx0, quadrant, half);
          else
            epilogue(std::integral_constant<int, 8>{}, n, od, ty0, tx0, quadrant, half);
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");   // TMEM reads ordered before the next full[] arrival
        }
        if (nxt.tile >= a.total_tiles) break;
        cur = nxt;
        if (++slot == R) { slot = 0; ++use; }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(a.tmem_cols));
}

using KernelFn = void (*)(const WsArgs);

template <int PASSES, int R, bool GN>
KernelFn get_kernel() {
  static SmemOptIn opt_in;
  KernelFn fn = conv_ws_kernel<PASSES, R, GN>;
  opt_in.ensure(fn, 220 * 1024);
  return fn;
}

template <int PASSES>
KernelFn pick_r(int r, bool gn) {
  if (gn) return r >= 3 ? get_kernel<PASSES, 3, true>() : get_kernel<PASSES, 2, true>();
  return r >= 3 ? get_kernel<PASSES, 3, false>() : get_kernel<PASSES, 2, false>();
}

struct TileCfg {
  int TH = 0, TW = 0, in_cols = 0, n_blk = 0, plane = 0, R = 0, stage_f = 0, halo_f = 0, ctas = 0;
  size_t smem = 0;
  double est = 1e30;   // modelled cycles for the whole launch
};

// Tile search.  For every (tile height, M blocks, CTAs per SM) that fits shared memory and TMEM, a rough cycle
// model of one tile (MMA time, split pass, copy issue, exposed latency, epilogue) times the number of tile waves
// is evaluated and the cheapest configuration wins.  The model only has to rank shapes; constants are from the
// ncu captures under profiles/.
void choose_tile(const dmvs_conv_desc& d, int KHe, int KWe, int N, int CC, int passes, int cin_pad, TileCfg& best) {
  static const int force_th = getenv("DMVS_WS_TH") ? atoi(getenv("DMVS_WS_TH")) : 0;   // tuning aids
  static const int force_r = getenv("DMVS_WS_R") ? atoi(getenv("DMVS_WS_R")) : 0;
  static const int force_ctas = getenv("DMVS_WS_CTAS") ? atoi(getenv("DMVS_WS_CTAS")) : 0;
  const int nchunks = cin_pad >> 3;
  static const int max_ctas = getenv("DMVS_WS_MAXCTAS") ? atoi(getenv("DMVS_WS_MAXCTAS")) : 2;   // 3-4 measured slower
  for (int ctas = max_ctas; ctas >= 1; --ctas) {
    if (force_ctas && ctas != force_ctas) continue;
    // co-resident CTAs share 227 KB of shared memory and 512 TMEM columns (allocations are powers of two)
    const size_t smem_limit = ctas >= 4 ? 55 * 1024 : (ctas == 3 ? 74 * 1024 : (ctas == 2 ? 112 * 1024 : 216 * 1024));
    const int max_blk = (ctas >= 3 ? 128 : (ctas == 2 ? 256 : 512)) / N;
    for (int th = 16; th >= 1; th >>= 1) {
      if (th > 1 && th / 2 >= d.Ho) continue;          // a shorter tile already covers the image height
      if (force_th && th != force_th) continue;
      for (int nb = max_blk; nb >= 1; --nb) {
        const int cols_max = nb * 128 / th;            // in_cols such that th*in_cols <= nb*128
        int tw_max = cols_max - (KWe - 1);
        if (tw_max > 250) tw_max = 250;
        if (tw_max < 1 || (tw_max < 8 && tw_max < d.Wo)) continue;
        const int ntx = ceil_div(d.Wo, tw_max);
        int TW = ceil_div(d.Wo, ntx);
        // operand rows of 16 bytes: a kernel-row offset of in_cols positions keeps the 8-row core matrices on 128-byte
        // lines only if in_cols is a multiple of 8
        static const int align8 = getenv("DMVS_WS_ALIGN8") ? atoi(getenv("DMVS_WS_ALIGN8")) : 0;
        if (align8) {
          const int padded = ((TW + KWe - 1 + 7) & ~7) - (KWe - 1);
          if (padded <= tw_max) TW = padded;
        }
        const int in_cols = TW + KWe - 1;
        const int in_rows = th + KHe - 1;
        const int n_blk = ceil_div(th * in_cols, 128);
        if (n_blk > max_blk) continue;
        const int plane = (n_blk * 128 + (KHe - 1) * in_cols + 8 + 7) & ~7;
        size_t work_f = (size_t)(passes == 3 ? 2 : 1) * 2 * plane * 4;
        const size_t halo_f = ((size_t)n_blk * (CC / 8) * 4 * (KWe - 1) * (KWe - 1) * 8 + 31) & ~(size_t)31;   // epilogue halo exchange
        work_f = (work_f + 31) & ~(size_t)31;
        const size_t wslab_f = (size_t)KHe * 2 * N * 4;
        for (int r = 3; r >= 2; --r) {
          if (force_r && r != force_r) continue;
          const size_t need = (r * work_f + (passes == 3 ? 2 : 1) * r * wslab_f + halo_f + 2 * (size_t)d.C1 + 8) * 4;
          if (need > smem_limit) continue;
          // ---- cycle model ------------------------------------------------------------------------
          const double mma = (double)n_blk * d.KH / d.stride * passes * (N / 2 > 32 ? N / 2 : 32);   // A read 32 clk or math N/2
          const double split = passes == 3 || d.in_stats ? 2.0 * plane / kWorkers * (d.in_stats ? 45.0 : 24.0) : 0.0;
          const double copies = (2.0 * in_rows * in_cols * 10.0 + wslab_f / 4.0 * passes * 6.0) / kWorkers;
          const double issue = (split + copies) * 8.0 / 4.0 + 200.0;        // 8 warps over 4 schedulers + barriers
          const double latency = r == 3 ? 600.0 : 1500.0;                    // exposed copy latency per stage
          const int stages = nchunks * d.KD * d.stride * d.stride;
          const double epi = (double)n_blk * (CC / 8) * (60.0 + 37.0 * KWe) + 200.0;
          // one CTA alone pays the exposed latency; co-resident CTAs interleave until the tensor pipe or the
          // instruction issue slots saturate
          const double alone = stages * ((mma > latency ? mma : latency) + issue) + epi + 300.0;
          const double mma_t = stages * mma, issue_t = stages * issue + epi;
          const double busy = (mma_t > issue_t ? mma_t : issue_t) + 0.25 * (mma_t > issue_t ? issue_t : mma_t);
          const double tile = alone / ctas > busy ? alone / ctas : busy;
          const long tiles = (long)ntx * ceil_div(d.Ho, th) * d.N * d.Do;
          const double waves = (double)ceil_div64(tiles, (int64_t)kNumSMs);   // per SM
          const double est = waves * tile;
          if (est < best.est) {
            best.TH = th; best.TW = TW; best.in_cols = in_cols; best.n_blk = n_blk; best.plane = plane; best.R = r;
            best.stage_f = (int)work_f; best.halo_f = (int)halo_f; best.smem = need; best.est = est; best.ctas = ctas;
          }
        }
      }
    }
  }
}

// Output-channel chunking shared with the host packer (packing.py::pack_ws): chunks of at most cc_max channels.
// Extent of a K-tap kernel in phase-plane shifts for stride S: tap k reads input S*o + k - pad = S*(o + s) + phase
// with shift s = floor((k - pad - phase) / S); smin is the smallest shift over all taps, the extent their span.
inline int floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
inline void ws_extent(int K, int pad, int S, int* smin, int* ext) {
  *smin = floor_div(-pad, S);
  *ext = floor_div(K - 1 - pad, S) - *smin + 1;
}

inline int ws_cc_max(int KW) {
  int cc = (256 / KW) & ~7;
  return cc > 64 ? 64 : cc;
}

}  // namespace

bool conv_ws_supported(const dmvs_conv_desc& d) {
  if (d.w_ws == nullptr || (d.stride != 1 && d.stride != 2) || d.KH > 16) return false;
  int smin, kwe;
  ws_extent(d.KW, d.pad_w, d.stride, &smin, &kwe);
  if (kwe > 8) return false;
  if (d.stride == 2 && (d.in_up2 || d.in_stats != nullptr)) return false;
  const bool vec_x = aligned16(d.x) && (d.x_ps % 4 == 0) && (d.C1 % 4 == 0);
  const bool vec_x2 = d.C2 == 0 || (aligned16(d.x2) && (d.x2_ps % 4 == 0) && (d.C2 % 4 == 0));
  if (!vec_x || !vec_x2) return false;
  if (d.in_stats != nullptr && (d.in_up2 || d.C2 != 0 || !aligned16(d.in_g1) || !aligned16(d.in_g0))) return false;
  const int64_t ps_max = d.x_ps > d.x2_ps ? d.x_ps : d.x2_ps;
  if ((int64_t)d.H * d.W * ps_max >= (1ll << 31)) return false;   // 32-bit row offsets inside one image plane
  if (d.out_stats != nullptr) {   // the epilogue reduces GroupNorm statistics per channel pair
    const int cpg = d.Cout / 4;
    if ((d.Cout % 4) != 0 || !(cpg == 2 || cpg == 4 || cpg % 8 == 0)) return false;
  }
  return true;
}

// plan_out (optional, host pointer): per launch {CC, N, TH, TW, n_blk, R, ctas, smem bytes}; nothing is launched.
int dispatch_conv_ws(const dmvs_conv_desc& d, cudaStream_t st, int32_t* plan_out = nullptr, int plan_cap = 0);

int plan_conv_ws(const dmvs_conv_desc& d, int32_t* out, int cap) { return dispatch_conv_ws(d, nullptr, out, cap); }

int dispatch_conv_ws(const dmvs_conv_desc& d, cudaStream_t st) { return dispatch_conv_ws(d, st, nullptr, 0); }

int dispatch_conv_ws(const dmvs_conv_desc& d, cudaStream_t st, int32_t* plan_out, int plan_cap) {
  if (!conv_ws_supported(d)) return DMVS_ERR_UNSUPPORTED;
  if (!aligned16(d.w_ws)) return DMVS_ERR_ALIGN;
  const int passes = d.precision == DMVS_PREC_WS_TF32 ? 1 : 3;
  WsArgs a;
  a.d = d;
  a.cin_pad = (d.C1 + d.C2 + 7) & ~7;
  a.cout_pad = (d.Cout + 15) & ~15;
  a.vec_y = aligned16(d.y) && (d.y_ps % 4 == 0);
  a.vec_res = d.res != nullptr && aligned16(d.res) && (d.res_ps % 4 == 0);
  a.vec_bias = d.bias != nullptr && aligned16(d.bias);
  static const int raw_hi = getenv("DMVS_WS_RAWHI") ? atoi(getenv("DMVS_WS_RAWHI")) : 0;
  a.raw_hi = raw_hi;
  a.Hs = d.in_up2 ? d.H / 2 : d.H;
  a.Ws = d.in_up2 ? d.W / 2 : d.W;

  a.S = d.stride;
  ws_extent(d.KH, d.pad_h, d.stride, &a.smin_h, &a.KHe);
  ws_extent(d.KW, d.pad_w, d.stride, &a.smin_w, &a.KWe);
  const int cc_max = ws_cc_max(a.KWe);
  if (cc_max < 8) return DMVS_ERR_UNSUPPORTED;
  int remaining = (d.Cout + 7) & ~7, co_base = 0;
  int64_t w_off = 0;
  int n_launch = 0;
  while (remaining > 0) {
    const int CC = remaining < cc_max ? remaining : cc_max;
    const int N = (a.KWe * CC + 15) & ~15;
    TileCfg t;
    choose_tile(d, a.KHe, a.KWe, N, CC, passes, a.cin_pad, t);
    if (!t.TH) return DMVS_ERR_UNSUPPORTED;
    a.co_base = co_base;
    a.CC = CC;
    a.N = N;
    a.TH = t.TH;
    a.TW = t.TW;
    a.in_rows = t.TH + a.KHe - 1;
    a.in_cols = t.in_cols;
    a.m_total = t.TH * t.in_cols;
    a.plane = t.plane;
    a.n_blk = t.n_blk;
    a.stage_f = t.stage_f;
    a.halo_f = t.halo_f;
    a.inv_in_cols = 1.0f / (float)t.in_cols;
    int lanes = 32;
    while (lanes < 2 * t.in_cols && lanes < kWorkers) lanes <<= 1;
    a.lanes_row = lanes;
    // packed slabs of this chunk: [hi | lo][KD][S*S phases][cin_pad/8][KHe][2][N][4]
    const int64_t plane_w = (int64_t)d.KD * a.S * a.S * (a.cin_pad >> 3) * a.KHe * 2 * N * 4;
    a.w_off = w_off;
    a.w_plane = plane_w;
    int cols = 32;
    while (cols < t.n_blk * N) cols <<= 1;
    a.tmem_cols = cols;
    a.tiles_x = ceil_div(d.Wo, t.TW);
    a.tiles_y = ceil_div(d.Ho, t.TH);
    const long tiles = (long)a.tiles_x * a.tiles_y * d.N * d.Do;
    if (tiles > 0x7fffffffL) return DMVS_ERR_UNSUPPORTED;
    a.total_tiles = (int)tiles;
    const long max_grid = (long)kNumSMs * t.ctas;
    const int grid = (int)(tiles < max_grid ? tiles : max_grid);
    if (plan_out != nullptr) {
      if (n_launch < plan_cap) {
        int32_t* o = plan_out + 8 * n_launch;
        o[0] = CC; o[1] = N; o[2] = t.TH; o[3] = t.TW; o[4] = t.n_blk; o[5] = t.R; o[6] = t.ctas; o[7] = (int32_t)t.smem;
      }
    } else {
      KernelFn fn = passes == 3 ? pick_r<3>(t.R, d.in_stats != nullptr) : pick_r<1>(t.R, d.in_stats != nullptr);
      launch_pdl(fn, dim3(grid), dim3(kWsThreads), t.smem, st, a);
      const int rc = launch_status();
      if (rc) return rc;
    }
    ++n_launch;
    co_base += CC;
    remaining -= CC;
    w_off += 2 * plane_w;
  }
  return plan_out != nullptr ? n_launch : 0;
}

}  // namespace dmvs
