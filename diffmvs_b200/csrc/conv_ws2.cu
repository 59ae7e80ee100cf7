// Width-stacked tcgen05 / TMEM implicit-GEMM convolution, second generation ("ws2"): the arithmetic and operand
// layouts of conv_ws.cu (kernel-row taps stacked along N, 3xTF32, shift-add epilogue - see the header of that file)
// behind a fully asynchronous, role-specialised pipeline:
//
//   warp 27     TMA producer   cp.async.bulk.tensor boxes of the raw fp32 halo tile (one 4-channel quad plane per copy,
//                              hardware zero fill = the convolution's padding) + cp.async.bulk of the stage's weight
//                              slabs, all completing on tma_full[slot] (mbarrier transaction count)
//   warps 20-26 split workers  raw -> (hi, lo) TF32 pair in place (+ GroupNorm+SiLU of the producer layer), then
//                              fence.proxy.async and one arrival per warp on op_full[slot]
//   warps 0-3   MMA issuers    FOUR issuing threads, one per scheduler, M blocks dealt round-robin: measured on B200
//                              (profiles/r2_pipe_probe*.txt, r2_mma_issue_experiments.txt) a single thread sustains one
//                              tcgen05.mma per ~100 clk whatever N is, four sustain one per ~70 clk in aggregate, while
//                              an M128 x N48 x K8 TF32 MMA needs 24 clk of math - the NUMBER of MMAs bounds the layers
//                              with many input channels.  Each issuer commits to empty[slot] (ring slot free) and,
//                              after a tile's last stage, to acc_full[set] (accumulators complete)
//   warps 4-19  epilogue       shift-add + fused epilogue of tile t out of accumulator set t&1 while the MMAs of tile
//                              t+1 fill the other set; acc_empty[set] hands the set back.  Four warps per TMEM lane
//                              quadrant, 8 output channels per work item, and a specialised straight-line path for the
//                              common bias/ReLU/residual case; when the staged tile is exactly 32 positions wide a lane
//                              quadrant is one tile row and no halo exchange between quadrants is needed.  The epilogue
//                              - not the tensor pipe - bounds the layers with <= 8 input channels.
//
// Variants selected on the host (dispatch_conv_ws2): kernel rows paired along K for <= 4 input channels (`pair`), phase
// launches with one-sided padding and strided output rows (ops.conv_up2).  One persistent CTA per SM; the ring is
// R = 2..4 stages deep, a stage = (tile, depth tap, stride phase, 8 input channels).  Compared with conv_ws.cu nothing in
// a CTA waits for global-memory latency any more: the producer runs up to R-1 stages ahead, the split of stage s+1
// overlaps the MMAs of stage s, and the epilogue overlaps the next tile.  Every launch is a programmatic dependent launch:
// barrier set-up, TMEM allocation and descriptor prefetch run before griddepcontrol.wait.
#include <cstdlib>
#include <type_traits>

#include "conv_common.cuh"
#include "tcgen05_common.cuh"

namespace dmvs {
namespace {

using namespace tc;

constexpr int kMmaWarps = 4;                                       // warps 0-3: one MMA-issuing thread per scheduler
constexpr int kEpiWarps = 16;                                      // warps 4-19: warp % 4 = TMEM lane quadrant, four warps per quadrant
constexpr int kSplitWarps = 7;                                     // warps 20-26
constexpr int kFirstEpiWarp = kMmaWarps;
constexpr int kFirstSplitWarp = kFirstEpiWarp + kEpiWarps;
constexpr int kTmaWarp = kFirstSplitWarp + kSplitWarps;           // warp 20
constexpr int kWs2Threads = 32 * (kTmaWarp + 1);                  // 672
constexpr int kSplitThreads = 32 * kSplitWarps;
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kMaxRing = 4;

struct alignas(64) Ws2Args {
  CUtensorMap map_x;     // (C, W, H, D, N) view of x, box {4, S*in_cols, S*in_rows, 1, 1}, traversal step S in W and H
  CUtensorMap map_x2;    // same for x2 (only read when C2 > 0)
  dmvs_conv_desc d;
  int cin_pad;           // (C1+C2) rounded up to 8
  int co_base;           // first output channel of this launch
  int CC;                // output channels of this launch (multiple of 8)
  int N;                 // MMA N = KWe*CC rounded up to 16
  int S;                 // stride (1 or 2): a stride-2 convolution runs as S*S stride-1 phases over decimated planes
  int KHe, KWe;          // kernel extent in phase-plane shifts
  int smin_h, smin_w;
  int TH, TW, in_rows, in_cols;
  int plane;             // positions per channel-quad plane (multiple of 8; incl. slack read by the last M block)
  int box_units;         // in_rows * in_cols: 16-byte units per quad plane written by one TMA box
  int n_blk;             // M=128 blocks per tile
  int acc_cols;          // TMEM columns of one accumulator set (n_blk * N)
  int tmem_cols;         // allocated TMEM columns (power of two >= 2 * acc_cols)
  int tiles_x, tiles_y, total_tiles;
  int R;                 // ring depth
  int stage_f;           // floats per operand plane set of one stage (2 * plane * 4)
  int wslab_f;           // floats per weight slab (KHe * 2 * N * 4)
  int halo_f;            // floats of ONE halo exchange buffer (two are allocated, alternating per tile)
  int vec_y, vec_res, vec_bias;
  int fast_epi;          // 1: the epilogue's straight-line path applies (see the kernel)
  int pair;              // 1: <= 4 input channels, kernel rows paired along K (see dispatch_conv_ws2): one quad plane per
                         //    stage, the MMA's second K quad is the same plane one tile row further down
  int KHm;               // row-MMAs per (block, stage): KHe, or ceil(KHe / 2) when paired
  int y_rs, res_rs;      // floats between consecutive output / residual rows (dense: Wo * pixel stride)
  int corr16;            // 1: two MMAs per kernel row instead of three - A_hi*B_hi in TF32 plus ONE kind::f16 MMA (K = 16)
                         //    computing A_lo*B_hi + A_hi*B_lo from fp16 copies of the correction operands (see dispatch)
  float inv_in_cols;
  int64_t w_off;         // offset (floats) of this launch's hi slabs inside w_ws
  int64_t w_plane;       // distance (floats) from a hi slab to its lo twin
  long long* dbg;        // optional pipeline timeline of CTA 0 (DMVS_WS2_DBG=1): [kDbgRoles][kDbgSlots] clock64 stamps
};

constexpr int kDbgRoles = 12, kDbgSlots = 64;
// stamps: 0 producer: stage issued | 1 split: stage landed | 2 split: stage split | 3 MMA: operands ready |
//         4 MMA: stage issued | 5 epilogue: accumulators ready | 6 epilogue: tile stored | 7 MMA: waited for accumulator set
//         8 epilogue: halo rows loaded | 9 epilogue: halo barrier passed | 10 first item's accumulators in registers |
//         11 first item stored
#define WS2_STAMP(role, idx)                                                                                  \
  do {                                                                                                        \
    if (a.dbg != nullptr && blockIdx.x == 0 && (idx) < kDbgSlots && lane == 0) a.dbg[(role) * kDbgSlots + (idx)] = clock64(); \
  } while (0)

struct Stage {
  int tile, kd, phase, chunk;
  int n, od, ty0, tx0;
};

template <bool GN>
__global__ void __launch_bounds__(kWs2Threads, 1) conv_ws2_kernel(const __grid_constant__ Ws2Args a) {
  const dmvs_conv_desc& d = a.d;
  extern __shared__ __align__(1024) float smem[];
  const int N = a.N;
  float* hi0 = smem;                                   // [R][stage_f]   raw tile lands here, split in place -> hi
  float* lo0 = hi0 + a.R * a.stage_f;                  // [R][stage_f]   (stage_f = one or two quad planes)
  float* w_hi0 = lo0 + a.R * a.stage_f;                // [R][wslab_f]
  float* w_lo0 = w_hi0 + a.R * a.wslab_f;              // [R][wslab_f]
  float* halo0 = w_lo0 + a.R * a.wslab_f;              // [2][halo_f]
  float* gn_s = halo0 + 2 * a.halo_f;                  // [2][C1] when GN
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t tma_full[kMaxRing], op_full[kMaxRing], empty_bar[kMaxRing], acc_full[2], acc_empty[2];
  __shared__ unsigned long long stat_s[2][8];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform in the compiler's eyes

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(a.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 32) {
    for (int i = 0; i < kMaxRing; ++i) {
      mbar_init(&tma_full[i], 1);
      mbar_init(&op_full[i], kSplitWarps);
      mbar_init(&empty_bar[i], kMmaWarps);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], kMmaWarps);
      mbar_init(&acc_empty[i], kEpiWarps);
    }
    mbar_init_fence();
    prefetch_tensormap(&a.map_x);
    if (d.C2 > 0) prefetch_tensormap(&a.map_x2);
  }
  if (tid < 16) stat_s[tid >> 3][tid & 7] = 0ull;
  fence_tc_before();
  __syncthreads();
  fence_tc_after();
  pdl_sync();   // everything above is on-chip set-up; the producing kernel's outputs are read from here on
  const uint32_t tmem_base = tmem_base_s;
  const int nchunks = a.cin_pad >> 3;
  const int nphase = a.S * a.S;

  auto kd_first = [&](int od) { const int v = d.pad_d - od * a.S; return v > 0 ? v : 0; };
  auto kd_last = [&](int od) { const int v = d.D - 1 + d.pad_d - od * a.S; return v < d.KD - 1 ? v : d.KD - 1; };
  auto first_stage_of = [&](int tile) {
    Stage s{tile, 0, 0, 0, 0, 0, 0, 0};
    if (tile < a.total_tiles) {
      const int tx = tile % a.tiles_x;
      const int r = tile / a.tiles_x;
      const int ty = r % a.tiles_y;
      const int z = r / a.tiles_y;
      s.n = z / d.Do;
      s.od = z - s.n * d.Do;
      s.ty0 = ty * a.TH;
      s.tx0 = tx * a.TW;
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

  Stage cur = first_stage_of((int)blockIdx.x);

  if (warp == kTmaWarp) {
    // =============================== TMA producer ==================================================================
    if (elect_one()) {
      int slot = 0, use = 0, n_issued = 0;
      const uint32_t box_bytes = (uint32_t)a.box_units * 16u, w_bytes = (uint32_t)a.wslab_f * 4u;
      while (cur.tile < a.total_tiles) {
        if (use > 0) mbar_wait(&empty_bar[slot], (uint32_t)((use - 1) & 1));   // the MMAs that read this slot have retired
        mbar_arrive_expect_tx(&tma_full[slot], (a.pair ? 1u : 2u) * box_bytes + 2u * w_bytes);
        const int pa = a.S == 2 ? (cur.phase >> 1) : 0, pb = a.S == 2 ? (cur.phase & 1) : 0;
        const int iy0 = a.S * (cur.ty0 + a.smin_h) + pa, ix0 = a.S * (cur.tx0 + a.smin_w) + pb;
        const int id = cur.od * a.S + cur.kd - d.pad_d;
        float* dst = hi0 + slot * a.stage_f;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (q == 1 && a.pair) break;
          const int ch = cur.chunk * 8 + q * 4;
          // channels past C1+C2 (cin_pad rounding) are out of bounds of the tensor map: zero filled
          if (ch < d.C1 || d.C2 == 0)
            tma_load_5d(dst + q * a.plane * 4, &a.map_x, &tma_full[slot], ch, ix0, iy0, id, cur.n);
          else
            tma_load_5d(dst + q * a.plane * 4, &a.map_x2, &tma_full[slot], ch - d.C1, ix0, iy0, id, cur.n);
        }
        const float* wsrc = d.w_ws + a.w_off + (int64_t)((cur.kd * nphase + cur.phase) * nchunks + cur.chunk) * a.wslab_f;
        bulk_load(w_hi0 + slot * a.wslab_f, wsrc, w_bytes, &tma_full[slot]);
        bulk_load(w_lo0 + slot * a.wslab_f, wsrc + a.w_plane, w_bytes, &tma_full[slot]);
        if (a.dbg != nullptr && blockIdx.x == 0 && n_issued < kDbgSlots) a.dbg[0 * kDbgSlots + n_issued] = clock64();
        ++n_issued;
        cur = advance(cur);
        if (++slot == a.R) { slot = 0; ++use; }
      }
    }
  } else if (warp < kMmaWarps) {
    // =============================== MMA issuers ===================================================================
    // warp w issues the MMAs of M blocks w, w + 4, ...; the descriptors' upper words are constant, the lower words
    // (start address field) advance by plain 32-bit adds, so one kernel row costs a handful of scalar instructions
    const bool leader = elect_one();
    const uint32_t idesc = idesc_tf32_m128(N), idesc16 = idesc_f16_m128(N);
    // paired rows: K quad 1 of the A operand is quad 0 one tile row further down (kernel row 2*khp + 1)
    const uint32_t lbo_a = (uint32_t)(a.pair ? a.in_cols : a.plane) * 16u, lbo_b = (uint32_t)N * 16u;
    const uint32_t a_step = (uint32_t)(a.pair ? 2 * a.in_cols : a.in_cols);
    const uint32_t b_step = 2u * (uint32_t)N;                // one kernel row of weights, in 16-byte units
    const uint32_t a_hiword = (uint32_t)(umma_desc(0, lbo_a, 128) >> 32), b_hiword = (uint32_t)(umma_desc(0, lbo_b, 128) >> 32);
    const uint32_t a_lbo = (uint32_t)umma_desc(0, lbo_a, 128), b_lbo = (uint32_t)umma_desc(0, lbo_b, 128);   // LBO field, address 0
    int slot = 0, use = 0, n_stage = 0;
    int acc = 0, acc_use = 0;                                // accumulator set of the current tile / earlier uses of that set
    bool tile_start = true;
    while (cur.tile < a.total_tiles) {
      if (tile_start && acc_use > 0) {                       // the epilogue of the tile two back has drained this set
        mbar_wait(&acc_empty[acc], (uint32_t)((acc_use - 1) & 1));
        fence_tc_after();
      }
      if (warp == 0) WS2_STAMP(7, n_stage);
      mbar_wait(&op_full[slot], (uint32_t)(use & 1));        // operands of this stage are split and fenced
      fence_tc_after();
      if (warp == 0) WS2_STAMP(3, n_stage);
      const uint32_t ah = a_lbo | (smem_u32(hi0 + slot * a.stage_f) >> 4), al = a_lbo | (smem_u32(lo0 + slot * a.stage_f) >> 4);
      const uint32_t bh = b_lbo | (smem_u32(w_hi0 + slot * a.wslab_f) >> 4), bl = b_lbo | (smem_u32(w_lo0 + slot * a.wslab_f) >> 4);
      const int pa = a.S == 2 ? (cur.phase >> 1) : 0;
      uint32_t rows = 0;   // kernel rows present in this phase: shift khe <-> kernel row S*(khe + smin_h) + pa + pad_h
      for (int khe = 0; khe < a.KHe; ++khe) {
        const int kh = a.S * (khe + a.smin_h) + pa + d.pad_h;
        if (kh >= 0 && kh < d.KH) rows |= 1u << khe;
      }
      if (a.pair) rows = (1u << a.KHm) - 1u;                  // stride 1: every kernel-row pair is present
      const uint32_t acc_base = tmem_base + (uint32_t)(acc * a.acc_cols);
      if (leader) {
        for (int blk = warp; blk < a.n_blk; blk += kMmaWarps) {
          const uint32_t d_tmem = acc_base + (uint32_t)(blk * N);
          uint32_t accum = tile_start ? 0u : 1u;
          uint32_t a_off = (uint32_t)(blk * 128), b_off = 0;
          uint32_t r = rows;
#pragma unroll 1
          for (int khe = 0; khe < a.KHm; ++khe, a_off += a_step, b_off += b_step, r >>= 1) {
            if (!(r & 1u)) continue;
            if (a.corr16) {
              umma_tf32_w(d_tmem, ah + a_off, a_hiword, bh + b_off, b_hiword, idesc, accum);
              umma_f16_w(d_tmem, al + a_off, a_hiword, bl + b_off, b_hiword, idesc16, 1u);   // [A_lo | A_hi] x [B_hi ; B_lo]
            } else {
              umma_tf32_w(d_tmem, al + a_off, a_hiword, bh + b_off, b_hiword, idesc, accum);
              umma_tf32_w(d_tmem, ah + a_off, a_hiword, bl + b_off, b_hiword, idesc, 1u);
              umma_tf32_w(d_tmem, ah + a_off, a_hiword, bh + b_off, b_hiword, idesc, 1u);
            }
            accum = 1u;
          }
        }
      }
      const Stage nxt = advance(cur);
      const bool tile_end = nxt.tile != cur.tile;
      if (leader) {
        umma_commit(&empty_bar[slot]);                        // this issuer's MMAs on the slot have retired (4 arrivals free it)
        if (tile_end) umma_commit(&acc_full[acc]);            // ... and its share of the tile's accumulators is complete
      }
      __syncwarp();
      if (warp == 0) WS2_STAMP(4, n_stage);
      ++n_stage;
      tile_start = tile_end;
      if (tile_end && ++acc == 2) { acc = 0; ++acc_use; }
      cur = nxt;
      if (++slot == a.R) { slot = 0; ++use; }
    }
  } else if (warp >= kFirstSplitWarp) {
    // =============================== split workers ================================================================
    const int wtid = tid - 32 * kFirstSplitWarp;
    int slot = 0, use = 0, gn_n = -1, n_stage = 0;
    while (cur.tile < a.total_tiles) {
      if (GN && cur.n != gn_n) {                             // GroupNorm affine of the producer is per sample
        asm volatile("bar.sync 1, %0;\n" ::"n"(kSplitThreads) : "memory");
        for (int c = wtid; c < d.C1; c += kSplitThreads) groupnorm_affine(d, cur.n, c, gn_s);
        asm volatile("bar.sync 1, %0;\n" ::"n"(kSplitThreads) : "memory");
        gn_n = cur.n;
      }
      mbar_wait(&tma_full[slot], (uint32_t)(use & 1));       // raw tile (and weights) of this stage have landed
      if (warp == kFirstSplitWarp) WS2_STAMP(1, n_stage);
      float* hi = hi0 + slot * a.stage_f;
      float* lo = lo0 + slot * a.stage_f;
      const int iy0 = cur.ty0 + a.smin_h, ix0 = cur.tx0 + a.smin_w;   // GN layers are stride 1
      if (a.corr16) {
        // fp16 correction operands: a thread owns a position's 8 channels.  hi (TF32-rounded fp32) stays in place; the
        // "lo" planes receive 16-byte units of 8 halves: plane 0 = fp16(x - hi) (exact residual, rounded once),
        // plane 1 = fp16(hi) saturated - the K = 16 operand row [A_lo | A_hi] of the kind::f16 MMA
        float4 g1a = make_float4(0.f, 0.f, 0.f, 0.f), g0a = g1a, g1b = g1a, g0b = g1a;
        bool oka = false, okb = false;
        if (GN) {
          const int ch = cur.chunk * 8;
          oka = ch < d.C1;
          okb = ch + 4 < d.C1;
          if (oka) { g1a = *reinterpret_cast<const float4*>(gn_s + ch); g0a = *reinterpret_cast<const float4*>(gn_s + d.C1 + ch); }
          if (okb) { g1b = *reinterpret_cast<const float4*>(gn_s + ch + 4); g0b = *reinterpret_cast<const float4*>(gn_s + d.C1 + ch + 4); }
        }
        float4* ph0 = reinterpret_cast<float4*>(hi);
        float4* ph1 = reinterpret_cast<float4*>(hi + a.plane * 4);
        uint4* pc0 = reinterpret_cast<uint4*>(lo);
        uint4* pc1 = reinterpret_cast<uint4*>(lo + a.plane * 4);
#pragma unroll 2
        for (int u = wtid; u < a.box_units; u += kSplitThreads) {
          float4 va = ph0[u], vb = ph1[u];
          if (GN) {
            const int row = (int)(((float)u + 0.5f) * a.inv_in_cols);
            const int col = u - row * a.in_cols;
            const int iy = iy0 + row, ix = ix0 + col;
            const bool inside = iy >= 0 && iy < d.H && ix >= 0 && ix < d.W;   // padding stays zero
            if (oka && inside) {
              va.x = siluf_(fmaf(va.x, g1a.x, g0a.x)); va.y = siluf_(fmaf(va.y, g1a.y, g0a.y));
              va.z = siluf_(fmaf(va.z, g1a.z, g0a.z)); va.w = siluf_(fmaf(va.w, g1a.w, g0a.w));
            } else {
              va = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (okb && inside) {
              vb.x = siluf_(fmaf(vb.x, g1b.x, g0b.x)); vb.y = siluf_(fmaf(vb.y, g1b.y, g0b.y));
              vb.z = siluf_(fmaf(vb.z, g1b.z, g0b.z)); vb.w = siluf_(fmaf(vb.w, g1b.w, g0b.w));
            } else {
              vb = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          // The TF32 operand is read raw: the tensor core ignores the 13 low mantissa bits (truncation, verified on B200),
          // so "hi" = the truncated value needs no store and the exact residual x - trunc(x) goes to the fp16 plane
          float4 ha, hb;
          ha.x = trunc_tf32(va.x); ha.y = trunc_tf32(va.y); ha.z = trunc_tf32(va.z); ha.w = trunc_tf32(va.w);
          hb.x = trunc_tf32(vb.x); hb.y = trunc_tf32(vb.y); hb.z = trunc_tf32(vb.z); hb.w = trunc_tf32(vb.w);
          if (GN) {            // the prologue changed the values: the raw operand has to be written back
            ph0[u] = va;
            ph1[u] = vb;
          }
          // the two correction products are balanced by powers of two (exact): A_lo * 16 meets B_hi / 16 and A_hi / 16 meets
          // B_lo * 16 (packing.pack_ws), which keeps the small residuals out of fp16's subnormal range
          constexpr float kUp = 16.0f, kDn = 0.0625f;
          pc0[u] = make_uint4(pack_f16x2((va.x - ha.x) * kUp, (va.y - ha.y) * kUp), pack_f16x2((va.z - ha.z) * kUp, (va.w - ha.w) * kUp),
                              pack_f16x2((vb.x - hb.x) * kUp, (vb.y - hb.y) * kUp), pack_f16x2((vb.z - hb.z) * kUp, (vb.w - hb.w) * kUp));
          pc1[u] = make_uint4(pack_f16x2(ha.x * kDn, ha.y * kDn), pack_f16x2(ha.z * kDn, ha.w * kDn),
                              pack_f16x2(hb.x * kDn, hb.y * kDn), pack_f16x2(hb.z * kDn, hb.w * kDn));
        }
      } else {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (q == 1 && a.pair) break;
        float4 g1 = make_float4(0.f, 0.f, 0.f, 0.f), g0 = g1;
        bool ch_ok = false;
        if (GN) {
          const int ch = cur.chunk * 8 + q * 4;
          ch_ok = ch < d.C1;
          if (ch_ok) {
            g1 = *reinterpret_cast<const float4*>(gn_s + ch);
            g0 = *reinterpret_cast<const float4*>(gn_s + d.C1 + ch);
          }
        }
        float4* ph = reinterpret_cast<float4*>(hi + q * a.plane * 4);
        float4* pl = reinterpret_cast<float4*>(lo + q * a.plane * 4);
#pragma unroll 4
        for (int u = wtid; u < a.box_units; u += kSplitThreads) {
          float4 v = ph[u];
          if (GN) {
            const int row = (int)(((float)u + 0.5f) * a.inv_in_cols);
            const int col = u - row * a.in_cols;
            const int iy = iy0 + row, ix = ix0 + col;
            if (ch_ok && iy >= 0 && iy < d.H && ix >= 0 && ix < d.W) {   // padding stays zero
              v.x = siluf_(fmaf(v.x, g1.x, g0.x));
              v.y = siluf_(fmaf(v.y, g1.y, g0.y));
              v.z = siluf_(fmaf(v.z, g1.z, g0.z));
              v.w = siluf_(fmaf(v.w, g1.w, g0.w));
            } else {
              v = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          float4 h, l;
          split_tf32(v.x, h.x, l.x);
          split_tf32(v.y, h.y, l.y);
          split_tf32(v.z, h.z, l.z);
          split_tf32(v.w, h.w, l.w);
          ph[u] = h;
          pl[u] = l;
        }
      }
      }
      fence_async_smem();
      __syncwarp();                                          // one arrival per warp: the lanes' writes are ordered before it
      if (lane == 0) mbar_arrive(&op_full[slot]);
      if (warp == kFirstSplitWarp) WS2_STAMP(2, n_stage);
      ++n_stage;
      cur = advance(cur);
      if (++slot == a.R) { slot = 0; ++use; }
    }
  } else {
    // =============================== epilogue warps ==============================================================
    // A lane owns one accumulator row (position p) and NCH = 8 or 16 output channels: tap kw of its output lives in row
    // p + kw, i.e. in lane + kw of the same warp (a shuffle) or, for the last KW-1 lanes, in the first rows of the next
    // lane quadrant / next M block (a small shared "halo", written in a first pass).  Warp w handles TMEM lane quadrant
    // w % 4 of every M block.
    const int quadrant = warp & 3;
    const int half = (warp - kFirstEpiWarp) >> 2;          // the warps of a quadrant take work items round-robin
    constexpr int kPerQuad = kEpiWarps / 4;
    const int etid = tid - 32 * kFirstEpiWarp;
    int dbg_tile = 0;                                      // tile index for the timeline stamps
    // FAST: the straight-line path - standard epilogue with optional bias, optional residual (before or after the
    // activation), ReLU on all channels or none, every access a full 128-bit vector (host-checked)
    // ALIGNED: the staged tile is exactly 32 positions wide, so a TMEM lane quadrant is one tile row and the lanes whose
    // shuffle sources would wrap into the next quadrant are the row's halo columns - never valid outputs: no halo
    // exchange, no barrier, and (row, column) of a lane need no division
    auto epilogue = [&](auto nch_tag, auto kwe_tag, auto fast_tag, auto aligned_tag, int n, int od, int ty0, int tx0,
                        uint32_t acc_base, uint32_t halo_sa, unsigned long long* stat) {
      constexpr int NCH = decltype(nch_tag)::value;
      constexpr int KWT = decltype(kwe_tag)::value;         // 1, 3 or 0 (= run-time kernel width)
      constexpr bool FAST = decltype(fast_tag)::value;
      constexpr bool ALIGNED = decltype(aligned_tag)::value;
      const int KWe = KWT ? KWT : a.KWe, KWm1 = KWe - 1;
      const int ncg = a.CC / NCH;
      const int n_items = a.n_blk * ncg;                    // (block, channel group) pairs, block-major
      const int halo_q = KWm1 * KWm1 * NCH;                 // floats per (item, quadrant)
      const bool plain = d.epi == DMVS_EPI_STD && (d.act == DMVS_ACT_NONE || d.act == DMVS_ACT_RELU);
      const int relu_from = d.act == DMVS_ACT_RELU ? d.act_c0 : 0x7fffffff;
      const int64_t img_base = (int64_t)(n * d.Do + od) * d.Ho;
      const uint32_t tquad = acc_base + ((uint32_t)(quadrant * 32) << 16);
      if (!ALIGNED && KWm1 > 0) {
        int blk = 0, cg = half;                              // items half, half + 2, ...
        while (cg >= ncg) { cg -= ncg; ++blk; }
#pragma unroll 1
        for (int it = half; it < n_items; it += kPerQuad) {        // pass 1: the first KW-1 rows of this quadrant, for its predecessor
          const uint32_t trow = tquad + (uint32_t)(blk * N + cg * NCH);
          const uint32_t hq = halo_sa + (uint32_t)((it * 4 + quadrant) * halo_q) * 4u;
          if (KWT == 3) {
            float v1[NCH], v2[NCH];
            tmem_ld<NCH>(trow + (uint32_t)a.CC, v1);
            tmem_ld<NCH>(trow + (uint32_t)(2 * a.CC), v2);
            tmem_wait_ld();
            tmem_pin(v1);
            tmem_pin(v2);
            if (lane < 2) {
              if (lane == 0) {
#pragma unroll
                for (int j = 0; j < NCH; j += 4) sts128(hq + (uint32_t)j * 4u, make_float4(v1[j], v1[j + 1], v1[j + 2], v1[j + 3]));
              }
#pragma unroll
              for (int j = 0; j < NCH; j += 4)
                sts128(hq + (uint32_t)((2 + lane) * NCH + j) * 4u, make_float4(v2[j], v2[j + 1], v2[j + 2], v2[j + 3]));
            }
          } else {
#pragma unroll 1
            for (int kw = 1; kw < KWe; ++kw) {
              float v[NCH];
              tmem_ld<NCH>(trow + (uint32_t)(kw * a.CC), v);
              tmem_wait_ld();
              tmem_pin(v);
              if (lane < kw) {
#pragma unroll
                for (int j = 0; j < NCH; j += 4)
                  sts128(hq + (uint32_t)(((kw - 1) * KWm1 + lane) * NCH + j) * 4u, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
              }
            }
          }
          cg += kPerQuad;
          while (cg >= ncg) { cg -= ncg; ++blk; }
        }
        if (warp == kFirstEpiWarp) WS2_STAMP(8, dbg_tile);
        asm volatile("bar.sync 2, %0;\n" ::"n"(kEpiThreads) : "memory");   // epilogue warps only: halo visible
        if (warp == kFirstEpiWarp) WS2_STAMP(9, dbg_tile);
      }
      int blk = 0, cg = half;
      while (cg >= ncg) { cg -= ncg; ++blk; }
#pragma unroll 1
      for (int it = half; it < n_items; it += kPerQuad) {   // pass 2: shift-add, fused epilogue, store
        const uint32_t trow = tquad + (uint32_t)(blk * N + cg * NCH);
        const bool have_next = quadrant < 3 || blk + 1 < a.n_blk;
        const uint32_t hn = halo_sa + (uint32_t)(((quadrant < 3 ? it : it + ncg) * 4 + ((quadrant + 1) & 3)) * halo_q) * 4u;
        // value of tap kw for this lane's output: row of lane + kw (shuffle) or the halo of the next quadrant / block
        auto shift_in = [&](float (&v)[NCH], int kw) {
#pragma unroll
          for (int j = 0; j < NCH; ++j) v[j] = __shfl_down_sync(0xffffffffu, v[j], kw);
          if (!ALIGNED && lane + kw >= 32) {
            if (have_next) {
              const uint32_t src = hn + (uint32_t)(((kw - 1) * KWm1 + (lane + kw - 32)) * NCH) * 4u;
#pragma unroll
              for (int j = 0; j < NCH; j += 4) {
                const float4 h4 = lds128(src + (uint32_t)j * 4u);
                v[j] = h4.x; v[j + 1] = h4.y; v[j + 2] = h4.z; v[j + 3] = h4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < NCH; ++j) v[j] = 0.0f;      // past the last block: never a valid output
            }
          }
        };
        const int c0 = a.co_base + cg * NCH;                  // first absolute output channel of this lane
        const int p = blk * 128 + quadrant * 32 + lane;
        const int py = ALIGNED ? blk * 4 + quadrant : (int)(((float)p + 0.5f) * a.inv_in_cols);
        const int px = ALIGNED ? lane : p - py * a.in_cols;
        const int oy = ty0 + py, ox = tx0 + px;
        const bool valid = px < a.TW && py < a.TH && oy < d.Ho && ox < d.Wo && (FAST || c0 < d.Cout);
        const int64_t opix = (img_base + oy) * d.Wo + ox;
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
          if (warp == kFirstEpiWarp && it == half) WS2_STAMP(10, dbg_tile);
          shift_in(v1, 1);
          shift_in(v2, 2);
#pragma unroll
          for (int j = 0; j < NCH; ++j) acc[j] = (acc[j] + v1[j]) + v2[j];
        } else {
          tmem_ld<NCH>(trow, acc);
          tmem_wait_ld();
          tmem_pin(acc);
          if (warp == kFirstEpiWarp && it == half) WS2_STAMP(10, dbg_tile);
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
        float ps[NCH / 2], pq[NCH / 2];                       // per channel pair: sum, sum of squares
#pragma unroll
        for (int j = 0; j < NCH / 2; ++j) ps[j] = pq[j] = 0.0f;
        if (FAST) {
          if (valid) {
            if (d.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < NCH; j += 4) {
                const float4 b4 = ldg4(d.bias + c0 + j);
                acc[j] += b4.x; acc[j + 1] += b4.y; acc[j + 2] += b4.z; acc[j + 3] += b4.w;
              }
            }
            const bool relu = d.act == DMVS_ACT_RELU;
            if (d.res_mode != DMVS_RES_NONE) {
              const float* rp = d.res + (img_base + oy) * a.res_rs + (int64_t)ox * d.res_ps + c0;
              if (d.res_up2) rp = d.res + (((int64_t)n * (d.Ho >> 1) + (oy >> 1)) * (d.Wo >> 1) + (ox >> 1)) * d.res_ps + c0;
              const bool pre_act = d.res_mode == DMVS_RES_PRE_ACT;
#pragma unroll
              for (int j = 0; j < NCH; j += 4) {
                const float4 r4 = ldg4(rp + j);
                const float r[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  float x = pre_act ? acc[j + k] + r[k] : acc[j + k];
                  if (relu) x = fmaxf(x, 0.0f);
                  acc[j + k] = pre_act ? x : x + r[k];
                }
              }
            } else if (relu) {
#pragma unroll
              for (int k = 0; k < NCH; ++k) acc[k] = fmaxf(acc[k], 0.0f);
            }
            float* yp = d.y + (img_base + oy) * a.y_rs + (int64_t)ox * d.y_ps + c0;
#pragma unroll
            for (int j = 0; j < NCH; j += 4) *reinterpret_cast<float4*>(yp + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
            if (d.out_stats != nullptr) {
#pragma unroll
              for (int k = 0; k < NCH; ++k) {
                ps[k >> 1] += acc[k];
                pq[k >> 1] += acc[k] * acc[k];
              }
            }
          }
        } else if (valid) {
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
            } else if (relu_from <= c0) {                     // ReLU on every channel
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
            const float sv = warp_sum(ps[j]), qv = warp_sum(pq[j]);
            const int c = c0 + 2 * j;
            if (lane == 0 && c < d.Cout) {
              const int g = c / cpg;
              atomicAdd(&stat[g * 2 + 0], stat_fixed(sv));
              atomicAdd(&stat[g * 2 + 1], stat_fixed(qv));
            }
          }
        }
        if (warp == kFirstEpiWarp && it == half) WS2_STAMP(11, dbg_tile);
        cg += kPerQuad;
        while (cg >= ncg) { cg -= ncg; ++blk; }
      }
      if (d.out_stats != nullptr) {
        asm volatile("bar.sync 2, %0;\n" ::"n"(kEpiThreads) : "memory");   // every epilogue warp's shared atomics are in
        if (etid < 8) {
          atomicAdd(reinterpret_cast<unsigned long long*>(d.out_stats) + n * 8 + etid, stat[etid]);
          stat[etid] = 0ull;                                   // ready for the tile after next (same buffer)
        }
      }
    };
    auto run_epilogue = [&](auto nch_tag, int n, int od, int ty0, int tx0, uint32_t acc_base, uint32_t halo_sa,
                            unsigned long long* stat) {
      using T = std::true_type;
      using F = std::false_type;
      using K0 = std::integral_constant<int, 0>;
      using K1 = std::integral_constant<int, 1>;
      using K3 = std::integral_constant<int, 3>;
      if (a.fast_epi && a.in_cols == 32) {
        if (a.KWe == 3) epilogue(nch_tag, K3{}, T{}, T{}, n, od, ty0, tx0, acc_base, halo_sa, stat);
        else if (a.KWe == 1) epilogue(nch_tag, K1{}, T{}, T{}, n, od, ty0, tx0, acc_base, halo_sa, stat);
        else epilogue(nch_tag, K0{}, T{}, T{}, n, od, ty0, tx0, acc_base, halo_sa, stat);
      } else if (a.fast_epi) {
        if (a.KWe == 3) epilogue(nch_tag, K3{}, T{}, F{}, n, od, ty0, tx0, acc_base, halo_sa, stat);
        else epilogue(nch_tag, K0{}, T{}, F{}, n, od, ty0, tx0, acc_base, halo_sa, stat);
      } else if (a.in_cols == 32) {
        epilogue(nch_tag, K0{}, F{}, T{}, n, od, ty0, tx0, acc_base, halo_sa, stat);
      } else {
        epilogue(nch_tag, K0{}, F{}, F{}, n, od, ty0, tx0, acc_base, halo_sa, stat);
      }
    };

    int t_local = 0;
    for (int tile = (int)blockIdx.x; tile < a.total_tiles; tile += (int)gridDim.x, ++t_local) {
      const Stage s = first_stage_of(tile);
      const int acc = t_local & 1;
      mbar_wait(&acc_full[acc], (uint32_t)((t_local >> 1) & 1));
      fence_tc_after();
      if (warp == kFirstEpiWarp) WS2_STAMP(5, t_local);
      dbg_tile = t_local;
      const uint32_t acc_base = tmem_base + (uint32_t)(acc * a.acc_cols);
      const uint32_t halo_sa = smem_u32(halo0 + acc * a.halo_f);
      run_epilogue(std::integral_constant<int, 8>{}, s.n, s.od, s.ty0, s.tx0, acc_base, halo_sa, stat_s[acc]);
      fence_tc_before();                                     // TMEM reads ordered before the hand-back
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (warp == kFirstEpiWarp) WS2_STAMP(6, t_local);
    }
  }
  fence_tc_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(a.tmem_cols));
}

using KernelFn = void (*)(const Ws2Args);

template <bool GN>
KernelFn get_kernel() {
  static SmemOptIn opt_in;
  KernelFn fn = conv_ws2_kernel<GN>;
  opt_in.ensure(fn, 225 * 1024);   // + 1 KB of static shared memory (barriers) = the 227 KB a CTA may use
  return fn;
}

struct TileCfg2 {
  int TH = 0, TW = 0, in_cols = 0, n_blk = 0, plane = 0, R = 0, halo_f = 0;
  size_t smem = 0;
  double est = 1e30;   // modelled cycles for the whole launch
};

inline int floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
inline void ws_extent(int K, int pad, int S, int* smin, int* ext) {
  *smin = floor_div(-pad, S);
  *ext = floor_div(K - 1 - pad, S) - *smin + 1;
}
inline int ws_cc_max(int KW) {
  int cc = (256 / KW) & ~7;
  return cc > 64 ? 64 : cc;
}

// Tile search: every (tile height, M blocks, ring depth) that fits 227 KB of shared memory and half of TMEM (two
// accumulator sets) is scored with a small throughput model - per tile the slower of the tensor pipe and the
// shared-memory port (MMA operand reads + split pass + TMA writes), plus a fixed per-tile cost - times the number of
// tile waves over the SMs.  Deeper rings win ties.
void choose_tile2(const dmvs_conv_desc& d, int S, int KHe_in, int KWe, int N, int CC, int cin_pad, bool pair, TileCfg2& best) {
  const int KHm = pair ? (KHe_in + 1) / 2 : KHe_in;      // row-MMAs per (block, stage)
  const int KHe = pair ? 2 * KHm : KHe_in;                // staged rows beyond the tile height + 1 (paired: rounded up to even)
  static const int force_th = getenv("DMVS_WS2_TH") ? atoi(getenv("DMVS_WS2_TH")) : 0;   // tuning aids
  static const int force_r = getenv("DMVS_WS2_R") ? atoi(getenv("DMVS_WS2_R")) : 0;
  static const int force_nb = getenv("DMVS_WS2_NB") ? atoi(getenv("DMVS_WS2_NB")) : 0;
  const int nchunks = cin_pad >> 3;
  const size_t smem_limit = 224 * 1024;
  const int max_blk = 256 / N;
  const size_t wslab_f = (size_t)KHm * 2 * N * 4;
  const int quads = pair ? 1 : 2;
  // Staged tiles exactly 32 positions wide (TW = 33 - KWe output columns) make a TMEM lane quadrant one tile row: the
  // epilogue then needs no halo exchange between quadrants, no barrier and no index division (measured: the epilogue
  // warps' instruction issue, not the tensor pipe, bounds the layers with few input channels).  Used whenever the
  // image is at least that wide; the general shapes remain for narrow maps.
  static const int no_align = getenv("DMVS_WS2_NOALIGN") ? atoi(getenv("DMVS_WS2_NOALIGN")) : 0;
  const bool aligned = !no_align && KWe <= 8 && d.Wo >= 33 - KWe;
  for (int th = 32; th >= 1; th >>= 1) {
    if (th > 1 && th / 2 >= d.Ho) continue;          // a shorter tile already covers the image height
    if (force_th && th != force_th) continue;
    for (int nb = max_blk; nb >= 1; --nb) {
      if (force_nb && nb != force_nb) continue;
      const int cols_max = nb * 128 / th;            // in_cols such that th*in_cols <= nb*128
      int tw_max = cols_max - (KWe - 1);
      if (tw_max > 250) tw_max = 250;
      if (tw_max * S + (KWe - 1) * S > 256) tw_max = 256 / S - (KWe - 1);   // TMA box limit
      if (tw_max < 1 || (tw_max < 8 && tw_max < d.Wo)) continue;
      int ntx = ceil_div(d.Wo, tw_max);
      int TW = ceil_div(d.Wo, ntx);
      if (aligned) {
        if (cols_max < 32 || th * 32 > nb * 128 || (nb > 1 && th * 32 <= (nb - 1) * 128)) continue;   // exactly nb blocks of 4 rows
        TW = 33 - KWe;
        ntx = ceil_div(d.Wo, TW);
      }
      const int in_cols = TW + KWe - 1;
      const int in_rows = th + KHe - 1;
      if (in_rows * S > 256) continue;
      const int n_blk = ceil_div(th * in_cols, 128);
      if (n_blk > max_blk) continue;
      const int plane = (n_blk * 128 + (KHe - 1) * in_cols + 8 + 7) & ~7;
      const size_t stage_f = (size_t)quads * plane * 4;
      const size_t halo_f = ((size_t)n_blk * (CC / 8) * 4 * (KWe - 1) * (KWe - 1) * 8 + 31) & ~(size_t)31;
      for (int r = kMaxRing; r >= 2; --r) {
        if (force_r && r != force_r) continue;
        const size_t need = (r * (2 * stage_f + 2 * wslab_f) + 2 * halo_f + 2 * (size_t)d.C1 + 64) * 4;
        if (need > smem_limit) continue;
        const int stages = nchunks * d.KD * S * S;
        const double rows_per_stage = pair ? (double)KHm : (double)d.KH / S;   // row-MMAs per phase (average)
        const double mma_n = (double)n_blk * rows_per_stage * 3.0;          // MMAs per stage
        const double mma_clk = mma_n * (N / 2 > 32 ? N / 2 : 32);
        const double port_clk = mma_n * (4096.0 + 32.0 * N) / 128.0          // A and B operand reads
                                + quads * in_rows * in_cols * (64.0 + 16.0) / 128.0   // split (16 read + 32 written) + TMA write
                                + 2.0 * wslab_f * 4.0 / 128.0;
        const double split_issue = quads * in_rows * in_cols * (d.in_stats ? 60.0 : 26.0) / 32.0 / 4.0;   // 8 warps on 4 schedulers
        double stage_clk = mma_clk > port_clk ? mma_clk : port_clk;
        if (split_issue > stage_clk) stage_clk = split_issue;
        const double fill = r >= 4 ? 0.0 : (r == 3 ? 40.0 : 700.0);        // exposed load latency per stage when the ring is shallow
        const double epi = (double)n_blk * (CC / 8) * (60.0 + 37.0 * KWe) * 2.0 + 300.0;   // 4 epilogue warps, overlapped
        double tile = stages * (stage_clk + fill + 60.0);
        if (epi > tile) tile = epi;
        tile += 250.0;
        const long tiles = (long)ntx * ceil_div(d.Ho, th) * d.N * d.Do;
        const double waves = (double)ceil_div64(tiles, (int64_t)kNumSMs);
        const double est = waves * tile;
        if (est < best.est) {
          best.TH = th; best.TW = TW; best.in_cols = in_cols; best.n_blk = n_blk; best.plane = plane; best.R = r;
          best.halo_f = (int)halo_f; best.smem = need; best.est = est;
        }
        break;   // shallower rings of the same shape are never better
      }
    }
  }
}

}  // namespace

// Pipeline timeline (tuning aid): with DMVS_WS2_DBG=1 every launch stamps CTA 0's first stages into a device buffer
long long* ws2_debug_buffer() {
  static long long* buf = nullptr;
  static const bool on = getenv("DMVS_WS2_DBG") && atoi(getenv("DMVS_WS2_DBG")) != 0;
  if (on && buf == nullptr) {
    if (cudaMalloc(&buf, sizeof(long long) * kDbgRoles * kDbgSlots) != cudaSuccess) buf = nullptr;
  }
  return on ? buf : nullptr;
}

bool conv_ws2_supported(const dmvs_conv_desc& d) {
  if (!conv_ws_supported(d)) return false;
  if (d.in_up2) return false;                                  // a TMA box cannot replicate pixels
  if (d.precision == DMVS_PREC_WS_TF32) return false;          // the single-pass mode stays on conv_ws.cu
  if ((d.x_ps % 4) != 0 || (d.C2 > 0 && (d.x2_ps % 4) != 0)) return false;
  return true;
}

int dispatch_conv_ws2(const dmvs_conv_desc& d, cudaStream_t st, int32_t* plan_out, int plan_cap) {
  if (!conv_ws2_supported(d)) return DMVS_ERR_UNSUPPORTED;
  if (!aligned16(d.w_ws)) return DMVS_ERR_ALIGN;
  Ws2Args a;
  a.d = d;
  a.cin_pad = (d.C1 + d.C2 + 7) & ~7;
  a.vec_y = aligned16(d.y) && (d.y_ps % 4 == 0);
  a.vec_res = d.res != nullptr && aligned16(d.res) && (d.res_ps % 4 == 0);
  a.vec_bias = d.bias != nullptr && aligned16(d.bias);
  a.S = d.stride;
  ws_extent(d.KH, d.pad_h, d.stride, &a.smin_h, &a.KHe);
  ws_extent(d.KW, d.pad_w, d.stride, &a.smin_w, &a.KWe);
  const int cc_max = ws_cc_max(a.KWe);
  if (cc_max < 8) return DMVS_ERR_UNSUPPORTED;
  // <= 4 input channels: a K = 8 MMA would multiply an all-zero channel quad.  Instead the second quad is the SAME staged
  // plane one tile row further down (the A descriptor's leading-dimension offset is one row of positions instead of one
  // plane), i.e. K = (kernel rows 2j, 2j+1) x 4 channels: ceil(KH/2) row-MMAs, one TMA box and half the split work per
  // stage.  The weights come pair-packed (`w_ws_pair`, packing.pack_ws_pair).
  a.pair = d.w_ws_pair != nullptr && aligned16(d.w_ws_pair) && a.S == 1 && d.C1 <= 4 && d.C2 == 0 && d.KH >= 2 &&
           d.in_stats == nullptr;
  a.KHm = a.pair ? (a.KHe + 1) / 2 : a.KHe;
  if (a.pair) a.d.w_ws = d.w_ws_pair;
  // DMVS_PREC_WS2_TF32_F16C: the two correction products of the 3xTF32 scheme, A_lo*B_hi + A_hi*B_lo, come from ONE
  // kind::f16 MMA (K = 16: operand rows [A_lo | A_hi] and [B_hi ; B_lo] as fp16, 11 significant bits - the same as the
  // TF32 copies they replace; A_hi saturates at 65504, where only the correction term degrades).  Two MMAs per kernel
  // row instead of three.  Needs the slabs packed with fp16 correction planes (`w_ws16`).
  a.corr16 = (d.precision == DMVS_PREC_WS2_TF32_F16C && d.w_ws16 != nullptr && !a.pair) ? 1 : 0;
  if (a.corr16) {
    if (!aligned16(d.w_ws16)) return DMVS_ERR_ALIGN;
    a.d.w_ws = d.w_ws16;
  }
  int remaining = (d.Cout + 7) & ~7, co_base = 0;
  int64_t w_off = 0;
  int n_launch = 0;
  while (remaining > 0) {
    const int CC = remaining < cc_max ? remaining : cc_max;
    const int N = (a.KWe * CC + 15) & ~15;
    TileCfg2 t;
    choose_tile2(d, a.S, a.KHe, a.KWe, N, CC, a.cin_pad, a.pair != 0, t);
    if (!t.TH) return DMVS_ERR_UNSUPPORTED;
    a.co_base = co_base;
    a.CC = CC;
    a.N = N;
    // straight-line epilogue: standard bias / residual / ReLU-on-all-channels arithmetic on whole 128-bit vectors
    a.fast_epi = d.epi == DMVS_EPI_STD && (d.act == DMVS_ACT_NONE || (d.act == DMVS_ACT_RELU && d.act_c0 <= 0)) && a.vec_y &&
                 (d.bias == nullptr || a.vec_bias) && (d.res_mode == DMVS_RES_NONE || a.vec_res) && co_base + CC <= d.Cout;
    const bool phase_launch = d.y_row_stride != 0 || d.res_row_stride != 0;
    if (phase_launch && !a.fast_epi) return DMVS_ERR_UNSUPPORTED;   // strided rows exist on the straight-line epilogue only
    a.y_rs = d.y_row_stride != 0 ? d.y_row_stride : d.Wo * d.y_ps;
    a.res_rs = d.res_row_stride != 0 ? d.res_row_stride : d.Wo * d.res_ps;
    a.TH = t.TH;
    a.TW = t.TW;
    a.in_rows = t.TH + (a.pair ? 2 * a.KHm : a.KHe) - 1;   // paired: the zero-weight half of the last pair reads real rows
    a.in_cols = t.in_cols;
    a.plane = t.plane;
    a.box_units = a.in_rows * a.in_cols;
    a.n_blk = t.n_blk;
    a.acc_cols = t.n_blk * N;
    a.R = t.R;
    a.stage_f = (a.pair ? 1 : 2) * t.plane * 4;
    a.wslab_f = a.KHm * 2 * N * 4;
    a.halo_f = t.halo_f;
    a.inv_in_cols = 1.0f / (float)t.in_cols;
    // packed slabs of this chunk: [hi | lo][KD][S*S phases][cin_pad/8][KHe][2][N][4]
    const int64_t plane_w = (int64_t)d.KD * a.S * a.S * (a.cin_pad >> 3) * a.KHm * 2 * N * 4;
    a.w_off = w_off;
    a.w_plane = plane_w;
    int cols = 32;
    while (cols < 2 * a.acc_cols) cols <<= 1;
    a.tmem_cols = cols;
    a.tiles_x = ceil_div(d.Wo, t.TW);
    a.tiles_y = ceil_div(d.Ho, t.TH);
    const long tiles = (long)a.tiles_x * a.tiles_y * d.N * d.Do;
    if (tiles > 0x7fffffffL) return DMVS_ERR_UNSUPPORTED;
    a.total_tiles = (int)tiles;
    a.dbg = plan_out != nullptr ? nullptr : ws2_debug_buffer();
    if (a.dbg != nullptr) cudaMemsetAsync(a.dbg, 0, sizeof(long long) * kDbgRoles * kDbgSlots, st);
    const int grid = (int)(tiles < (long)kNumSMs ? tiles : (long)kNumSMs);
    if (plan_out != nullptr) {
      if (n_launch < plan_cap) {
        int32_t* o = plan_out + 8 * n_launch;
        o[0] = CC; o[1] = N; o[2] = t.TH; o[3] = t.TW; o[4] = t.n_blk; o[5] = t.R; o[6] = 1; o[7] = (int32_t)t.smem;
      }
    } else {
      const int Hs = d.H, Ws = d.W;
      if (tensor_map_encoder() == nullptr) return DMVS_ERR_UNSUPPORTED;
      if (!make_activation_map(&a.map_x, d.x, d.C1, d.x_ps, Ws, Hs, d.D, d.N, 4, a.in_cols, a.in_rows, a.S)) return DMVS_ERR_UNSUPPORTED;
      if (d.C2 > 0) {
        if (!make_activation_map(&a.map_x2, d.x2, d.C2, d.x2_ps, Ws, Hs, d.D, d.N, 4, a.in_cols, a.in_rows, a.S))
          return DMVS_ERR_UNSUPPORTED;
      } else {
        a.map_x2 = a.map_x;
      }
      KernelFn fn = d.in_stats != nullptr ? get_kernel<true>() : get_kernel<false>();
      launch_pdl(fn, dim3(grid), dim3(kWs2Threads), t.smem, st, a);
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

int read_ws2_debug(long long* host_out, int count) {
  long long* buf = ws2_debug_buffer();
  if (buf == nullptr) return 0;
  const int n = count < kDbgRoles * kDbgSlots ? count : kDbgRoles * kDbgSlots;
  if (cudaMemcpy(host_out, buf, sizeof(long long) * n, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return n;
}

int dispatch_conv_ws2(const dmvs_conv_desc& d, cudaStream_t st) { return dispatch_conv_ws2(d, st, nullptr, 0); }
int plan_conv_ws2(const dmvs_conv_desc& d, int32_t* out, int cap) { return dispatch_conv_ws2(d, nullptr, out, cap); }

}  // namespace dmvs
