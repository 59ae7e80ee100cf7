// Tensor-core implicit-GEMM convolution (2-D / 3-D, channels-last), sharing the staged input tile and the
// fused epilogue with the FFMA back end (conv_common.cuh).
//
//   GEMM view per CTA:  M = 32 x TH output pixels,  N = COUT_S output channels,  K = taps x input channels.
//   * A (activations) is never im2col'ed: the halo tile sits in shared memory as [row][col][CKP] and
//     `ldmatrix.x4` reads the 16x8 fp32 fragment of 16 consecutive pixels at a tap offset (rows = pixels,
//     16-byte row segments = 4 channels) - conflict-free thanks to the padded pixel pitch.
//   * B (weights) is staged as [tap][cout][CKP] (input channel contiguous) and read with `ldmatrix` too.
//   * math: mma.sync.m16n8k8 TF32 with fp32 accumulation.  With passes == 3 every operand is split in
//     registers into hi = rna_tf32(x), lo = x - hi and the product is ah*bh + al*bh + ah*bl (error ~2^-21,
//     fp32 class), which keeps the <=1e-3 depth-parity bar with two orders of magnitude to spare; passes == 1
//     is the plain-TF32 mode (what cuDNN runs under torch defaults).
//   sm_100a note: legacy mma.sync peaks at 278 TFLOP/s TF32 on B200 (profiles/r1_mma_probe.txt), 4x the FFMA
//   pipe; the tcgen05 path is the next step for the FLOP-heaviest layers.
#include "conv_common.cuh"

namespace dmvs {
namespace {

__device__ __forceinline__ void ldmatrix_x4(unsigned (&r)[4], const float* smem_row) {
  const unsigned addr = static_cast<unsigned>(__cvta_generic_to_shared(smem_row));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2(unsigned (&r)[2], const float* smem_row) {
  const unsigned addr = static_cast<unsigned>(__cvta_generic_to_shared(smem_row));
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// x = hi + lo with hi = rna_tf32(x) and lo = rna_tf32(x - hi): both roundings are to nearest, so the
// dropped part (|x| * 2^-22) carries no systematic sign.
__device__ __forceinline__ void split_tf32(unsigned x, unsigned& hi, unsigned& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(hi) : "f"(__uint_as_float(x)));
  const float rest = __uint_as_float(x) - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(lo) : "f"(rest));
}

// One pipeline stage = one (output tile, depth tap kd, input-channel chunk c0).
struct Stage {
  int tile, kd, c0;
};

// NT = n8 tiles per warp, WC = warps along output channels, PX = output rows per warp, S = stride,
// PASSES = 1 (TF32) or 3 (3xTF32).
//
// Persistent CTAs walk the output tiles round-robin; the cp.async loads of stage s+1 (input halo tile and
// weight slab, double-buffered in shared memory) are in flight while the tensor cores work on stage s and
// while the epilogue of a finished tile streams out - loads, math and stores of one SM overlap without
// relying on a second resident CTA.
template <int NT, int WC, int PX, int S, int PASSES>
__global__ void __launch_bounds__(kConvThreads, 2) conv_mma_kernel(const __grid_constant__ ConvArgs a) {
  constexpr int WP = 8 / WC;
  constexpr int TH = WP * PX;
  constexpr int COUT_S = NT * 8 * WC;
  constexpr int OP = COUT_S + 4;
  const dmvs_conv_desc& d = a.d;

  extern __shared__ __align__(16) float smem[];
  const int in_sz = a.in_rows * a.in_cols * a.CKP;          // floats per input buffer
  const int w_sz = d.KH * d.KW * COUT_S * a.CKP;            // floats per weight buffer
  float* in_s0 = smem;                                      // [2][in_rows][in_cols][CKP]
  float* w_s0 = in_s0 + 2 * in_sz;                          // [2][KH*KW][COUT_S][CKP]
  float* out_s = w_s0 + 2 * w_sz;                           // [TH*32][OP]
  float* gn_s = out_s + TH * kTileW * OP;                   // [2][C1] when in_stats
  __shared__ unsigned long long stat_s[8];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wc = warp % WC, wp = warp / WC;
  const int tiles_x = ceil_div(d.Wo, kTileW), tiles_y = ceil_div(d.Ho, TH);
  const int total_tiles = tiles_x * tiles_y * d.N * d.Do;

  // ldmatrix source rows of this lane.  A: matrices (rows 0-7,k0-3) (rows 8-15,k0-3) (rows 0-7,k4-7) (rows 8-15,k4-7)
  const int lm = lane >> 3, lr = lane & 7;
  const int a_row = lr + (lm & 1) * 8;       // pixel within the m16 tile
  const int a_kofs = (lm >> 1) * 4;          // channel offset within the k8 step
  // B (x4): matrices (tile j,k0-3) (tile j,k4-7) (tile j+1,k0-3) (tile j+1,k4-7); (x2): first two only
  const int b_n = (lm >> 1) * 8 + lr;
  const int b_kofs = (lm & 1) * 4;
  const int ck4 = a.CK >> 2;
  const int w_units = d.KH * d.KW * COUT_S * ck4;   // 16-byte units of the weight slab
  const int ksteps = a.CK >> 3;
  const int row_pitch = S * a.in_cols * a.CKP;

  auto decode = [&](int tile, int& n, int& od, int& ty0, int& tx0) {
    const int tx = tile % tiles_x;
    const int r = tile / tiles_x;
    const int ty = r % tiles_y;
    const int z = r / tiles_y;
    n = z / d.Do;
    od = z - n * d.Do;
    ty0 = ty * TH;
    tx0 = tx * kTileW;
  };
  auto kd_first = [&](int od) { const int v = d.pad_d - od * S; return v > 0 ? v : 0; };
  auto kd_last = [&](int od) { const int v = d.D - 1 + d.pad_d - od * S; return v < d.KD - 1 ? v : d.KD - 1; };
  auto first_stage = [&](int tile) {
    Stage s{tile, 0, 0};
    if (tile < total_tiles) {
      int n, od, ty0, tx0;
      decode(tile, n, od, ty0, tx0);
      s.kd = kd_first(od);
    }
    return s;
  };
  auto advance = [&](const Stage& c) {
    Stage s = c;
    s.c0 += a.CK;
    if (s.c0 < a.cin_pad) return s;
    s.c0 = 0;
    int n, od, ty0, tx0;
    decode(c.tile, n, od, ty0, tx0);
    if (++s.kd <= kd_last(od)) return s;
    return first_stage(c.tile + (int)gridDim.x);
  };
  int gn_n = -1;
  auto issue = [&](const Stage& s, int buf) {
    int n, od, ty0, tx0;
    decode(s.tile, n, od, ty0, tx0);
    if (d.in_stats != nullptr && n != gn_n) {   // GroupNorm affine of the producer is per sample
      __syncthreads();
      for (int c = tid; c < d.C1; c += kConvThreads) groupnorm_affine(d, n, c, gn_s);
      __syncthreads();
      gn_n = n;
    }
    const int id = od * S + s.kd - d.pad_d;
    stage_input_tile(a, in_s0 + buf * in_sz, gn_s, n, id, ty0 * S - d.pad_h, tx0 * S - d.pad_w, s.c0);
    float* w_s = w_s0 + buf * w_sz;   // global [kd][tap][cout_pad8][cin_pad8] -> shared [tap][COUT_S][CKP]
#pragma unroll 1
    for (int idx = tid; idx < w_units; idx += kConvThreads) {
      const int c4 = idx & (ck4 - 1);
      const int r = idx >> a.ck4_shift;            // tap * COUT_S + co
      const int co = r % COUT_S;
      const int tap = r / COUT_S;
      const bool ok = s.c0 + c4 * 4 < a.cin_pad;
      const int64_t off =
          ((int64_t)(s.kd * d.KH * d.KW + tap) * a.w_cstride + a.co_base + co) * a.cin_pad + s.c0 + c4 * 4;
      cp_async16(w_s + r * a.CKP + c4 * 4, ok ? d.w_t + off : d.w_t, ok);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };

  float acc[PX][2][NT][4];
#pragma unroll
  for (int p = 0; p < PX; ++p)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[p][mt][j][e] = 0.0f;

  Stage cur = first_stage((int)blockIdx.x);
  if (cur.tile >= total_tiles) return;
  int buf = 0;
  issue(cur, 0);
  for (;;) {
    const Stage nxt = advance(cur);
    const bool has_next = nxt.tile < total_tiles;
    if (has_next) {
      issue(nxt, buf ^ 1);
      asm volatile("cp.async.wait_group 1;\n" ::: "memory");   // everything but the newest group has landed
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    __syncthreads();

    // ---- tensor-core math on stage `cur` -----------------------------------------------------------
    {
      const float* a_base = in_s0 + buf * in_sz + ((wp * PX * S) * a.in_cols + a_row * S) * a.CKP + a_kofs;
      const float* b_base = w_s0 + buf * w_sz + (wc * NT * 8 + b_n) * a.CKP + b_kofs;
#pragma unroll 1
      for (int kh = 0; kh < d.KH; ++kh) {
#pragma unroll 1
        for (int kw = 0; kw < d.KW; ++kw) {
          const float* ap = a_base + (kh * a.in_cols + kw) * a.CKP;
          const float* bp = b_base + ((kh * d.KW + kw) * COUT_S) * a.CKP;
#pragma unroll 1
          for (int ks = 0; ks < ksteps; ++ks) {
            unsigned bh[NT][2], bl[NT][2];
#pragma unroll
            for (int j = 0; j < NT; j += 2) {
              if (j + 1 < NT) {
                unsigned r4[4];
                ldmatrix_x4(r4, bp + j * 8 * a.CKP + ks * 8);
                bh[j][0] = r4[0]; bh[j][1] = r4[1]; bh[j + 1][0] = r4[2]; bh[j + 1][1] = r4[3];
              } else {
                unsigned r2[2];
                ldmatrix_x2(r2, bp + j * 8 * a.CKP + ks * 8);
                bh[j][0] = r2[0]; bh[j][1] = r2[1];
              }
            }
            if (PASSES == 3) {
#pragma unroll
              for (int j = 0; j < NT; ++j) {
                split_tf32(bh[j][0], bh[j][0], bl[j][0]);
                split_tf32(bh[j][1], bh[j][1], bl[j][1]);
              }
            }
#pragma unroll
            for (int p = 0; p < PX; ++p) {
              // both m16 tiles of the row are loaded up front so their MMA chains interleave
              unsigned ah[2][4];
              ldmatrix_x4(ah[0], ap + p * row_pitch + ks * 8);
              ldmatrix_x4(ah[1], ap + p * row_pitch + 16 * S * a.CKP + ks * 8);
              if (PASSES == 3) {
                unsigned al[2][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                  for (int e = 0; e < 4; ++e) split_tf32(ah[mt][e], ah[mt][e], al[mt][e]);
                // The tensor core adds into its accumulator with truncation; chaining hundreds of k-steps
                // through it biases long reductions (7x7x64: ~2e-5).  Each k8 step is therefore formed from a
                // zero accumulator and folded into the running sum with a round-to-nearest FADD.
                float t4[2][NT][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                  for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) t4[mt][j][e] = 0.0f;
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                  for (int j = 0; j < NT; ++j) mma_tf32(t4[mt][j], al[mt], bh[j][0], bh[j][1]);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                  for (int j = 0; j < NT; ++j) mma_tf32(t4[mt][j], ah[mt], bl[j][0], bl[j][1]);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                  for (int j = 0; j < NT; ++j) mma_tf32(t4[mt][j], ah[mt], bh[j][0], bh[j][1]);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                  for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[p][mt][j][e] += t4[mt][j][e];
              } else {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                  for (int j = 0; j < NT; ++j) mma_tf32(acc[p][mt][j], ah[mt], bh[j][0], bh[j][1]);
              }
            }
          }
        }
      }
    }

    // ---- tile finished: accumulators -> shared output tile -> fused epilogue --------------------------
    if (!has_next || nxt.tile != cur.tile) {
      int n, od, ty0, tx0;
      decode(cur.tile, n, od, ty0, tx0);
      const int g = lane >> 2, t = lane & 3;   // C fragment: rows g / g+8, columns 2t, 2t+1
#pragma unroll
      for (int p = 0; p < PX; ++p)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            float* o = out_s + ((wp * PX + p) * kTileW + mt * 16 + g) * OP + (wc * NT + j) * 8 + 2 * t;
            *reinterpret_cast<float2*>(o) = make_float2(acc[p][mt][j][0], acc[p][mt][j][1]);
            *reinterpret_cast<float2*>(o + 8 * OP) = make_float2(acc[p][mt][j][2], acc[p][mt][j][3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[p][mt][j][e] = 0.0f;
          }
      if (tid < 8) stat_s[tid] = 0ull;
      __syncthreads();
      epilogue_tile<TH, COUT_S>(a, out_s, stat_s, n, od, ty0, tx0);
    }
    __syncthreads();   // every warp is done with this stage's buffers (and out_s) before they are refilled
    if (!has_next) break;
    cur = nxt;
    buf ^= 1;
  }
}

using KernelFn = void (*)(const ConvArgs);

template <int NT, int WC, int PX, int S, int PASSES>
KernelFn get_kernel() {
  static SmemOptIn opt_in;
  KernelFn fn = conv_mma_kernel<NT, WC, PX, S, PASSES>;
  opt_in.ensure(fn, 200 * 1024);
  return fn;
}

template <int NT, int WC, int MAXPX, int PASSES>
KernelFn pick_px(int px, int s) {
  if (px > MAXPX) px = MAXPX;
  if (s == 1) {
    if (px >= 4) { if constexpr (MAXPX >= 4) return get_kernel<NT, WC, 4, 1, PASSES>(); }
    if (px >= 2) return get_kernel<NT, WC, 2, 1, PASSES>();
    return get_kernel<NT, WC, 1, 1, PASSES>();
  }
  if (px >= 4) { if constexpr (MAXPX >= 4) return get_kernel<NT, WC, 4, 2, PASSES>(); }
  if (px >= 2) return get_kernel<NT, WC, 2, 2, PASSES>();
  return get_kernel<NT, WC, 1, 2, PASSES>();
}

struct Config { int wc, max_px; };
Config config_of(int chunk) {
  switch (chunk) {
    case 8: return {1, 4};
    case 16: return {1, 4};
    case 32: return {1, 2};
    case 64: return {2, 2};
    default: return {4, 2};
  }
}
template <int PASSES>
KernelFn pick_kernel_p(int chunk, int px, int s) {
  switch (chunk) {
    case 8: return pick_px<1, 1, 4, PASSES>(px, s);
    case 16: return pick_px<2, 1, 4, PASSES>(px, s);
    case 32: return pick_px<4, 1, 2, PASSES>(px, s);
    case 64: return pick_px<4, 2, 2, PASSES>(px, s);
    default: return pick_px<4, 4, 2, PASSES>(px, s);
  }
}
KernelFn pick_kernel(int chunk, int px, int s, int passes) {
  return passes == 3 ? pick_kernel_p<3>(chunk, px, s) : pick_kernel_p<1>(chunk, px, s);
}

}  // namespace

int dispatch_conv_mma(const dmvs_conv_desc& d, cudaStream_t st) {
  if (!aligned16(d.w_t)) return DMVS_ERR_ALIGN;
  ConvArgs a;
  a.d = d;
  a.cin_pad = (d.C1 + d.C2 + 7) & ~7;
  a.w_cstride = (d.Cout + 7) & ~7;   // padded Cout of the transposed weights
  const bool vec_x = aligned16(d.x) && (d.x_ps % 4 == 0) && (d.C1 % 4 == 0);
  const bool vec_x2 = d.C2 == 0 || (aligned16(d.x2) && (d.x2_ps % 4 == 0) && (d.C2 % 4 == 0));
  a.fast_in = vec_x && vec_x2 && d.in_stats == nullptr;
  a.vec_y = aligned16(d.y) && (d.y_ps % 4 == 0);
  a.vec_res = d.res != nullptr && aligned16(d.res) && (d.res_ps % 4 == 0);
  a.Hs = d.in_up2 ? d.H / 2 : d.H;
  a.Ws = d.in_up2 ? d.W / 2 : d.W;
  a.passes = d.precision == DMVS_PREC_TF32 ? 1 : 3;

  const int S = d.stride;
  int remaining = a.w_cstride;
  int co_base = 0;
  while (remaining > 0) {
    int chunk = 128;
    while (chunk > remaining) chunk >>= 1;   // power of two >= 8
    // tile search: prefer two CTAs per SM (100 KB), fall back to one CTA (200 KB), then to a narrower chunk
    int px = 0, ck = 0;
    size_t smem = 0;
    Config cfg = config_of(chunk);
    for (;;) {
      cfg = config_of(chunk);
      for (size_t budget = kSmemBudget; budget <= 2 * (size_t)kSmemBudget && !ck; budget += kSmemBudget) {
        for (px = cfg.max_px; px >= 1; px >>= 1) {
          const int th = (8 / cfg.wc) * px;
          const int in_rows = (th - 1) * S + d.KH;
          const int in_cols = (kTileW - 1) * S + d.KW;
          ck = 0;
          for (int c = 16; c >= 8; c >>= 1) {
            if (c > a.cin_pad) continue;
            const int ckp = c + 4;
            // double-buffered input tile + weight slab, output tile, GroupNorm affine
            const size_t need = (2 * ((size_t)in_rows * in_cols * ckp + (size_t)d.KH * d.KW * chunk * ckp) +
                                 (size_t)th * kTileW * (chunk + 4) + 2 * (size_t)d.C1) * 4;
            if (need <= budget) { ck = c; smem = need; break; }
          }
          if (ck) {
            const long blocks = (long)ceil_div(d.Wo, kTileW) * ceil_div(d.Ho, th) * d.N * d.Do;
            if (blocks >= 2 * kNumSMs || px == 1) break;
          }
        }
      }
      if (ck || chunk == 8) break;
      chunk >>= 1;
    }
    if (!ck) return DMVS_ERR_UNSUPPORTED;
    if (px < 1) px = 1;
    KernelFn fn = pick_kernel(chunk, px, S, a.passes);
    const int th = (8 / cfg.wc) * px;
    a.co_base = co_base;
    a.CK = ck;
    a.CKP = ck + 4;
    a.ck4_shift = ck == 8 ? 1 : 2;
    a.in_rows = (th - 1) * S + d.KH;
    a.in_cols = (kTileW - 1) * S + d.KW;
    const long tiles = (long)ceil_div(d.Wo, kTileW) * ceil_div(d.Ho, th) * d.N * d.Do;
    if (tiles > 0x7fffffffL) return DMVS_ERR_UNSUPPORTED;
    const int ctas_per_sm = smem <= (size_t)kSmemBudget ? 2 : 1;
    const int grid = (int)(tiles < (long)kNumSMs * ctas_per_sm ? tiles : (long)kNumSMs * ctas_per_sm);
    fn<<<grid, kConvThreads, smem, st>>>(a);
    const int rc = launch_status();
    if (rc) return rc;
    co_base += chunk;
    remaining -= chunk;
  }
  return 0;
}

}  // namespace dmvs
