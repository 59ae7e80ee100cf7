// Pipeline probe for the round-2 convolution kernel (sm_100a).  Three questions, each answered by a measurement:
//   1. tcgen05.mma SS-mode cost per instruction (M=128, K=8 tf32, un-swizzled K-major A/B as in conv_ws.cu) versus N:
//      is it max(32, N/2) clk, and do B reads add?
//   2. cp.async.bulk.tensor throughput for channels-last activation tiles with inner boxes of 16 B (planar-by-quad,
//      un-swizzled), 32 B (SWIZZLE_32B), 64 B (SWIZZLE_64B) and 128 B (SWIZZLE_128B), streaming from HBM.
//   3. Does a K-major SWIZZLE_{32,64,128}B operand written by TMA feed tcgen05.mma correctly when the descriptor start
//      address is offset by an ARBITRARY number of rows (the kernel-row tap of the convolution), and by 32-byte K slices?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu     (no -lcuda needed)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); exit(1); }
  return (EncodeTiledFn)fn;
}

// channels-last map [N][H][W][C] fp32 -> 4-D tensor map (C, W, H, N), box (bc, bw, bh, 1)
static CUtensorMap make_map(EncodeTiledFn enc, const float* base, int N, int H, int W, int C, int ps, int bc, int bw, int bh,
                            CUtensorMapSwizzle swz) {
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ps * 4, (cuuint64_t)W * ps * 4, (cuuint64_t)H * W * ps * 4};
  cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d (bc=%d swz=%d)\n", (int)r, bc, (int)swz); exit(1); }
  return m;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 26)) { printf("mbar timeout\n"); __trap(); }
  }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(
                   smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t v = 0;
  v |= (uint64_t)((saddr >> 4) & 0x3fff);
  v |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  v |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  v |= 1ull << 46;
  v |= (uint64_t)layout << 61;
  return v;
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
               "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t make_idesc(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------------------------
// 1. MMA issue rate
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) mma_rate_kernel(int N, int iters, int taps, int in_cols, int tmem_cols, long long* cycles_out) {
  extern __shared__ __align__(1024) float smem[];
  const int plane = 2048;                     // positions per quad plane
  float* A_s = smem;                          // [2][plane][4]
  float* B_s = smem + 2 * plane * 4;          // [taps][2][N][4]
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * plane * 4 + taps * 2 * N * 4; i += 128) smem[i] = (float)((i * 37) % 19 - 9) * 0.125f;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n"); }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tb = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(N);
    const uint64_t da0 = umma_desc(smem_u32(A_s), plane * 16, 128, 0), db0 = umma_desc(smem_u32(B_s), N * 16, 128, 0);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int blk = it & 3;
      for (int t = 0; t < taps; ++t)
        umma_tf32(tb + (uint32_t)((blk & 1) * N), da0 + (uint32_t)(blk * 128 + t * in_cols), db0 + (uint32_t)(t * 2 * N), idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) *cycles_out = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tb), "r"(tmem_cols));
}

// 1b. the same with consecutive MMAs spread round-robin over G accumulators (independent chains): separates a
// dependent-accumulation latency from a per-instruction issue floor
__global__ void __launch_bounds__(128) mma_chain_kernel(int N, int iters, int G, int kslices, long long* cycles_out) {
  extern __shared__ __align__(1024) float smem[];
  const int plane = 2048;
  float* A_s = smem;                          // [2*kslices][plane][4]
  float* B_s = smem + 2 * kslices * plane * 4;   // [2*kslices][N][4]
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * kslices * (plane + N) * 4; i += 128) smem[i] = (float)((i * 37) % 19 - 9) * 0.125f;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n"); }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tb = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(N);
    const uint64_t da0 = umma_desc(smem_u32(A_s), plane * 16, 128, 0), db0 = umma_desc(smem_u32(B_s), N * 16, 128, 0);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int g = it % G;
      for (int ks = 0; ks < kslices; ++ks)   // kslices back-to-back K=8 MMAs into the same accumulator (a K=8*kslices stage)
        umma_tf32(tb + (uint32_t)(g * N), da0 + (uint32_t)(ks * 2 * plane + (it & 7) * 128), db0 + (uint32_t)(ks * 2 * N), idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) *cycles_out = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tb), "r"(512));
}

// ------------------------------------------------------------------------------------------------------------------
// 2. TMA tile streaming
// ------------------------------------------------------------------------------------------------------------------
constexpr int kRing = 2;
__global__ void __launch_bounds__(64) tma_stream_kernel(const __grid_constant__ CUtensorMap map, int C, int bc, int TH, int TW, int in_rows,
                                                        int in_cols, int tiles_x, int tiles_y, int total_tiles, float* sink) {
  extern __shared__ __align__(1024) float smem[];
  __shared__ __align__(8) uint64_t full[kRing], empty[kRing];
  const int tid = threadIdx.x;
  const int tile_bytes = in_rows * in_cols * C * 4;
  const int slot_f = ((tile_bytes + 1023) & ~1023) / 4;
  if (tid == 0) {
    for (int i = 0; i < kRing; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;\n");
  }
  __syncthreads();
  if (tid == 0) {   // producer
    int slot = 0, use = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1);
      const int tx = tile % tiles_x, r = tile / tiles_x, ty = r % tiles_y, n = r / tiles_y;
      mbar_expect_tx(&full[slot], tile_bytes);
      float* dst = smem + slot * slot_f;
      for (int c = 0; c < C; c += bc)
        tma_load_4d(dst + (c / bc) * in_rows * in_cols * bc, &map, &full[slot], c, tx * TW - 1, ty * TH - 1, n);
      if (++slot == kRing) { slot = 0; ++use; }
    }
  } else if (tid == 32) {   // consumer: touches one word per tile, frees the slot
    int slot = 0, use = 0;
    float acc = 0.f;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&full[slot], use & 1);
      acc += smem[slot * slot_f + 5];
      mbar_arrive(&empty[slot]);
      if (++slot == kRing) { slot = 0; ++use; }
    }
    if (acc == 123.456f) *sink = acc;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// 3. swizzled K-major operand written by TMA, consumed at a row offset
// ------------------------------------------------------------------------------------------------------------------
// A: tensor [1][H=rows][W=1][C=K] loaded as a box (K, 1.., rows) with swizzle `swz` (inner bytes K*4 = 32/64/128);
// B: un-swizzled [quad][N][4] written by threads.  D[m][n] = sum_k A[m + shift][k] * B[n][k].
__global__ void __launch_bounds__(128) swz_mma_kernel(const __grid_constant__ CUtensorMap map, const float* Bg, float* D, int rows, int K, int N,
                                                      int shift, int layout, int sbo, int base_off_mode) {
  extern __shared__ __align__(1024) float smem[];
  float* A_s = smem;                         // rows * K floats, swizzled by TMA
  float* B_s = smem + ((rows * K + 255) & ~255);
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar, mbar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quads = K / 4;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;\n"); }
  for (int i = tid; i < N * quads; i += 128) {
    const int q = i % quads, n = i / quads;
    *reinterpret_cast<float4*>(B_s + (q * N + n) * 4) = *reinterpret_cast<const float4*>(Bg + n * K + q * 4);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tb = tmem_base;
  if (tid == 0) {
    mbar_expect_tx(&bar, rows * K * 4);
    tma_load_4d(A_s, &map, &bar, 0, 0, 0, 0);
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t idesc = make_idesc(N);
    const uint32_t row_bytes = K * 4;
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint32_t a_addr = smem_u32(A_s) + shift * row_bytes + ks * 32;
      uint64_t da = umma_desc(a_addr, 16, sbo, layout);
      if (base_off_mode) da |= (uint64_t)((a_addr >> 7) & 7) << 49;
      const uint64_t db = umma_desc(smem_u32(B_s + (2 * ks) * N * 4), N * 16, 128, 0);
      umma_tf32(tb, da, db, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(&mbar);
  }
  mbar_wait(&mbar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tb), "r"(256));
}

static int run_swz(EncodeTiledFn enc, int K, int N, int shift, int base_off_mode) {
  const int rows = 128 + 72;
  std::vector<float> A(rows * K), B(N * K), D(128 * N), R(128 * N);
  srand(77 + K + N + shift);
  for (auto& v : A) v = (float)((rand() % 17) - 8) / 8.0f;
  for (auto& v : B) v = (float)((rand() % 13) - 6) / 4.0f;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[(m + shift) * K + k] * B[n * K + k];
      R[m * N + n] = (float)s;
    }
  float *dA, *dB, *dD;
  CHECK(cudaMalloc(&dA, A.size() * 4)); CHECK(cudaMalloc(&dB, B.size() * 4)); CHECK(cudaMalloc(&dD, D.size() * 4));
  CHECK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CHECK(cudaMemset(dD, 0xff, D.size() * 4));
  const CUtensorMapSwizzle swz = K == 8 ? CU_TENSOR_MAP_SWIZZLE_32B : (K == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
  const int layout = K == 8 ? 6 : (K == 16 ? 4 : 2);
  // tensor [N=1][H=1][W=rows][C=K]: box (K, rows<=256, 1, 1)
  CUtensorMap map = make_map(enc, dA, 1, 1, rows, K, K, K, rows, 1, swz);
  const size_t smem = (size_t)((rows * K + 255) & ~255) * 4 + (size_t)(K / 4) * N * 16 + 1024;
  CHECK(cudaFuncSetAttribute(swz_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  swz_mma_kernel<<<1, 128, smem>>>(map, dB, dD, rows, K, N, shift, layout, 8 * K * 4, base_off_mode);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("swz K=%d shift=%d: CUDA error %s\n", K, shift, cudaGetErrorString(e)); exit(1); }
  CHECK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0; double maxerr = 0;
  for (int i = 0; i < 128 * N; ++i) { double er = fabs((double)D[i] - R[i]); if (er > maxerr) maxerr = er; if (!(er <= 1e-4)) ++bad; }
  printf("swizzle K=%2d (%3d-byte rows) N=%3d row-shift=%3d base_off=%d : mismatches %d/%d max|err| %.3g\n", K, K * 4, N, shift,
         base_off_mode, bad, 128 * N, maxerr);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return bad;
}

int main() {
  EncodeTiledFn enc = get_encode();
  cudaDeviceProp prop;
  CHECK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs\n", prop.name, prop.multiProcessorCount);
  const int sms = prop.multiProcessorCount;

  // ---- 3. swizzle + row offsets (first: a failure here changes the plan) ----
  int bad = 0;
  for (int K : {8, 16, 32})
    for (int shift : {0, 1, 5, 8, 64, 66, 67})
      bad += run_swz(enc, K, 48, shift, 0) != 0;
  if (bad) {
    printf("-- some swizzled row offsets FAILED with base_offset = 0; retry with base_offset = (addr >> 7) & 7\n");
    for (int K : {8, 16, 32})
      for (int shift : {1, 5, 66, 67}) run_swz(enc, K, 48, shift, 1);
  }

  // ---- 1. MMA cost per instruction ----
  long long* d_cyc;
  CHECK(cudaMalloc(&d_cyc, 8));
  CHECK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int ctas : {1, 2}) {
    for (int N : {16, 32, 48, 64, 96, 128, 192, 256}) {
      const int taps = 3, iters = 400;
      const size_t smem = (size_t)(2 * 2048 * 4 + taps * 2 * N * 4) * 4;
      if (ctas == 2 && (smem > 100 * 1024 || N > 128)) continue;
      mma_rate_kernel<<<sms * ctas, 128, smem>>>(N, iters, taps, 64, ctas == 2 ? 256 : 512, d_cyc);
      CHECK(cudaDeviceSynchronize());
      long long cyc;
      CHECK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
      printf("mma M128 N%3d K8 tf32 SS, %d CTA/SM: %.1f clk per MMA (CTA 0)  [A read 4096 B + B read %d B]\n", N, ctas,
             (double)cyc / (iters * taps), N * 32);
    }
  }

  // ---- 1b. dependent chain vs independent accumulators ----
  CHECK(cudaFuncSetAttribute(mma_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int N : {32, 48, 96}) {
    for (int kslices : {1, 2}) {
      for (int G : {1, 2, 3, 4, 8}) {
        if (G * N > 512) continue;
        const int iters = 1200;
        const size_t smem = (size_t)2 * kslices * (2048 + N) * 16;
        mma_chain_kernel<<<sms, 128, smem>>>(N, iters, G, kslices, d_cyc);
        CHECK(cudaDeviceSynchronize());
        long long cyc;
        CHECK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
        printf("mma chain N%3d, %d K-slices per accumulator visit, %d accumulators round-robin: %.1f clk per MMA\n", N, kslices, G,
               (double)cyc / (iters * kslices));
      }
    }
  }

  // ---- 2. TMA streaming ----
  {
    const int Nimg = 7, H = 576, W = 800;
    float* sink;
    CHECK(cudaMalloc(&sink, 4));
    CHECK(cudaFuncSetAttribute(tma_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    struct Mode { int C, bc; CUtensorMapSwizzle swz; const char* name; };
    const Mode modes[] = {{16, 4, CU_TENSOR_MAP_SWIZZLE_NONE, "C16 box 4ch (16 B inner, planar quads)"},
                          {16, 8, CU_TENSOR_MAP_SWIZZLE_32B, "C16 box 8ch (32 B inner, SW32)"},
                          {16, 16, CU_TENSOR_MAP_SWIZZLE_64B, "C16 box 16ch (64 B inner, SW64)"},
                          {32, 4, CU_TENSOR_MAP_SWIZZLE_NONE, "C32 box 4ch (16 B inner)"},
                          {32, 8, CU_TENSOR_MAP_SWIZZLE_32B, "C32 box 8ch (32 B inner, SW32)"},
                          {32, 32, CU_TENSOR_MAP_SWIZZLE_128B, "C32 box 32ch (128 B inner, SW128)"},
                          {64, 4, CU_TENSOR_MAP_SWIZZLE_NONE, "C64 box 4ch (16 B inner)"},
                          {64, 8, CU_TENSOR_MAP_SWIZZLE_32B, "C64 box 8ch (32 B inner, SW32)"},
                          {64, 32, CU_TENSOR_MAP_SWIZZLE_128B, "C64 box 32ch (128 B inner, SW128)"}};
    for (const Mode& m : modes) {
      float* x;
      const size_t elems = (size_t)Nimg * H * W * m.C;
      CHECK(cudaMalloc(&x, elems * 4));
      CHECK(cudaMemset(x, 0, elems * 4));
      for (int TH : {8, 16}) {
        const int TW = 62, in_rows = TH + 2, in_cols = TW + 2;
        CUtensorMap map = make_map(enc, x, Nimg, H, W, m.C, m.C, m.bc, in_cols, in_rows, m.swz);
        const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH, total = tiles_x * tiles_y * Nimg;
        const int tile_bytes = in_rows * in_cols * m.C * 4;
        const size_t smem = (size_t)kRing * ((tile_bytes + 1023) & ~1023) + 1024;
        if (smem > 200 * 1024) continue;
        for (int ctas : {1, 2}) {
          if (ctas * smem > 220 * 1024) continue;
          cudaEvent_t e0, e1;
          CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
          tma_stream_kernel<<<sms * ctas, 64, smem>>>(map, m.C, m.bc, TH, TW, in_rows, in_cols, tiles_x, tiles_y, total, sink);
          CHECK(cudaDeviceSynchronize());
          CHECK(cudaEventRecord(e0));
          tma_stream_kernel<<<sms * ctas, 64, smem>>>(map, m.C, m.bc, TH, TW, in_rows, in_cols, tiles_x, tiles_y, total, sink);
          CHECK(cudaEventRecord(e1));
          CHECK(cudaDeviceSynchronize());
          float ms;
          CHECK(cudaEventElapsedTime(&ms, e0, e1));
          const double useful = (double)elems * 4, moved = (double)total * tile_bytes;
          printf("tma %-40s tile %2dx%d, %d CTA/SM: %.3f ms  %.0f GB/s unique  %.0f GB/s incl. halo  (%d boxes/tile)\n", m.name, TH, TW,
                 ctas, ms, useful / ms * 1e-6, moved / ms * 1e-6, m.C / m.bc);
        }
      }
      cudaFree(x);
    }
  }
  printf("PROBE DONE\n");
  return 0;
}
