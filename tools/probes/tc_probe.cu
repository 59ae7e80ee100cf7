// tcgen05 probe: validates the shared-memory descriptor encoding used by the implicit-GEMM convolution:
//   A = 128 consecutive "pixels" x K channels read from a planar-by-channel-quad buffer A_s[q][row][4 floats]
//       (K-major, no swizzle: 8 rows x 16 B core matrices, SBO = 128 B, LBO = plane pitch) at an arbitrary
//       16-byte-aligned row shift (the tap offset of the convolution),
//   B = N x K weights as B_s[q][n][4 floats].
// D (TMEM, fp32) is read back with tcgen05.ld and compared on the host.  Inputs are exactly representable in
// TF32 so the expected result is exact.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= 1ull << 46;   // descriptor version (Blackwell)
  return d;          // layout_type = 0 (no swizzle), base_offset = 0
}

template <int N>
__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int rows, int K, int shift, int* tmem_out) {
  extern __shared__ __align__(128) float smem[];
  const int quads = K / 4;
  float* A_s = smem;                       // [quads][rows][4]
  float* B_s = A_s + quads * rows * 4;     // [quads][N][4]
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
  }
  // A global [rows][K] -> A_s[q][row][4];  B global [N][K] -> B_s[q][n][4]
  for (int i = tid; i < rows * quads; i += 128) {
    const int q = i % quads, r = i / quads;
    *reinterpret_cast<float4*>(A_s + (q * rows + r) * 4) = *reinterpret_cast<const float4*>(A + r * K + q * 4);
  }
  for (int i = tid; i < N * quads; i += 128) {
    const int q = i % quads, n = i / quads;
    *reinterpret_cast<float4*>(B_s + (q * N + n) * 4) = *reinterpret_cast<const float4*>(B + n * K + q * 4);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tbase = tmem_base;
  if (tid == 0) *tmem_out = (int)tbase;

  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint64_t da = make_desc(smem_u32(A_s + ((2 * ks) * rows + shift) * 4), rows * 16, 128);
      const uint64_t db = make_desc(smem_u32(B_s + ((2 * ks) * N) * 4), N * 16, 128);
      const uint32_t accum = ks > 0 ? 1u : 0u;
      asm volatile(
          "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tbase),
          "l"(da), "l"(db), "r"(idesc), "r"(accum)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar)) : "memory");
  }
  // wait for the MMAs
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // each warp reads its 32 lanes; 16 columns per load
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tbase), "r"(64));
}

template <int N>
int run(int K, int shift) {
  const int rows = 128 + 32;
  std::vector<float> A(rows * K), B(N * K), D(128 * N), R(128 * N);
  srand(1234 + N + K + shift);
  for (auto& v : A) v = (float)((rand() % 17) - 8) / 8.0f;
  for (auto& v : B) v = (float)((rand() % 13) - 6) / 4.0f;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[(m + shift) * K + k] * B[n * K + k];
      R[m * N + n] = (float)s;
    }
  float *dA, *dB, *dD; int* dT;
  CHECK(cudaMalloc(&dA, A.size() * 4)); CHECK(cudaMalloc(&dB, B.size() * 4)); CHECK(cudaMalloc(&dD, D.size() * 4)); CHECK(cudaMalloc(&dT, 4));
  CHECK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CHECK(cudaMemset(dD, 0xff, D.size() * 4));
  const size_t smem = (size_t)(K / 4) * (rows + N) * 16;
  CHECK(cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe<N><<<1, 128, smem>>>(dA, dB, dD, rows, K, shift, dT);
  CHECK(cudaDeviceSynchronize());
  CHECK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  int tb; CHECK(cudaMemcpy(&tb, dT, 4, cudaMemcpyDeviceToHost));
  double maxerr = 0; int bad = 0;
  for (int i = 0; i < 128 * N; ++i) { double e = fabs((double)D[i] - R[i]); if (e > maxerr) maxerr = e; if (e > 1e-4) ++bad; }
  printf("N=%3d K=%3d shift=%2d tmem_base=0x%x  max|err|=%.3e  mismatches=%d/%d  D[0]=%g R[0]=%g D[last]=%g R[last]=%g\n", N, K, shift, tb,
         maxerr, bad, 128 * N, D[0], R[0], D[128 * N - 1], R[128 * N - 1]);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dT);
  return bad;
}

int main() {
  int bad = 0;
  bad += run<16>(8, 0);
  bad += run<16>(64, 0);
  bad += run<16>(64, 3);
  bad += run<32>(32, 1);
  bad += run<64>(64, 5);
  printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
  return bad != 0;
}
