// Throughput probe: legacy mma.sync (tf32 m16n8k8, bf16 m16n8k16) and FFMA on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_mma_tf32(float* out, int iters) {
  float c[8][4] = {};
  unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f800000u, 0x3f900000u, 0x3fa00000u}, b[2] = {0x3f800000u, 0x3f800001u + threadIdx.x};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mma_bf16(float* out, int iters) {
  float c[8][4] = {};
  unsigned a[4] = {0x3f803f80u + threadIdx.x, 0x3f803f80u, 0x3f903f80u, 0x3fa03f80u}, b[2] = {0x3f803f80u, 0x3f803f81u + threadIdx.x};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0; for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out, int iters, float x, float y) {
  float c[32];
  for (int j = 0; j < 32; ++j) c[j] = threadIdx.x + j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 32; ++j) c[j] = fmaf(c[j], x, y);
  }
  float s = 0; for (int j = 0; j < 32; ++j) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    int blocks = 148 * 2, threads = warps * 32 / 2;
    if (threads > 1024) continue;
    float ms = timeit([&] { k_mma_tf32<<<blocks, threads>>>(out, iters); });
    double macs = (double)blocks * (threads / 32) * iters * 8 * 16 * 8 * 8;
    printf("mma.sync tf32 m16n8k8 : %2d warps/SM  %8.1f TFLOP/s\n", warps, 2 * macs / ms / 1e9);
    ms = timeit([&] { k_mma_bf16<<<blocks, threads>>>(out, iters); });
    macs = (double)blocks * (threads / 32) * iters * 8 * 16 * 8 * 16;
    printf("mma.sync bf16 m16n8k16: %2d warps/SM  %8.1f TFLOP/s\n", warps, 2 * macs / ms / 1e9);
    ms = timeit([&] { k_ffma<<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
    double fl = (double)blocks * threads * iters * 32 * 2;
    printf("ffma                  : %2d warps/SM  %8.1f TFLOP/s\n", warps, fl / ms / 1e9);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
