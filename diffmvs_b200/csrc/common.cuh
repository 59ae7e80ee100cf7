// Shared device helpers for the diffmvs_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "diffmvs_b200.h"

#define DMVS_STR2(x) #x
#define DMVS_STR(x) DMVS_STR2(x)

namespace dmvs {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }
__device__ __forceinline__ float siluf_(float v) { return v / (1.0f + expf(-v)); }

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case DMVS_ACT_RELU: return fmaxf(v, 0.0f);
    case DMVS_ACT_SIGMOID: return sigmoidf_(v);
    case DMVS_ACT_TANH: return tanhf(v);
    case DMVS_ACT_SILU: return siluf_(v);
    default: return v;
  }
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// GroupNorm statistics (sum, sum of squares per sample and group) are accumulated as 64-bit FIXED-POINT integers
// (2^-20 units): integer addition is associative, so the result does not depend on the order in which warps, CTAs and
// tiles arrive - two runs are bit-identical, like the reference's single-kernel GroupNorm.  Partials are rounded to
// nearest (unbiased); 2^-20 resolution on per-warp partials of magnitude >= 1 keeps the relative error of the
// statistics below 1e-7, and the range (8.8e12) covers activations up to ~1e3 RMS on the largest maps.
constexpr double kStatScale = 1048576.0;
__device__ __forceinline__ unsigned long long stat_fixed(float v) {
  return (unsigned long long)__double2ll_rn((double)v * kStatScale);
}
__device__ __forceinline__ double stat_value(long long v) { return (double)v * (1.0 / kStatScale); }

// disp_to_depth / depth_to_disp of the reference (module.py:220-235), evaluated from the same
// `depth_min`, `depth_max` tensors the reference passes around (diffusion.py:140-146).
struct DepthRange {
  float min_disp, span;
  __device__ __forceinline__ DepthRange(float depth_min, float depth_max) {
    min_disp = 1.0f / depth_max;
    const float max_disp = 1.0f / depth_min;
    span = max_disp - min_disp;
  }
  __device__ __forceinline__ float to_depth(float n) const {
    // __fmaf_rn would fuse the rounding; the reference rounds the product first.
    const float scaled = fmaxf(__fadd_rn(min_disp, __fmul_rn(span, n)), 1e-6f);
    return 1.0f / scaled;
  }
  __device__ __forceinline__ float to_norm(float depth) const {
    return __fdiv_rn(__fsub_rn(1.0f / depth, min_disp), span);
  }
};

// Number of kernels this library has launched in this process (bench.py reports it as gpu_launches).
extern unsigned long long g_launch_count;

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting: remember per kernel AND per device that the
// opt-in happened (a process may drive several GPUs, e.g. torch tensors on cuda:1 while cuda:0 is also in use).
struct SmemOptIn {
  bool done[64] = {};
  template <typename Fn>
  void ensure(Fn fn, int bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0, done[0] = false;
    if (!done[dev]) {
      cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      done[dev] = true;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch: a kernel launched through launch_pdl may become resident while its
// predecessor on the stream is still draining; its on-chip prologue (barrier init, TMEM allocation,
// descriptor prefetch, index arithmetic) then overlaps the predecessor's tail and the launch latency.
// Contract: such a kernel calls pdl_sync() before its FIRST access to global memory, and nothing
// before that point reads or writes global memory.  griddepcontrol.wait returns only when every
// prerequisite grid has completed and its writes are visible, so no ordering between kernels changes.
// DMVS_PDL=0 launches everything fully serialised (A/B switch).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("DMVS_PDL");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(static_cast<Args&&>(args))...);
}

inline int launch_status() {
  ++g_launch_count;
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

}  // namespace dmvs
