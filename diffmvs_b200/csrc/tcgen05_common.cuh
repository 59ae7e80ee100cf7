// Inline-PTX helpers shared by the tcgen05 / TMEM / TMA kernels (sm_100a): shared-memory matrix descriptors, MMA issue,
// TMEM loads, mbarriers, bulk-tensor (TMA) copies, the 3xTF32 operand split.
#pragma once
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no -lcuda)

#include "common.cuh"

namespace dmvs {
namespace tc {

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

// instruction descriptor of kind::tf32: D = f32, A = B = tf32, both K-major, N >> 3, M = 128
__device__ __forceinline__ uint32_t idesc_tf32_m128(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

// instruction descriptor of kind::f16 with fp16 operands (format 0), D = f32, both K-major, K = 16 per instruction
__device__ __forceinline__ uint32_t idesc_f16_m128(int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

// the same with the descriptors given as (lower word, upper word) pairs: the upper words (SBO, version, layout) are
// loop invariant, the lower words (start address, LBO) advance by 32-bit adds
__device__ __forceinline__ void umma_tf32_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// kind::f16 twin of umma_tf32_w (K = 16 halves = the same 32 bytes per operand row)
__device__ __forceinline__ void umma_f16_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %6, 0;\nmov.b64 da, {%1, %2};\nmov.b64 db, {%3, %4};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// two floats -> packed fp16 pair (first argument in the LOW half), round to nearest, saturating at +-65504
__device__ __forceinline__ uint32_t pack_f16x2(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}

// arrives on `bar` once every MMA issued so far by this thread has retired (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// One lane of a converged warp (the compiler keeps the guarded code on the uniform datapath)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void fence_tc_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_tc_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// generic-proxy writes of this thread become visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// 128-bit shared-memory accesses by 32-bit shared address (a generic pointer costs an S2UR + address conversion per use)
__device__ __forceinline__ void sts128(uint32_t saddr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
  return v;
}

// ---- mbarriers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    if (!done && ++spins > (1u << 24)) __trap();   // watchdog: a lost arrival must not hang the GPU
  }
}

// ---- TMA -----------------------------------------------------------------------------------------------------------
// 5-D tiled tensor load (coordinates innermost first); completes `bytes of the box` on `bar`; out-of-bounds elements
// (negative coordinates included) are written as zeros - the convolution's zero padding
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// contiguous bulk copy global -> shared (16-byte granularity)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(map) : "memory");
}

// ---- 3xTF32 split ---------------------------------------------------------------------------------------------------
// x = hi + lo + r with hi, lo exactly representable in TF32 and |r| < 2^-21 |x|.  hi is x rounded to nearest (integer
// add of half an ulp, then mask - the result of cvt.rna.tf32.f32 for finite values, which ptxas expands to four
// instructions on sm_100a); lo = x - hi is exact in fp32 and is truncated to TF32 explicitly, so the result does not
// depend on what the tensor core does with the 13 low mantissa bits.  4 instructions per value.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = __uint_as_float(__float_as_uint(x - hi) & 0xffffe000u);
}

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// ---- TMEM -> registers: NCH consecutive fp32 columns of this thread's lane (32x32b shape); no wait inside -----------
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

}  // namespace tc

// ---- host: tensor-map encoder through the runtime's driver entry point ---------------------------------------------
typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TensorMapEncodeTiledFn tensor_map_encoder() {
  static TensorMapEncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return reinterpret_cast<TensorMapEncodeTiledFn>(p);
  }();
  return fn;
}

// fp32 channels-last activation [N][D][H][W][C] (pixel stride ps floats) as a 5-D tensor (C, W, H, D, N) with a box of
// `box_c` channels x box_w x box_h pixels traversed with step `step` in W and H (a stride-2 layer loads one phase plane)
inline bool make_activation_map(CUtensorMap* map, const float* base, int C, int ps, int W, int H, int D, int N, int box_c, int box_w,
                                int box_h, int step) {
  TensorMapEncodeTiledFn enc = tensor_map_encoder();
  if (!enc || box_w * step > 256 || box_h * step > 256) return false;
  const cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  const cuuint64_t strides[4] = {(cuuint64_t)ps * 4, (cuuint64_t)W * ps * 4, (cuuint64_t)H * W * ps * 4,
                                 (cuuint64_t)D * H * W * ps * 4};
  const cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)(box_w * step), (cuuint32_t)(box_h * step), 1, 1};
  const cuuint32_t estr[5] = {1, (cuuint32_t)step, (cuuint32_t)step, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace dmvs
