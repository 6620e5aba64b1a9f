// tcgen05 / TMEM / mbarrier primitives for sm_100a (inline PTX; no CUTLASS dependency).
// Bit layouts follow the PTX ISA "tcgen05" matrix/instruction descriptors (cross-checked against
// cute/arch/mma_sm100_desc.hpp of the CUTLASS headers shipped in this image).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t ok = 0;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"     // suspends (no issue slots) up to the hint or completion
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity), "r"(4000u)
        : "memory");
  } while (!ok);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}

// one lane of a CONVERGED warp (the compiler treats the guarded region as single-threaded/uniform: no per-instruction election)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t"
      "}\n"
      : "=r"(pred)::"memory");
  return pred != 0;
}

// L2 prefetch of a contiguous global range (16-byte aligned, size multiple of 16): no register or shared-memory destination
__device__ __forceinline__ void prefetch_l2(const void* gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256), pointer 32-byte aligned
__device__ __forceinline__ void ldg256(const float* p, float (&v)[8]) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, float a, float b, float c, float d, float e, float f, float g, float h) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e), "f"(f), "f"(g), "f"(h) : "memory");
}

// ---- proxies / fences --------------------------------------------------------------------------------------
// generic-proxy shared-memory writes (st.shared) -> visible to the async proxy (tensor-core descriptor reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM allocation (one full warp executes; ncols power of two >= 32) ---------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- descriptors --------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, no swizzle ("interleave"): the operand is made of 8-row x 16-byte core
// matrices (8 rows 16 B apart = 128 contiguous bytes);  SBO = byte distance between consecutive 8-row groups along M/N,
// LBO = byte distance between the two core matrices along K of one K=16 (bf16) MMA.  All addresses/offsets 16 B aligned.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version 1 (Blackwell)
  return d;                 // base_offset = 0, lbo_mode = 0, layout_type = 0 (SWIZZLE_NONE)
}

// Descriptor words for the MMA issue loops: the 64-bit descriptor is {lo, hi} with lo = addr>>4 | (LBO>>4)<<16 and
// hi = SBO>>4 | version<<14; stepping through taps / K slices only adds a constant to `lo` (addresses < 256 KB never carry
// into the LBO field), so the single issuing thread spends ~1 integer add per operand per MMA.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }
__device__ __forceinline__ uint64_t desc_make(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// Instruction descriptor for kind::f16: BF16 x BF16 -> FP32, A and B K-major, dense.
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Instruction descriptor for kind::f16 with FP16 operands: FP16 x FP16 -> FP32 (a_format = b_format = 0), A and B K-major, dense.
__host__ __device__ constexpr uint32_t idesc_f16_f32(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// all previously issued MMAs of this thread -> arrive on the mbarrier when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: warp w (w = warp_id % 4) reads lanes 32w..32w+31; thread t gets lane 32w+t, 16 or 32 columns ----
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// two values -> packed bf16x2 hi and lo words (x = hi + lo): 1 cvt + 2 unpack + 2 sub + 1 cvt
__device__ __forceinline__ void split_pack2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - h0, x1 - h1);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// fp16 variant: x = hi + lo with hi = fp16(x), lo = fp16(x - hi): 11 + 11 mantissa bits (bf16 pairs carry 8 + 8).  Values must stay
// inside the fp16 range (|x| < 65504); a lo part below 2^-14 is subnormal and keeps an ABSOLUTE precision of 2^-25.
__device__ __forceinline__ void split_pack2_f16(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// split x = hi + lo with hi = bf16(x), lo = bf16(x - hi)
__device__ __forceinline__ void split_bf16(float x, float& hi, float& lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  hi = __bfloat162float(h);
  lo = x - hi;
}

}  // namespace tc
