// Self-test of the tcgen05 primitives in tc.cuh (exported so the GPU test-suite pins them on real hardware):
//   test 0: plain 128 x N x K bf16 GEMM from canonical no-swizzle K-major tiles
//   test 1: "shifted window" operand addressing used by the implicit-GEMM convolutions: rows 16 B apart inside a
//           larger buffer, start address 16-B (not 128-B) aligned, arbitrary LBO / SBO.
#include "common.cuh"
#include "tc.cuh"

#define ST_K 32
#define ST_N 32

__global__ void __launch_bounds__(128) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                          const float* __restrict__ X0, const float* __restrict__ X1,
                                                          float* __restrict__ D0, float* __restrict__ D1) {
  __shared__ __align__(128) __nv_bfloat16 sA[(ST_K / 8) * 16 * 64];   // [kgroup][mgroup][8 rows][8]
  __shared__ __align__(128) __nv_bfloat16 sB[(ST_K / 8) * 4 * 64];    // [kgroup][ngroup][8 rows][8]
  __shared__ __align__(128) __nv_bfloat16 sX0[144 * 8];               // [row][8]
  __shared__ __align__(128) __nv_bfloat16 sX1[144 * 8];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * ST_K; i += 128) {
    const int m = i / ST_K, k = i % ST_K;
    sA[(((k >> 3) * 16 + (m >> 3)) * 8 + (m & 7)) * 8 + (k & 7)] = __float2bfloat16_rn(A[i]);
  }
  for (int i = tid; i < ST_N * ST_K; i += 128) {
    const int n = i / ST_K, k = i % ST_K;
    sB[(((k >> 3) * 4 + (n >> 3)) * 8 + (n & 7)) * 8 + (k & 7)] = __float2bfloat16_rn(B[i]);
  }
  for (int i = tid; i < 144 * 8; i += 128) {
    sX0[i] = __float2bfloat16_rn(X0[i]);
    sX1[i] = __float2bfloat16_rn(X1[i]);
  }
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base, 64);
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;
  const uint32_t idesc = tc::idesc_bf16_f32(128, ST_N);
  if (tid == 0) {
    // test 0 -> TMEM columns [0,32)
    for (int j = 0; j < ST_K / 16; j++) {
      const uint64_t ad = tc::smem_desc(tc::smem_u32(sA) + j * 2 * 2048, 2048, 128);
      const uint64_t bd = tc::smem_desc(tc::smem_u32(sB) + j * 2 * 512, 512, 128);
      tc::mma_bf16(tm, ad, bd, idesc, j > 0);
    }
    // test 1 -> TMEM columns [32,64): A row m = [ X0[m+1][0:8] | X1[m+3][0:8] ], K = 16, B = first 16 k of sB rows
    {
      const uint32_t a0 = tc::smem_u32(sX0) + 1 * 16;
      const uint32_t a1 = tc::smem_u32(sX1) + 3 * 16;
      const uint64_t ad = tc::smem_desc(a0, a1 - a0, 128);
      const uint64_t bd = tc::smem_desc(tc::smem_u32(sB), 512, 128);
      tc::mma_bf16(tm + 32, ad, bd, idesc, 0);
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  float v[16];
  for (int c = 0; c < 2; c++) {
    tc::tmem_ld16(tm + lane_base + c * 16, v);
    for (int i = 0; i < 16; i++) D0[tid * ST_N + c * 16 + i] = v[i];
    tc::tmem_ld16(tm + lane_base + 32 + c * 16, v);
    for (int i = 0; i < 16; i++) D1[tid * ST_N + c * 16 + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, 64);
}

// A (128,32), B (32,32), X0/X1 (144,8) fp32 inputs (rounded to bf16 inside); D0, D1 (128,32) fp32 outputs.
extern "C" int strive_tc_selftest(const float* A, const float* B, const float* X0, const float* X1, float* D0, float* D1, void* stream) {
  STRIVE_CHECK(A && B && X0 && X1 && D0 && D1, STRIVE_EINVAL, "strive_tc_selftest: null argument");
  KPROF("tc_selftest", (cudaStream_t)stream, tc_selftest_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(A, B, X0, X1, D0, D1));
  STRIVE_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// test 2: CTA pair (cta_group::2).  Cluster of two CTAs on one TPC; CTA r holds rows [128 r, 128 r + 128) of A and rows
// [N/2 r, N/2 r + N/2) of B at the SAME shared-memory offsets; the leader (rank 0) issues ONE M = 256 MMA per K step that
// reads both CTAs' operands and writes D (128 x N per CTA) into both CTAs' tensor memory; tcgen05.commit multicasts the
// completion to the barrier of both CTAs.  Every wait is bounded: a protocol error ends as a flag, not as a hang.
// ------------------------------------------------------------------------------------------------------
#define SP_K 32

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, int tries) {
  const uint32_t a = tc::smem_u32(bar);
  for (int i = 0; i < tries; i++) {
    uint32_t ok = 0;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity), "r"(1000u) : "memory");
    if (ok) return true;
  }
  return false;
}

template <int SP_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) tc_selftest_pair_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                                                         float* __restrict__ D, int* __restrict__ flag) {
  __shared__ __align__(1024) __nv_bfloat16 sA[(SP_K / 8) * 16 * 64];             // [kgroup][mgroup 16][8 rows][8]
  __shared__ __align__(1024) __nv_bfloat16 sB[(SP_K / 8) * (SP_N / 16) * 64];    // [kgroup][ngroup N/16][8 rows][8]: this CTA's N/2 rows
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  for (int i = tid; i < 128 * SP_K; i += 128) {
    const int m = i / SP_K, k = i % SP_K;
    sA[(((k >> 3) * 16 + (m >> 3)) * 8 + (m & 7)) * 8 + (k & 7)] = __float2bfloat16_rn(A[(size_t)(rank * 128 + m) * SP_K + k]);
  }
  for (int i = tid; i < (SP_N / 2) * SP_K; i += 128) {
    const int n = i / SP_K, k = i % SP_K;
    sB[(((k >> 3) * (SP_N / 16) + (n >> 3)) * 8 + (n & 7)) * 8 + (k & 7)] = __float2bfloat16_rn(B[(size_t)(rank * (SP_N / 2) + n) * SP_K + k]);
  }
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tmem_base)), "r"((uint32_t)SP_N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs: operands staged, barriers initialised, tensor memory allocated
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = tc::idesc_bf16_f32(256, SP_N);
    for (int j = 0; j < SP_K / 16; j++) {
      const uint64_t ad = tc::smem_desc(tc::smem_u32(sA) + j * 2 * 2048, 2048, 128);
      const uint64_t bd = tc::smem_desc(tc::smem_u32(sB) + j * 2 * (SP_N / 16) * 128, (SP_N / 16) * 128, 128);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(ad),
                   "l"(bd), "r"(idesc), "r"((uint32_t)(j > 0))
                   : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(tc::smem_u32(&bar)),
                 "h"((uint16_t)3)
                 : "memory");
  }
  const bool ok = mbar_wait_bounded(&bar, 0, 200000);
  if (!ok && tid == 0) atomicOr(flag, 1 << rank);
  tc::tc_fence_after();
  if (ok) {
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float v[16];
    for (int c = 0; c < SP_N / 16; c++) {
      tc::tmem_ld16(tm + lane_base + c * 16, v);
      for (int i = 0; i < 16; i++) D[(size_t)(rank * 128 + tid) * SP_N + c * 16 + i] = v[i];
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // nobody frees tensor memory the peer's MMA may still write
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"((uint32_t)SP_N) : "memory");
}

// A (256,32), B (N,32) fp32 inputs (rounded to bf16 inside); D (256,N) fp32 output, N = 64 or 128; flag (1 int32, device): bit r set =
// CTA r timed out.
extern "C" int strive_tc_selftest_pair(const float* A, const float* B, float* D, int32_t* flag, int32_t n, void* stream) {
  STRIVE_CHECK(A && B && D && flag, STRIVE_EINVAL, "strive_tc_selftest_pair: null argument");
  STRIVE_CHECK(n == 64 || n == 128, STRIVE_EINVAL, "strive_tc_selftest_pair: n must be 64 or 128");
  cudaStream_t st = (cudaStream_t)stream;
  STRIVE_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t), st));
  if (n == 64) tc_selftest_pair_kernel<64><<<2, 128, 0, st>>>(A, B, D, flag);
  else tc_selftest_pair_kernel<128><<<2, 128, 0, st>>>(A, B, D, flag);
  STRIVE_LAUNCH_CHECK();
  return 0;
}
