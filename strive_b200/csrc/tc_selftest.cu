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
