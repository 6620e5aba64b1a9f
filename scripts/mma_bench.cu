// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, SS mode, no-swizzle K-major operands) as a function
// of N and of the operand layout, issued back to back by one thread.  Used to size the map-encoder MMA tiles
// (profiles/r01_mma_microbench.txt).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I strive_b200/csrc
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc.cuh"

struct Case {
  int N, sbo_a, lbo_a, a_step, nacc, b_step;
};

template <int N, int SBO_A, int LBO_A, int A_STEP, int NACC, int B_STEP, int N2 = 0, int PATTERN = 0>
__global__ void __launch_bounds__(128) bench(int iters, long long* out) {
  constexpr Case c = {N, SBO_A, LBO_A, A_STEP, NACC, B_STEP};
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x;
  for (int i = tid; i < 200 * 1024 / 16; i += 128) reinterpret_cast<int4*>(smem)[i] = make_int4(0, 0, 0, 0);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  if (tid < 32) tc::tmem_alloc(&tmem_base, 512);
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;
  if (tid < 32) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::idesc_bf16_f32(128, c.N);
      const uint32_t idesc2 = tc::idesc_bf16_f32(128, N2 ? N2 : N);
      const uint32_t a0 = tc::desc_lo(tc::smem_u32(smem), c.lbo_a), ah = tc::desc_hi(c.sbo_a);
      const uint32_t b0 = tc::desc_lo(tc::smem_u32(smem) + 128 * 1024, 128 * (c.N / 8)), bh = tc::desc_hi(128);
      long long t0 = clock64();
#pragma unroll 1
      for (int it = 0; it < iters; it += 16) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
          const uint32_t al = a0 + ((u * c.a_step) >> 4);
          const uint32_t bl = b0 + (((u % 8) * c.b_step) >> 4);
          // PATTERN 0: alternate N / N2 every MMA;  PATTERN 1: 8 x N then 8 x N2
          const bool second = N2 != 0 && (PATTERN == 0 ? (u & 1) : (u >= 8));
          tc::mma_bf16(tm + (u % c.nacc) * c.N, tc::desc_make(al, ah), tc::desc_make(bl, bh), second ? idesc2 : idesc, 1u);
        }
      }
      long long t1 = clock64();
      tc::mma_commit(&bar);
      tc::mbar_wait(&bar, 0);
      long long t2 = clock64();
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
    }
    __syncwarp();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tm, 512);
}


template <int N, int SBO_A, int LBO_A, int A_STEP, int NACC, int B_STEP, int N2 = 0, int PATTERN = 0>
void run(long long* d) {
  const int iters = 4096;
  auto k = bench<N, SBO_A, LBO_A, A_STEP, NACC, B_STEP, N2, PATTERN>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int grid : {1, 148}) {
    k<<<grid, 128, 200 * 1024>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    if (N2) printf("[N2=%d pattern=%d: expected %.1f] ", N2, PATTERN, ((4096.0 + 32 * N) / 128 + (4096.0 + 32 * N2) / 128) / 2);
    printf("N=%3d sbo=%5d lbo=%5d a_step=%4d b_step=%4d nacc=%d grid=%3d : issue %.1f cyc/MMA, complete %.1f cyc/MMA (math floor %d) %s\n", N, SBO_A, LBO_A, A_STEP,
           B_STEP, NACC, grid, (double)h[0] / iters, (double)h[1] / iters, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 2 * sizeof(long long));
  // canonical layout: core matrices of 128 B, SBO 256 (two K halves interleaved), LBO 128
  run<16, 256, 128, 4096, 1, 0>(d);
  run<32, 256, 128, 4096, 1, 0>(d);
  run<64, 256, 128, 4096, 1, 0>(d);
  run<128, 256, 128, 4096, 1, 0>(d);
  run<256, 256, 128, 4096, 1, 0>(d);
  run<16, 256, 128, 0, 1, 0>(d);
  run<64, 256, 128, 0, 1, 0>(d);
  run<16, 256, 128, 4096, 4, 0>(d);
  run<32, 256, 128, 4096, 4, 0>(d);
  run<64, 256, 128, 4096, 4, 0>(d);
  // shifted-window layouts of the encoder: conv1 (rows 16 B apart, LBO 16, SBO = 2 patch rows), conv2 (SBO = plane row)
  run<16, 1120, 16, 560, 1, 512>(d);
  run<32, 1120, 16, 560, 1, 512>(d);
  run<32, 320, 11200, 160, 1, 1024>(d);
  run<64, 320, 11200, 160, 1, 2048>(d);
  run<96, 320, 11200, 160, 1, 2048>(d);
  run<128, 320, 11200, 160, 1, 2048>(d);
  // mixed-N sequences (conv2..4 issue N=64 and N=32 MMAs into the same accumulator)
  run<64, 320, 11200, 160, 1, 2048, 32, 0>(d);
  run<64, 320, 11200, 160, 1, 2048, 32, 1>(d);
  run<128, 320, 11200, 160, 1, 2048, 64, 0>(d);
  run<128, 320, 11200, 160, 1, 2048, 64, 1>(d);
  return 0;
}
