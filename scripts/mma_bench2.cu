// Micro-benchmark 2: cycles per tcgen05.mma for the operand-reuse forms (collector::a keep/reuse, .ws with collector::b),
// kind::i8, and TS mode (A in TMEM), M=128, no-swizzle K-major operands.  One thread issues back to back.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I strive_b200/csrc scripts/mma_bench2.cu -o /tmp/mma_bench2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc.cuh"

#define MMA_FORM(NAME, TEXT)                                                                                             \
  __device__ __forceinline__ void NAME(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {                              \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t" TEXT " [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a),   \
                 "l"(b), "r"(idesc), "r"(1u)                                                                              \
                 : "memory");                                                                                             \
  }
MMA_FORM(mma_plain, "tcgen05.mma.cta_group::1.kind::f16")
MMA_FORM(mma_a_fill, "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill")
MMA_FORM(mma_a_use, "tcgen05.mma.cta_group::1.kind::f16.collector::a::use")
MMA_FORM(mma_a_last, "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse")
MMA_FORM(mma_ws_plain, "tcgen05.mma.ws.cta_group::1.kind::f16")
MMA_FORM(mma_ws_fill, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill")
MMA_FORM(mma_ws_use, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::use")
MMA_FORM(mma_i8, "tcgen05.mma.cta_group::1.kind::i8")
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
               "r"(a_tmem), "l"(b), "r"(idesc), "r"(1u)
               : "memory");
}

// MODE 0 plain SS | 1 collector::a (fill then GROUP-1 reuse, B changes) | 2 .ws plain | 3 .ws collector::b0 (fill then GROUP-1 use, A changes)
// 4 kind::i8 | 5 TS (A in TMEM)
template <int MODE, int N, int GROUP>
__global__ void __launch_bounds__(128) bench(int iters, long long* out) {
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
      uint32_t idesc = tc::idesc_bf16_f32(128, N);
      if (MODE == 4) idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // s32 accum, s8 x s8
      // shifted-window style A (conv2 layout: rows 16 B apart in a parity plane) and canonical B
      const uint32_t a0 = tc::desc_lo(tc::smem_u32(smem), 11200), ah = tc::desc_hi(320);
      const uint32_t b0 = tc::desc_lo(tc::smem_u32(smem) + 128 * 1024, 128 * (N / 8)), bh = tc::desc_hi(128);
      long long t0 = clock64();
#pragma unroll 1
      for (int it = 0; it < iters; it += 16) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
          const int g = u / GROUP, j = u % GROUP;
          uint32_t al = a0, bl = b0;
          if (MODE == 1) { al += (g * 160) >> 4; bl += ((u % 8) * 2048) >> 4; }          // A fixed within a group, B changes
          else if (MODE == 3) { al += (u * 160) >> 4; bl += ((g % 8) * 2048) >> 4; }     // B fixed within a group, A changes
          else { al += (u * 160) >> 4; bl += ((u % 8) * 2048) >> 4; }
          const uint64_t ad = tc::desc_make(al, ah), bd = tc::desc_make(bl, bh);
          const uint32_t d = tm + (u % 4) * 64;
          if (MODE == 0) mma_plain(d, ad, bd, idesc);
          if (MODE == 1) { if (j == 0) mma_a_fill(d, ad, bd, idesc); else if (j == GROUP - 1) mma_a_last(d, ad, bd, idesc); else mma_a_use(d, ad, bd, idesc); }
          if (MODE == 2) mma_ws_plain(d, ad, bd, idesc);
          if (MODE == 3) { if (j == 0) mma_ws_fill(d, ad, bd, idesc); else mma_ws_use(d, ad, bd, idesc); }
          if (MODE == 4) mma_i8(d, ad, bd, idesc);
          if (MODE == 5) mma_ts(d, tm + 256 + (u % 8) * 8, bd, idesc);
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

template <int MODE, int N, int GROUP>
void run(long long* d, const char* what) {
  const int iters = 4096;
  auto k = bench<MODE, N, GROUP>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<148, 128, 200 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-34s N=%3d group=%2d : issue %.1f cyc/MMA, complete %.1f cyc/MMA (math floor %d, SS model %.0f) %s\n", what, N, GROUP,
         (double)h[0] / iters, (double)h[1] / iters, N / 2, (4096.0 + 32 * N) / 128, e == cudaSuccess ? "" : cudaGetErrorString(e));
  if (e != cudaSuccess) exit(1);
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 2 * sizeof(long long));
  run<0, 32, 1>(d, "SS plain");
  run<0, 64, 1>(d, "SS plain");
  run<0, 128, 1>(d, "SS plain");
  run<0, 256, 1>(d, "SS plain");
  run<5, 32, 1>(d, "TS (A in TMEM)");
  run<5, 64, 1>(d, "TS (A in TMEM)");
  run<5, 128, 1>(d, "TS (A in TMEM)");
  run<1, 32, 2>(d, "SS collector::a fill/lastuse");
  run<1, 64, 2>(d, "SS collector::a fill/lastuse");
  run<1, 64, 4>(d, "SS collector::a fill/use/lastuse");
  run<4, 32, 1>(d, "SS kind::i8 (K=32)");
  run<4, 64, 1>(d, "SS kind::i8 (K=32)");
  run<4, 128, 1>(d, "SS kind::i8 (K=32)");
  run<2, 64, 1>(d, ".ws plain");
  run<2, 128, 1>(d, ".ws plain");
  run<3, 64, 4>(d, ".ws collector::b0 fill/use");
  run<3, 64, 8>(d, ".ws collector::b0 fill/use");
  run<3, 128, 8>(d, ".ws collector::b0 fill/use");
  run<2, 32, 1>(d, ".ws plain");
  run<3, 32, 8>(d, ".ws collector::b0 fill/use");
  return 0;
}
