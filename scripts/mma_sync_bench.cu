// Throughput of the legacy warp-level tensor path (mma.sync) on sm_100a, per SM: m16n8k8 tf32 and m16n8k16 bf16 with fp32
// accumulation, 8 independent accumulator chains per warp.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_sync_bench scripts/mma_sync_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int KIND>
__global__ void bench(float* out, int iters, long long* cyc) {
  float acc[8][4];
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, threadIdx.x * 5u, threadIdx.x * 7u};
  uint32_t b[2] = {threadIdx.x * 11u, threadIdx.x * 13u};
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4 * sizeof(float));
  cudaMallocManaged(&cyc, 8);
  const int iters = 4096;
  for (int kind = 0; kind < 2; kind++)
    for (int warps = 4; warps <= 32; warps *= 2) {
      if (kind == 0) bench<0><<<148, warps * 32>>>(out, iters, cyc); else bench<1><<<148, warps * 32>>>(out, iters, cyc);
      cudaDeviceSynchronize();
      const double mmas = (double)iters * 8 * warps;
      const double macs = mmas * (kind == 0 ? 16 * 8 * 8 : 16 * 8 * 16);
      printf("%s warps/SM %2d: %.2f cycles per MMA per SM, %.0f MAC/clk/SM (FFMA peak = 128)\n", kind == 0 ? "tf32 m16n8k8 " : "bf16 m16n8k16", warps,
             (double)*cyc / mmas, macs / (double)*cyc);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
