// CTA-wide weight stream for the decoder's fp32 MLP/GRU kernels (rollout.cu).
//
// The per-agent layers are GEMV-shaped per warp (a warp owns R rows, a lane owns OUT/32 columns) and the matrices are
// small (16-84 KB), so what bounds them is how the weights reach the FMAs: read per warp from L1/L2 they cost one L2
// round trip per k-group and every warp streams every matrix (measured: node kernels at 0.3 % of the fp32 peak, edge
// kernels L2-bandwidth bound).  Here ONE producer lane streams the matrices a kernel multiplies by -- in program order,
// chunk by chunk, running ahead across layer boundaries -- into a ring of shared-memory stages with cp.async.bulk
// (completion on an mbarrier), and all consumer warps of the CTA read each chunk from shared memory
// (conflict-free 128-bit reads: 32 lanes x V contiguous floats), wait full -> use -> arrive empty.
// The k order of every dot product is unchanged (sequential fmaf chain), so results are bit-identical to the
// register/L1 version these replace.
#pragma once
#include "tc.cuh"

#ifndef WP_STAGES
#define WP_STAGES 4
#endif
#define WP_STAGE_FLOATS 4096                       // 16 KB per stage
#define WP_SMEM_FLOATS (WP_STAGES * WP_STAGE_FLOATS + 16)   // stages + 2*WP_STAGES mbarriers (8 B each), 16-byte multiple

#ifndef WP_WAIT_HINT_NS
#define WP_WAIT_HINT_NS 2000
#endif

struct WPipe {
  float* stage;
  uint64_t* full;
  uint64_t* empty;
  uint32_t cnt;    // chunks produced (producer lane) / consumed (consumer warp) so far
};

// rows of a [red][OUT] matrix per stage
__host__ __device__ constexpr int wp_chunk_rows(int out) { return (WP_STAGE_FLOATS / out) & ~3; }

__device__ __forceinline__ void wp_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wp_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc),
               "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}

// L1 prefetch of a small parameter vector (bias / LayerNorm affine): the kernels read these between GEMMs, where an L2 round
// trip would be fully exposed
__device__ __forceinline__ void wp_prefetch_l1(const float* __restrict__ p, int nfloats, int lane) {
  for (int i = lane * 32; i < nfloats; i += 32 * 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + i));
}

// bounded wait: a producer/consumer schedule mismatch must end in a trap (launch error -> RuntimeError), never in a hang
__device__ __forceinline__ void wp_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = tc::smem_u32(bar);
  uint32_t ok = 0;
#pragma unroll 1
  for (int spin = 0; spin < (1 << 22); spin++) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity), "r"((uint32_t)WP_WAIT_HINT_NS)
        : "memory");
    if (ok) return;
  }
  __trap();
}

// Called by every thread of the CTA before the roles split.  `base` is 16-byte aligned shared memory of WP_SMEM_FLOATS floats.
__device__ __forceinline__ WPipe wp_init(float* base, int consumer_warps) {
  WPipe p;
  p.stage = base;
  p.full = reinterpret_cast<uint64_t*>(base + WP_STAGES * WP_STAGE_FLOATS);
  p.empty = p.full + WP_STAGES;
  p.cnt = 0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < WP_STAGES; s++) {
      tc::mbar_init(&p.full[s], 1);
      tc::mbar_init(&p.empty[s], consumer_warps);
    }
    tc::fence_mbar_init();
  }
  __syncthreads();
  return p;
}

// producer lane: stream matrix Wm [red][out] (row-major, 16-byte aligned, red % 4 == 0)
__device__ __forceinline__ void wp_produce(WPipe& p, const float* __restrict__ Wm, int red, int out) {
  const int ch = wp_chunk_rows(out);
  for (int k0 = 0; k0 < red; k0 += ch) {
    const int rows = min(ch, red - k0);
    const uint32_t s = p.cnt % WP_STAGES;
    wp_wait(&p.empty[s], ((p.cnt / WP_STAGES) & 1) ^ 1);
    const uint32_t bytes = (uint32_t)(rows * out) * 4u;
    wp_mbar_expect_tx(&p.full[s], bytes);
    wp_bulk_g2s(p.stage + s * WP_STAGE_FLOATS, Wm + (size_t)k0 * out, bytes, &p.full[s]);
    p.cnt++;
  }
}

// consumer warp skips a matrix without using it (keeps the ring in step when a warp has no rows for it)
__device__ __forceinline__ void wp_skip(WPipe& p, int red, int out, int lane) {
  const int ch = wp_chunk_rows(out);
  for (int k0 = 0; k0 < red; k0 += ch) {
    const uint32_t s = p.cnt % WP_STAGES;
    wp_wait(&p.full[s], (p.cnt / WP_STAGES) & 1);
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&p.empty[s]);
    p.cnt++;
  }
}

// acc[r][v] += sum_k xs[r*ldx + k] * Wm[k*OUT + lane*V + v], Wm streamed through the pipe (same k order as warp_gemm).
#ifdef WP_NOINLINE
#define WP_GEMM_INLINE __noinline__
#else
#define WP_GEMM_INLINE __forceinline__
#endif
template <int OUT, int R>
__device__ WP_GEMM_INLINE void pipe_gemm(WPipe& p, int red, const float* xs, int ldx, float (&acc)[R][OUT / 32], int lane) {
  constexpr int V = OUT / 32;
  constexpr int CH = wp_chunk_rows(OUT);
  for (int k0 = 0; k0 < red; k0 += CH) {
    const int rows = min(CH, red - k0);
    const uint32_t s = p.cnt % WP_STAGES;
    wp_wait(&p.full[s], (p.cnt / WP_STAGES) & 1);
    const float* wp = p.stage + s * WP_STAGE_FLOATS + lane * V;
    const float* xk = xs + k0;
    // ~128 FMAs per unrolled body: enough independent work for one warp per scheduler without blowing the instruction
    // cache (these kernels run every instruction once per warp; unroll 4 at R=4 was 27 % stall_no_inst)
    constexpr int UNR = (R * V >= 16) ? 2 : ((R * V >= 8) ? 4 : 8);
#pragma unroll UNR
    for (int k = 0; k < rows; k += 4) {
      float w[4][V];
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        if constexpr (V == 4) {
          const float4 t = *reinterpret_cast<const float4*>(wp + (k + kk) * OUT);
          w[kk][0] = t.x; w[kk][1] = t.y; w[kk][2] = t.z; w[kk][3] = t.w;
        } else if constexpr (V == 2) {
          const float2 t = *reinterpret_cast<const float2*>(wp + (k + kk) * OUT);
          w[kk][0] = t.x; w[kk][1] = t.y;
        } else {
          w[kk][0] = wp[(k + kk) * OUT];
        }
      }
#pragma unroll
      for (int r = 0; r < R; r++) {
        const float4 xv = *reinterpret_cast<const float4*>(xk + r * ldx + k);
#pragma unroll
        for (int v = 0; v < V; v++) {
          acc[r][v] = fmaf(xv.x, w[0][v], acc[r][v]);
          acc[r][v] = fmaf(xv.y, w[1][v], acc[r][v]);
          acc[r][v] = fmaf(xv.z, w[2][v], acc[r][v]);
          acc[r][v] = fmaf(xv.w, w[3][v], acc[r][v]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&p.empty[s]);
    p.cnt++;
  }
}

// GRU gate GEMM through the pipe: Wm [red][192], lane owns columns g*64 + lane*2 + {0,1} for g = 0,1,2 (r,z,n).
template <int R>
__device__ __forceinline__ void pipe_gemm_gru(WPipe& p, int red, const float* xs, int ldx, float (&acc)[R][6], int lane) {
  constexpr int CH = wp_chunk_rows(192);
  for (int k0 = 0; k0 < red; k0 += CH) {
    const int rows = min(CH, red - k0);
    const uint32_t s = p.cnt % WP_STAGES;
    wp_wait(&p.full[s], (p.cnt / WP_STAGES) & 1);
    const float* wp = p.stage + s * WP_STAGE_FLOATS + lane * 2;
    const float* xk = xs + k0;
#pragma unroll 2
    for (int k = 0; k < rows; k += 4) {
      float w[4][6];
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
#pragma unroll
        for (int g = 0; g < 3; g++) {
          const float2 t = *reinterpret_cast<const float2*>(wp + (k + kk) * 192 + g * 64);
          w[kk][g * 2] = t.x;
          w[kk][g * 2 + 1] = t.y;
        }
      }
#pragma unroll
      for (int r = 0; r < R; r++) {
        const float4 xv = *reinterpret_cast<const float4*>(xk + r * ldx + k);
#pragma unroll
        for (int v = 0; v < 6; v++) {
          acc[r][v] = fmaf(xv.x, w[0][v], acc[r][v]);
          acc[r][v] = fmaf(xv.y, w[1][v], acc[r][v]);
          acc[r][v] = fmaf(xv.z, w[2][v], acc[r][v]);
          acc[r][v] = fmaf(xv.w, w[3][v], acc[r][v]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&p.empty[s]);
    p.cnt++;
  }
}
