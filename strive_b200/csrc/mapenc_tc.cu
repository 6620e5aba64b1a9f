// Tensor-core (tcgen05 + TMEM) implicit-GEMM convolutions of the map encoder, warp-specialised:
//   warps 0..10  producers  : stage operand tiles into a shared-memory ring (GroupNorm + ReLU + bf16 hi/lo split fused)
//   warp  11     MMA issuer : one elected thread issues tcgen05.mma, frees ring slots with tcgen05.commit
//   warps 12..15 epilogue   : TMEM -> registers -> +bias -> NHWC fp32 store + fp64 GroupNorm statistics
// Two TMEM accumulators, so the epilogue of tile i overlaps the MMAs of tile i+1; all hand-offs are mbarriers.
//
// Numerics: bf16 operand SPLITTING with fp32 accumulation in TMEM.  x = hi + lo (hi = bf16(x), lo = bf16(x - hi)).
//   conv1: the input is the binary crop (exact in bf16); weights are split, [W_hi | W_lo] stacked along N -> 1 MMA per K step.
//   conv2..4: activations and weights both split -> A_hi*[W_hi | W_lo] + A_lo*W_hi (2 MMAs, dropped lo*lo term 2^-18).
//   conv5..fc: hi*hi + lo*hi + hi*lo as 3 MMAs.
// Measured against the fp64 oracle: 1.4e-5 abs on O(1) features (tests allow 1e-4), at 1/3 of the dense bf16 tensor rate.
//
// Operand addressing ("shifted window", conv1..conv4): the input tile is written to shared memory ONCE, columns
// de-interleaved by parity (stride-2 conv -> consecutive output pixels are consecutive 16-byte rows of a parity plane).
// Every filter tap is then only a different start address in the K-major no-swizzle matrix descriptor: no im2col copy.
#define STRIVE_PDL_CLASS 2   // bit of strive_set_pdl() that enables programmatic dependent launch for this file's kernels
#include "common.cuh"
#include "tc.cuh"
#include <cstdio>
#include <cstdlib>

#define TC_NPROD 11
#define TC_MMA_WARP 11
#define TC_EPI_WARP0 12
#define TC_THREADS 512
#define TC_PROD_THREADS (TC_NPROD * 32)

// Pipeline diagnostics (strive_tc_trace): per kernel, cycles each role spent blocked on its mbarriers, summed over CTAs.
//   [0] producer: waiting for a free ring slot   [1] producer: loop total
//   [2] MMA: waiting for a filled slot           [3] MMA: waiting for a free accumulator   [4] MMA: loop total
//   [5] epilogue: waiting for an accumulator     [6] epilogue: loop total                  [7] CTAs
__device__ unsigned long long g_tc_trace[4][8];
#define TRACE_T() clock64()
__device__ __forceinline__ void trace_add(int k, int slot, long long v) { atomicAdd(&g_tc_trace[k][slot], (unsigned long long)v); }

extern "C" int strive_tc_trace(unsigned long long* out32, int reset) {
  if (out32 && cudaMemcpyFromSymbol(out32, g_tc_trace, sizeof(unsigned long long) * 32) != cudaSuccess) return -1;
  if (reset) {
    unsigned long long z[32] = {0};
    if (cudaMemcpyToSymbol(g_tc_trace, z, sizeof(z)) != cudaSuccess) return -1;
  }
  return 0;
}

// Timing experiments only (strive_tc_debug, scripts/mapenc_dbg_sweep.py): bit0 epilogue skips its global stores, bit1 producers skip
// their shared stores, bit2 producers skip their global loads, bit3 the MMA warp issues no MMAs, bit5 tc_gemm skips its weight
// copies.  Results are garbage when set.  The flags exist only in builds with -DSTRIVE_TC_DEBUG=1 (scripts/build_variants.sh):
// the operand producers sit at the register limit and every extra predicate costs them (measured: +3 % on conv1, spills in conv2).
#ifndef STRIVE_TC_DEBUG
#define STRIVE_TC_DEBUG 0
#endif
static int g_tc_dbg = 0;
extern "C" int strive_tc_debug(int flags) {
  g_tc_dbg = flags;
  return 0;
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

__device__ __forceinline__ void gn_stats(const double* __restrict__ st, int crop, double cnt, float& mean, float& rstd) {
  const double mu = st[(size_t)crop * 2] / cnt;
  double var = st[(size_t)crop * 2 + 1] / cnt - mu * mu;
  if (var < 0.0) var = 0.0;
  mean = (float)mu;
  rstd = (float)(1.0 / sqrt(var + 1e-5));
}

// Per-CTA table of the GroupNorm statistics (mean, rstd) of the crops a CTA works on, filled ONCE in the prologue by the first 64
// threads (one crop per thread): the float64 division + rsqrt behind gn_stats saturate the FP64 pipe when every producer warp runs
// them at every crop change (ncu: 16 % of the stall samples of conv3), and a producer-wide barrier around one warp doing it for
// everybody stalls ten warps.  A CTA's crops are a contiguous range; ranges longer than the table fall back to gn_stats.
#define GN_TAB 64
__device__ __forceinline__ void gn_table_fill(float* s_cm, float* s_cr, const double* __restrict__ st, int crop_first, int crop_last, double cnt, int tid) {
  if (tid < GN_TAB && crop_first + tid <= crop_last) gn_stats(st, crop_first + tid, cnt, s_cm[tid], s_cr[tid]);
}
__device__ __forceinline__ void gn_table_get(const float* s_cm, const float* s_cr, const double* __restrict__ st, int crop, int crop_first, double cnt,
                                             float& mean, float& rstd) {
  const int i = crop - crop_first;
  if (i < GN_TAB) { mean = s_cm[i]; rstd = s_cr[i]; }
  else gn_stats(st, crop, cnt, mean, rstd);
}

// ======================================================================================================
// conv1: 4 -> 16, k7 s2 on the INTEGER tensor path (tcgen05.mma kind::i8, s32 accumulators).
// The input is the binary crop -- exact as int8 {0,1}.  Each output channel's weights are written in 31-bit fixed point
//     w[o][k] ~= sc[o] * q[o][k],   q = round(w / max|w[o]| * 127 * 2^16) = sum_d digit_d 256^d,  digit_d in [-128, 127]
// and the three digit planes are stacked along N (rows n = 16 d + o), so ONE MMA (K = 32 = 8 taps x 4 channels) multiplies
// a pixel window with all of them; the s32 accumulators are exact and the epilogue recombines
//     out = sc[o] * float(acc2 * 65536 + acc1 * 256 + acc0) + bias        (exact int32, one conversion, one fma)
// |w - sc q| <= 2^-24 max|w[o]| (the fp32 half-ulp of the largest weight); summed over <= 196 active taps that is ~1e-8 on O(0.1)
// outputs, ten times below the rounding noise of the reference's own fp32 accumulation.  Three planes, not four: the
// epilogue is bound by the TMEM read bandwidth (64 B/clk: 4 planes = 131 KB per item took longer than the MMAs).  Against the bf16 hi/lo version
// this halves the operand bytes per tap (int8, K = 32 per instruction) -- the kernel is bound by the 128 B/clk shared-memory
// operand fetch (profiles/r01_ncu_full_v2_encoder.txt), so bytes are time.
// CTA item = 16 x 32 outputs = 4 sub-tiles (M = 128 rows = 16 output rows x 8 output columns of ONE column parity).
// Operand rows must be 16 bytes apart and a pixel is 4 bytes, so consecutive rows are 4 input pixels = 2 output columns
// apart: even output columns read the patch as stored, odd ones read a second copy shifted by 2 pixels.
// ======================================================================================================
#define T1_ROWS 37                    // input rows of an item: 2 * 16 + 5
#define T1_COLS 72                    // input columns 0..68 are needed (+ the zero-weight 8th tap), padded to a multiple of 4
#define T1_ROWB (T1_COLS * 4)         // 288 bytes per patch row
#define T1_COPY (T1_ROWS * T1_ROWB)   // 10656 bytes
#define T1_PATCH_BYTES (2 * T1_COPY)
#define T1_NP 48                      // N = 3 digit planes x 16 channels
#define T1_WBYTES (7 * 2 * T1_NP * 16)   // [ky][khalf][n = 16 d + o][16 k-bytes]; 16 fp32 scales follow in global memory
#define T1_RB 8                       // 8 row blocks x 4 column blocks of 16 x 32 outputs cover 125 x 125
#define T1_CB 4
#define T1_NBUF 4
// roles: warps 0-5 producers (the patch is 666 32-bit loads), warp 6 MMA issuer, warps 8-15 epilogue (two per TMEM lane quarter:
// recombining four digit planes is ~10 instructions per output, the epilogue -- not the tensor core -- was the bottleneck with 4 warps)
#define T1_NPROD 6
#define T1_PROD_THREADS (T1_NPROD * 32)
#define T1_NGRP 3
#define T1_MMA_WARP 6
#define T1_EPI_WARP0 8

// round-half-even(g / dx) exactly as torch.round(float64 quotient) (reference datasets/nuscenes_utils.py:254-255): multiply by
// the reciprocal; only when the product lands within 1e-6 of a .5 boundary (where the two could round differently) divide.
__device__ __forceinline__ int round_div_exact(float g, double dx, double inv) {
  const double gd = (double)g;
  const double q = gd * inv;
  int r = __double2int_rn(q);      // saturates for |q| >= 2^31: still "outside the map" 
  const double fr = fabs(q - (double)r);
  if (fr > 0.499999) r = __double2int_rn(gd / dx);
  return r;
}


// ------------------------------------------------------------------------------------------------------
// crop_pack: the rotated nearest-neighbour crop itself (exact get_map_obs arithmetic), written ONCE per crop as
// [256][256] bytes with bit c = layer c (64 KB per crop instead of the reference's 256 KB uint8 + 4 MB of int64 indices).
// Block = one 64 x 64 tile of crop pixels.  Its footprint in the raster is a rotated square of ~77 px: the bounding box
// (<= 116 x 116 px at 4 px/m) is first copied into shared memory with coalesced 32-bit loads, then every sample is a
// shared-memory byte read (a direct byte gather costs ~20 L1 sectors per warp instruction: the v2 kernel sat at 92 % of
// the L1 throughput).  Two block-uniform code paths with identical pixel-index arithmetic:
//   CLEAN   : pose finite, |q| < 5e4, the whole footprint inside the raster and staged -> no NaN / range / bounds tests
//   generic : everything else (NaN poses, crops over the map border, footprints that do not fit) reads global memory
// Samples within 0.49 of a rounding tie are resolved exactly (float64 quotient) after the fast pass.
// A thread owns 4 consecutive pixels of 4 rows -> 32-bit stores.
// ------------------------------------------------------------------------------------------------------
#define CP_TILE 64
#define CP_BOX_BYTES 18432
template <bool CLEAN>
__device__ __forceinline__ void crop_pack_rows(const uint8_t* __restrict__ s_box, const uint8_t* __restrict__ base, const float* __restrict__ lin_l,
                                               int H, int W, int P, int bx0, int by0, int bw, int bh, int pitch, float px, float py, float hc,
                                               float hs, float inv0f, float inv1f, float thr, const float (&whs)[4], const float (&whc)[4],
                                               int row0, int rg, int c4, unsigned short* s_q, int* s_nq, uint8_t* dst_tile) {
  unsigned mask = 0u;     // bit 4j + k: sample (row j, pixel k) of this thread needs the exact path
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int r = rg + 16 * j;
    const float l = __ldg(lin_l + row0 + r);
    const float lhc = __fmul_rn(l, hc), lhs = __fmul_rn(l, hs);     // gen_car_coords (:232-233), every product rounded on its own
    uint32_t word = 0u;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float gx = __fadd_rn(__fsub_rn(lhc, whs[k]), px);
      float gy = __fadd_rn(__fadd_rn(lhs, whc[k]), py);
      if (!CLEAN) {
        if (isnan(gx)) gx = 0.f;     // xys[torch.isnan(xys)] = 0.0 (:251)
        if (isnan(gy)) gy = 0.f;
      }
      // |fp32 quotient - float64 quotient| <= |q| 2^-23: accept the fp32 rounding unless it is that close to a .5 tie (thr)
      const float qx = gx * inv0f, qy = gy * inv1f;
      float rx, ry;
      int xp, yp;
      if (CLEAN) {
        // |q| < 5e4 here: q + 1.5 * 2^23 rounds to the nearest integer (ties to even, like rintf) and carries it in its low
        // mantissa bits -- FADD/IADD instead of the quarter-rate FRND + F2I
        const float mx = __fadd_rn(qx, 12582912.f), my = __fadd_rn(qy, 12582912.f);
        rx = __fsub_rn(mx, 12582912.f); ry = __fsub_rn(my, 12582912.f);
        xp = __float_as_int(mx) - 0x4B400000; yp = __float_as_int(my) - 0x4B400000;
      } else {
        rx = rintf(qx); ry = rintf(qy);
        xp = (int)rx; yp = (int)ry;
      }
      bool slow = !(fabsf(qx - rx) < thr && fabsf(qy - ry) < thr);
      if (!CLEAN) slow = slow || !(fabsf(qx) < 6e4f && fabsf(qy) < 6e4f);
      mask |= slow ? (1u << (4 * j + k)) : 0u;
      unsigned v;
      if (CLEAN) {
        v = s_box[(yp - by0) * pitch + (xp - bx0)];     // always inside the staged box (a tie sample is overwritten below)
      } else {
        if (slow || (unsigned)yp >= (unsigned)H || (unsigned)xp >= (unsigned)W) { xp = 0; yp = 0; }     // :260-262
        const unsigned ux = (unsigned)(xp - bx0), uy = (unsigned)(yp - by0);
        v = (ux < (unsigned)bw && uy < (unsigned)bh) ? (unsigned)s_box[uy * pitch + ux]
                                                     : (unsigned)__ldg(base + (size_t)((unsigned)yp * (unsigned)P + (unsigned)xp));
      }
      word |= (v & 15u) << (8 * k);
    }
    *reinterpret_cast<uint32_t*>(dst_tile + r * 256 + c4) = word;
  }
  if (mask) {
    int at = atomicAdd(s_nq, __popc(mask));
    while (mask) {
      const int bit = __ffs(mask) - 1;
      mask &= mask - 1;
      s_q[at++] = (unsigned short)((rg + 16 * (bit >> 2)) * CP_TILE + c4 + (bit & 3));
    }
  }
}

__global__ void __launch_bounds__(256, 5) crop_pack_kernel(StriveMap map, const float* __restrict__ pose, const int32_t* __restrict__ map_of,
                                                        uint8_t* __restrict__ packed_crop, int n) {
  STRIVE_PDL_TRIGGER();
  STRIVE_PDL_WAIT();
  __shared__ __align__(16) uint8_t s_box[CP_BOX_BYTES];
  __shared__ unsigned short s_q[CP_TILE * CP_TILE];
  __shared__ int s_nq;
  __shared__ int s_geo[6];     // bx0, by0, bw, bh, pitch, clean
  __shared__ float s_thr;
  const int crop = blockIdx.y, tid = threadIdx.x;
  const int row0 = (blockIdx.x >> 2) * CP_TILE, col0 = (blockIdx.x & 3) * CP_TILE;
  const int m = map_of[crop];
  const float px = pose[crop * 4 + 0], py = pose[crop * 4 + 1], hc = pose[crop * 4 + 2], hs = pose[crop * 4 + 3];
  const double dx0 = map.dx[m * 2 + 0], dx1 = map.dx[m * 2 + 1];
  const double inv0 = 1.0 / dx0, inv1 = 1.0 / dx1;
  const float inv0f = (float)inv0, inv1f = (float)inv1;
  const int H = map.H, W = map.W, P = map.packed_pitch;
  const uint8_t* base = map.packed + (size_t)m * H * P;
  if (tid == 0) {
    // bounding box of the tile in raster pixels from its 4 corners (+-2 px: the footprint is their convex hull up to rounding)
    s_nq = 0;
    int bx0 = 0, by0 = 0, bw = 0, bh = 0, pitch = 4, clean = 0;
    const float l0 = __ldg(map.lin_l + row0), l1 = __ldg(map.lin_l + row0 + CP_TILE - 1);
    const float w0 = __ldg(map.lin_w + col0), w1 = __ldg(map.lin_w + col0 + CP_TILE - 1);
    float xmin = 3e38f, xmax = -3e38f, ymin = 3e38f, ymax = -3e38f;
    bool finite = true;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float l = (k & 1) ? l1 : l0, w = (k & 2) ? w1 : w0;
      const float qx = (l * hc - w * hs + px) * inv0f, qy = (l * hs + w * hc + py) * inv1f;
      finite = finite && fabsf(qx) < 5e4f && fabsf(qy) < 5e4f;     // false for NaN too
      xmin = fminf(xmin, qx); xmax = fmaxf(xmax, qx); ymin = fminf(ymin, qy); ymax = fmaxf(ymax, qy);
    }
    // every |term| of gx, gy finite and small: no intermediate can overflow or become NaN in the sample arithmetic
    finite = finite && fabsf(px) < 1e6f && fabsf(py) < 1e6f && fabsf(hc) < 1e3f && fabsf(hs) < 1e3f;
    if (finite) {
      const int fx0 = (int)floorf(xmin) - 2, fx1 = (int)ceilf(xmax) + 2, fy0 = (int)floorf(ymin) - 2, fy1 = (int)ceilf(ymax) + 2;
      int x0 = max(fx0 & ~15, 0), y0 = max(fy0, 0), x1 = min(fx1, W - 1), y1 = min(fy1, H - 1);
      if (x1 >= x0 && y1 >= y0) {
        const int chunks = (x1 - x0 + 16) >> 4;          // 16-byte chunks; x0 and the row pitch P are multiples of 16: a row segment never leaves its row
        const int pc = chunks | 1;                        // odd chunk pitch spreads rows over the banks
        if (chunks <= 16 && (y1 - y0 + 1) * pc * 16 <= CP_BOX_BYTES) {
          bx0 = x0; by0 = y0; bw = chunks * 16; bh = y1 - y0 + 1; pitch = pc * 16;
          clean = (fx0 >= 0 && fy0 >= 0 && fx1 <= W - 1 && fy1 <= H - 1) ? 1 : 0;
        }
      }
    }
    s_geo[0] = bx0; s_geo[1] = by0; s_geo[2] = bw; s_geo[3] = bh; s_geo[4] = pitch; s_geo[5] = clean;
    // tie margin: 4e-7 |q| (> 3x the fp32 quotient error bound) when |q| is known, else the 0.01 that covers |q| < 6e4
    const float qabs = fmaxf(fmaxf(fabsf(xmin), fabsf(xmax)), fmaxf(fabsf(ymin), fabsf(ymax)));
    s_thr = clean ? 0.5f - fmaxf(qabs * 4e-7f + 1e-5f, 1e-4f) : 0.49f;
  }
  __syncthreads();
  const int bx0 = s_geo[0], by0 = s_geo[1], bw = s_geo[2], bh = s_geo[3], pitch = s_geo[4];
  const bool clean = s_geo[5] != 0;
  const float thr = s_thr;
  {
    // a warp copies 2 box rows per step (16 lanes x 16 bytes each); cp.async keeps every load of the box in flight at once
    // (a register-staged loop serialised ~14 L2 round trips per warp and cost half of the kernel's instructions)
    const int lane = tid & 31, ch = lane & 15;
    if (ch * 16 < bw) {
      const uint8_t* src = base + (size_t)by0 * P + bx0 + ch * 16;
      const uint32_t sdst = tc::smem_u32(s_box) + ch * 16;
      for (int r = (tid >> 5) * 2 + (lane >> 4); r < bh; r += 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst + (uint32_t)(r * pitch)), "l"(src + (size_t)r * P) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  const int rg = tid >> 4, c4 = (tid & 15) * 4;
  const float4 w4 = __ldg(reinterpret_cast<const float4*>(map.lin_w + col0 + c4));
  const float whs[4] = {__fmul_rn(w4.x, hs), __fmul_rn(w4.y, hs), __fmul_rn(w4.z, hs), __fmul_rn(w4.w, hs)};
  const float whc[4] = {__fmul_rn(w4.x, hc), __fmul_rn(w4.y, hc), __fmul_rn(w4.z, hc), __fmul_rn(w4.w, hc)};
  uint8_t* dst_tile = packed_crop + ((size_t)crop * 256 + row0) * 256 + col0;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (clean)
    crop_pack_rows<true>(s_box, base, map.lin_l, H, W, P, bx0, by0, bw, bh, pitch, px, py, hc, hs, inv0f, inv1f, thr, whs, whc, row0, rg, c4, s_q, &s_nq, dst_tile);
  else
    crop_pack_rows<false>(s_box, base, map.lin_l, H, W, P, bx0, by0, bw, bh, pitch, px, py, hc, hs, inv0f, inv1f, thr, whs, whc, row0, rg, c4, s_q, &s_nq, dst_tile);
  __syncthreads();     // also orders the word stores above before the byte patches below (same block)
  for (int k = tid; k < s_nq; k += 256) {
    const int rr = s_q[k] >> 6, cc = s_q[k] & 63;
    const float ll = __ldg(map.lin_l + row0 + rr), wc = __ldg(map.lin_w + col0 + cc);
    float gx = __fadd_rn(__fsub_rn(__fmul_rn(ll, hc), __fmul_rn(wc, hs)), px);
    float gy = __fadd_rn(__fadd_rn(__fmul_rn(ll, hs), __fmul_rn(wc, hc)), py);
    if (isnan(gx)) gx = 0.f;
    if (isnan(gy)) gy = 0.f;
    int xp = round_div_exact(gx, dx0, inv0), yp = round_div_exact(gy, dx1, inv1);
    if ((unsigned)yp >= (unsigned)H || (unsigned)xp >= (unsigned)W) { xp = 0; yp = 0; }
    dst_tile[rr * 256 + cc] = __ldg(base + (size_t)((unsigned)yp * (unsigned)P + (unsigned)xp)) & 15u;
  }
}

// Bias passed BY VALUE (constant bank): the epilogue adds it as an immediate-constant operand, no shared-memory traffic.
struct BiasArg {
  float b[64];
};

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor of kind::i8: s8 x s8 -> s32, A and B K-major, dense, no saturation
__host__ __device__ constexpr uint32_t idesc_s8_s32(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16i(uint32_t taddr, int (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld48i(uint32_t taddr, int (&v)[48]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47])
               : "r"(taddr + 32)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// split-phase TMEM loads for the conv1 epilogue pipeline: issue (no wait) / wait / pin the destination registers behind the wait
__device__ __forceinline__ void tmem_ld8i_issue(uint32_t taddr, int* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld_wait_pin(int (&v)[N]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < N; i++) asm volatile("" : "+r"(v[i]));     // no use of a loaded register may be scheduled above the wait
}

__global__ void __launch_bounds__(TC_THREADS) tc_conv1_kernel(const uint8_t* __restrict__ packed_crop, const uint8_t* __restrict__ wpack,
                                                              const BiasArg bias, float* __restrict__ out, double* __restrict__ out_stats, int n) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sP = smem + T1_WBYTES;
  __shared__ __align__(8) uint64_t full[T1_NBUF], empty[T1_NBUF], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < T1_WBYTES / 16; i += TC_THREADS) reinterpret_cast<int4*>(sW)[i] = __ldg(reinterpret_cast<const int4*>(wpack) + i);
  // the patch columns that no item writes (copy 1, columns 70-71) are never read; zero everything once anyway
  for (int i = tid; i < T1_NBUF * T1_PATCH_BYTES / 16; i += TC_THREADS) reinterpret_cast<int4*>(sP)[i] = make_int4(0, 0, 0, 0);
  if (tid == 0) {
    for (int b = 0; b < T1_NBUF; b++) { tc::mbar_init(&full[b], T1_PROD_THREADS / T1_NGRP); tc::mbar_init(&empty[b], 1); }
    for (int a = 0; a < 2; a++) { tc::mbar_init(&acc_full[a], 1); tc::mbar_init(&acc_empty[a], 256); }
    tc::fence_mbar_init();
  }
  if (warp == T1_MMA_WARP) tc::tmem_alloc(&tmem_base, 512);     // 2 accumulator sets (256 columns apart) x 4 sub-tiles x 48 columns
  tc::fence_async_smem();
  STRIVE_PDL_TRIGGER();
  STRIVE_PDL_WAIT_PTRS(packed_crop, out_stats);          // prologue above: weights / GroupNorm affine (constants), barriers, TMEM; below: the previous kernel's output
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;
  constexpr int IPC = T1_RB * T1_CB;     // items per crop
  const int items = n * IPC;
  // contiguous item range per CTA: the items of a crop stay on one SM (L1/L2 locality, one statistics flush per crop)
  const int item_lo = (int)(((long long)items * blockIdx.x) / gridDim.x);
  const int item_hi = (int)(((long long)items * (blockIdx.x + 1)) / gridDim.x);

  if (warp < T1_NPROD) {
    // ---------------- producers: 4 packed pixels (one 32-bit load) -> 4 x int8x4, written to both patch copies ----------------
    // Three groups of two warps take the items round-robin: an item is "load 666 words, wait, expand, store, fence" -- a pure
    // latency chain (fence.proxy.async waits for every outstanding load, so there is no prefetching across the fence) -- and
    // three chains in flight hide it.
    constexpr int QPR = T1_COLS / 4, NQ = T1_ROWS * QPR;                      // 18 quads per row, 666 per item
    constexpr int GT = T1_PROD_THREADS / T1_NGRP;                             // 64 threads per group
    constexpr int QPT = (NQ + GT - 1) / GT;                                   // 11
    const int grp = tid / GT, gt = tid % GT;
    long long tw = 0, t_start = TRACE_T();
    for (int cnt = grp; item_lo + cnt < item_hi; cnt += T1_NGRP) {
      const int item = item_lo + cnt;
      const int crop = item / IPC, st = item % IPC;
      const int row0 = (st / T1_CB) * 32, col0 = (st % T1_CB) * 64;          // first input row / column of the item
      const int b = cnt % T1_NBUF;
      const uint8_t* src = packed_crop + (size_t)crop * 65536 + (size_t)row0 * 256 + col0;
      uint32_t wv[QPT];
#pragma unroll
      for (int k = 0; k < QPT; k++) {
        const int i = gt + k * GT, r = i / QPR, c = i - r * QPR;
        const bool ok = i < NQ && row0 + r < 256 && col0 + 4 * c < 256;          // zero padding outside the crop
        wv[k] = ok ? __ldg(reinterpret_cast<const uint32_t*>(src + r * 256 + 4 * c)) : 0u;
      }
      const long long tq = TRACE_T();
      tc::mbar_wait(&empty[b], ((cnt / T1_NBUF) & 1) ^ 1);
      tw += TRACE_T() - tq;
      uint8_t* dst = sP + (size_t)b * T1_PATCH_BYTES;
#pragma unroll
      for (int k = 0; k < QPT; k++) {
        const int i = gt + k * GT, r = i / QPR, c = i - r * QPR;
        if (i < NQ) {
          uint32_t e[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t v = (wv[k] >> (8 * j)) & 15u;     // bit c = layer c  ->  byte c = 0 / 1
            e[j] = (v & 1u) | ((v & 2u) << 7) | ((v & 4u) << 14) | ((v & 8u) << 21);
          }
          uint8_t* d0 = dst + r * T1_ROWB + c * 16;
          *reinterpret_cast<uint4*>(d0) = make_uint4(e[0], e[1], e[2], e[3]);
          uint8_t* d1 = d0 + T1_COPY;                        // copy 1 column c holds pixel c + 2
          if (c > 0) *reinterpret_cast<uint2*>(d1 - 8) = make_uint2(e[0], e[1]);
          *reinterpret_cast<uint2*>(d1) = make_uint2(e[2], e[3]);
        }
      }
      tc::fence_async_smem();
      tc::mbar_arrive(&full[b]);
    }
    if (tid == 0) { trace_add(0, 0, tw); trace_add(0, 1, TRACE_T() - t_start); trace_add(0, 7, 1); }
  } else if (warp == T1_MMA_WARP) {
    const uint32_t idesc = idesc_s8_s32(128, T1_NP);
    const uint32_t wbase = tc::smem_u32(sW);
    int cnt = 0;
    long long twf = 0, twa = 0, t_start = TRACE_T();
    for (int item = item_lo; item < item_hi; item++, cnt++) {
      const int b = cnt % T1_NBUF, a = cnt & 1;
      const long long tq0 = TRACE_T();
      tc::mbar_wait(&acc_empty[a], ((cnt >> 1) & 1) ^ 1);
      const long long tq1 = TRACE_T();
      tc::mbar_wait(&full[b], (cnt / T1_NBUF) & 1);
      twa += tq1 - tq0;
      twf += TRACE_T() - tq1;
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t pbase = tc::smem_u32(sP + (size_t)b * T1_PATCH_BYTES);
        const uint32_t a_hi = tc::desc_hi(2 * T1_ROWB), b_hi = tc::desc_hi(128);
        const uint32_t b_lo0 = tc::desc_lo(wbase, T1_NP * 16);
#pragma unroll
        for (int sub = 0; sub < 4; sub++) {
          const int xb = sub >> 1, par = sub & 1;
          const uint32_t a_lo0 = tc::desc_lo(pbase + par * T1_COPY + xb * 128, 16);
          const uint32_t d = tm + a * 256 + sub * T1_NP;
#pragma unroll
          for (int ky = 0; ky < 7; ky++) {
            const uint64_t ad = tc::desc_make(a_lo0 + ((ky * T1_ROWB) >> 4), a_hi);
            const uint64_t bd = tc::desc_make(b_lo0 + ((ky * 2 * T1_NP * 16) >> 4), b_hi);
            mma_i8(d, ad, bd, idesc, ky ? 1u : 0u);
          }
        }
        tc::mma_commit(&empty[b]);
        tc::mma_commit(&acc_full[a]);
      }
      __syncwarp();
    }
    if (lane == 0) { trace_add(0, 2, twf); trace_add(0, 3, twa); trace_add(0, 4, TRACE_T() - t_start); }
  } else if (warp >= T1_EPI_WARP0) {
    // ---------------- epilogue: warp (q, half) reads TMEM lanes 32q..32q+31 of sub-tiles 2 half, 2 half + 1 ----------------
    const int q = warp & 3, half = (warp - T1_EPI_WARP0) >> 2;
    const int m = q * 32 + lane, oyl = m >> 3, jx = m & 7;
    float sc[16];
#pragma unroll
    for (int c = 0; c < 16; c++) sc[c] = __ldg(reinterpret_cast<const float*>(wpack + T1_WBYTES) + c);
    int cnt = 0, cur_crop = -1;
    double d1 = 0.0, d2 = 0.0;
    long long twe = 0, t_start = TRACE_T();
    for (int item = item_lo; item < item_hi; item++, cnt++) {
      const int crop = item / IPC, st = item % IPC;
      const int oy0 = (st / T1_CB) * 16, ox0 = (st % T1_CB) * 32;
      const int a = cnt & 1;
      if (crop != cur_crop) {
        if (cur_crop >= 0) {
          d1 = warp_sum_f64(d1);
          d2 = warp_sum_f64(d2);
          if (lane == 0) {
            atomicAdd(out_stats + (size_t)cur_crop * 2, d1);
            atomicAdd(out_stats + (size_t)cur_crop * 2 + 1, d2);
          }
        }
        cur_crop = crop;
        d1 = 0.0;
        d2 = 0.0;
      }
      const long long tq = TRACE_T();
      tc::mbar_wait(&acc_full[a], (cnt >> 1) & 1);
      twe += TRACE_T() - tq;
      tc::tc_fence_after();
      float s1 = 0.f, s2 = 0.f;
      // Four steps per item: (sub-tile par, channel block j) = 8 channels x 3 digit planes = 24 TMEM columns.  The load of step
      // s + 1 is issued right after the wait for step s, so its latency hides behind the recombination and the stores of step s
      // (one 48-column load + wait per sub-tile left this warp idle for the whole TMEM round trip: the epilogue was 93 % busy
      // and the MMA warp waited for accumulators a quarter of the time).
      const uint32_t tb0 = tm + ((uint32_t)(q * 32) << 16) + a * 256 + half * 2 * T1_NP;
      int buf[2][24];
      auto issue = [&](int st_, int (&v)[24]) {
        const uint32_t tb = tb0 + (st_ >> 1) * T1_NP + (st_ & 1) * 8;
        tmem_ld8i_issue(tb, &v[0]);
        tmem_ld8i_issue(tb + 16, &v[8]);
        tmem_ld8i_issue(tb + 32, &v[16]);
      };
      issue(0, buf[0]);
#pragma unroll
      for (int st_ = 0; st_ < 4; st_++) {
        int (&cur)[24] = buf[st_ & 1];
        tmem_ld_wait_pin<24>(cur);
        if (st_ < 3) {
          issue(st_ + 1, buf[(st_ + 1) & 1]);
        } else {
          tc::tc_fence_before();
          tc::mbar_arrive(&acc_empty[a]);
        }
        const int par = st_ >> 1, j = st_ & 1;
        const int oy = oy0 + oyl, ox = ox0 + half * 16 + 2 * jx + par;
        if (oy < 125 && ox < 125) {
          float v[8];
#pragma unroll
          for (int c = 0; c < 8; c++) {
            // |acc_d| <= 196 * 128: the three planes recombine exactly in int32 (|t| < 1.64e9 < 2^31); ONE conversion rounds the
            // exact integer to fp32 (I2F runs on the XU pipe: 16 per output pixel, far below its rate; the magic-number
            // conversions of the separate planes cost 3x the issue slots)
            const int t = (cur[16 + c] * 256 + cur[8 + c]) * 256 + cur[c];
            v[c] = fmaf(__int2float_rn(t), sc[j * 8 + c], bias.b[j * 8 + c]);
            s1 += v[c];
            s2 = fmaf(v[c], v[c], s2);
          }
          // channel-blocked activations [crop][C/8][H][W][8]: a lane stores 32 contiguous bytes per block
          float* dst = out + ((((size_t)crop * 2 + j) * 125 + oy) * 125 + ox) * 8;
          tc::stg256(dst, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
        }
      }
      d1 += (double)s1;
      d2 += (double)s2;
    }
    if (cur_crop >= 0) {
      d1 = warp_sum_f64(d1);
      d2 = warp_sum_f64(d2);
      if (lane == 0) {
        atomicAdd(out_stats + (size_t)cur_crop * 2, d1);
        atomicAdd(out_stats + (size_t)cur_crop * 2 + 1, d2);
      }
    }
    if (warp == T1_EPI_WARP0 && lane == 0) { trace_add(0, 5, twe); trace_add(0, 6, TRACE_T() - t_start); }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == T1_MMA_WARP) {
    __syncwarp();
    tc::tmem_dealloc(tm, 512);
  }
}

// ======================================================================================================
// conv2..4: stride-2 kxk conv, CIN multiple of 16, NCH output channels per CTA (blockIdx.y chunks), NHWC fp32 in/out.
// CTA tile = 16 x 8 outputs (M = 128).  Weights of the chunk stay resident in shared memory; the input tile is staged
// per 16-channel chunk into an NBUF-deep ring.  Per filter tap and 16 channels TWO MMAs:
//     D[:, 0:2NCH]  += A_hi * [W_hi | W_lo]^T      (N = 2 NCH: hi and lo weights stacked along N, A_hi fetched once)
//     D[:, 0:NCH]   += A_lo * W_hi^T               (N = NCH)
// and the epilogue adds the two column halves.  Measured SS-mode cost of one M=128,K=16 MMA is (4096 + 32 N) / 128 cycles
// (scripts/mma_bench.cu): the A-tile fetch dominates, so stacking N is worth 27 % over three N = NCH MMAs.
// ======================================================================================================
// (24-warp CTAs with 19 producer warps were measured SLOWER: the producers are bound by shared-memory store issue, not latency)
#define T2_NPROD 11
#define T2_MMA_WARP 11
#define T2_EPI_WARP0 12
#define T2_THREADS 512
#define T2_PROD_THREADS (T2_NPROD * 32)
template <int CIN, int KS, int HIN, int HOUT, int COUT, int NCH_, int NBUF>
struct TcCfg {
  static constexpr int NCH = NCH_;
  static constexpr int C2 = CIN / 16;
  static constexpr int TAPS = KS * KS;
  static constexpr int PH = 30 + KS;
  static constexpr int PW = 14 + KS;
  static constexpr int PQ = 8 + (KS - 1) / 2;
  static constexpr int A_PREC_BYTES = 2 * PH * 2 * PQ * 16;
  static constexpr int A_BYTES = 2 * A_PREC_BYTES;
  static constexpr int TAP_BYTES = 64 * NCH;                 // [khalf 2][prec 2][NCH rows][8 k] bf16
  static constexpr int W_BYTES = C2 * TAPS * TAP_BYTES;
  static constexpr int TILES_Y = (HOUT + 15) / 16;
  static constexpr int TILES_X = (HOUT + 7) / 8;
  static constexpr int TILES = TILES_Y * TILES_X;
  static constexpr int TMEM_COLS = 4 * NCH;                  // 2 accumulator sets x (hi part | lo part)
  static constexpr int DUMP_OFF = (2 * PQ - 1) * 16;         // last slot of the odd-column plane of patch row 0: never written by a real
                                                             // pixel (PW is odd) nor read by a tap window -> dump slot for the idle work items
  static constexpr size_t SMEM = (size_t)W_BYTES + (size_t)NBUF * A_BYTES;
};

template <int CIN, int KS, int HIN, int HOUT, int COUT, int NCH, int NBUF, bool OUT_BLK>
__global__ void __launch_bounds__(T2_THREADS) tc_conv_kernel(const float* __restrict__ in, const double* __restrict__ in_stats,
                                                             const float* __restrict__ gam, const float* __restrict__ bet,
                                                             const uint8_t* __restrict__ wpack, const BiasArg bias,
                                                             float* __restrict__ out, double* __restrict__ out_stats, int n, int dbg_arg) {
  const int dbg = STRIVE_TC_DEBUG ? dbg_arg : 0;     // timing-experiment flags: compiled out of the product build
  using Cfg = TcCfg<CIN, KS, HIN, HOUT, COUT, NCH, NBUF>;
  constexpr int PH = Cfg::PH, PW = Cfg::PW, PQ = Cfg::PQ, C2 = Cfg::C2, TAPS = Cfg::TAPS;
  constexpr int TK = CIN == 16 ? 1 : (CIN == 32 ? 2 : 3);
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sA = smem + Cfg::W_BYTES;
  __shared__ __align__(8) uint64_t full[NBUF], empty[NBUF], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(16) float s_gam[CIN], s_bet[CIN];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunk = (COUT == NCH) ? 0 : (int)blockIdx.y;
  {
    const int4* src = reinterpret_cast<const int4*>(wpack + (size_t)nchunk * Cfg::W_BYTES);
    for (int i = tid; i < Cfg::W_BYTES / 16; i += T2_THREADS) reinterpret_cast<int4*>(sW)[i] = __ldg(src + i);
  }
  for (int i = tid; i < CIN; i += T2_THREADS) { s_gam[i] = gam[i]; s_bet[i] = bet[i]; }
  if (tid == 0) {
    for (int b = 0; b < NBUF; b++) { tc::mbar_init(&full[b], T2_PROD_THREADS); tc::mbar_init(&empty[b], 1); }
    for (int a = 0; a < 2; a++) { tc::mbar_init(&acc_full[a], 1); tc::mbar_init(&acc_empty[a], 128); }
    tc::fence_mbar_init();
  }
  if (warp == T2_MMA_WARP) tc::tmem_alloc(&tmem_base, Cfg::TMEM_COLS);
  tc::fence_async_smem();
  STRIVE_PDL_TRIGGER();
  STRIVE_PDL_WAIT_PTRS(in, in_stats);          // prologue above: weights / GroupNorm affine (constants), barriers, TMEM; below: the previous kernel's output
  const int items = n * Cfg::TILES;
  // contiguous item range per CTA: consecutive tiles of the same crop share the GroupNorm statistics and L2 lines
  const int item_lo = (int)(((long long)items * blockIdx.x) / gridDim.x);
  const int item_hi = (int)(((long long)items * (blockIdx.x + 1)) / gridDim.x);
  const int crop_first = item_lo / Cfg::TILES;
  __shared__ float s_cm[GN_TAB], s_cr[GN_TAB];
  gn_table_fill(s_cm, s_cr, in_stats, crop_first, (item_hi - 1) / Cfg::TILES, (double)CIN * HIN * HIN, tid);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;

  if (warp < T2_NPROD) {
    // ---------------- producers ----------------
    // Work item = 8 channels (one channel block, 32 bytes) of one input pixel; activations are channel-blocked
    // [crop][C/8][H][W][8].  Thread t owns items t, t + 352, ...: its block half (t & 1) is fixed, so its GroupNorm affine
    // lives in 16 registers, and the (row, col) of each of its items inside the input patch never changes: global and
    // shared offsets are computed once.  A warp reads two 512-byte runs per 256-bit load and writes one 16-byte operand row
    // per item (8 distinct 16-byte bank groups per quarter-warp: conflict-free).
    constexpr int NPIX = PH * PW, NITEM = NPIX * 2;
    constexpr int KI = (NITEM + T2_PROD_THREADS - 1) / T2_PROD_THREADS;
    constexpr int CG = PH * 2 * PQ * 16;
    const int half = tid & 1;
    int rc[KI], soff[KI];
#pragma unroll
    for (int k = 0; k < KI; k++) {
      const int i = tid + k * T2_PROD_THREADS;
      rc[k] = -1;
      soff[k] = Cfg::DUMP_OFF;
      if (i < NITEM) {
        const int p = i >> 1;
        const int row = p / PW, col = p - row * PW;
        rc[k] = (row << 8) | col;
        soff[k] = ((row * 2 + (col & 1)) * PQ + (col >> 1)) * 16 + half * CG;
      }
    }
    auto load_chunk = [&](int item, int c2, float (&x)[KI][8], unsigned& okmask) {
      const int crop = item / Cfg::TILES, tile = item - crop * Cfg::TILES;
      const int ty0 = (tile / Cfg::TILES_X) * 16, tx0 = (tile % Cfg::TILES_X) * 8;
      const int rows_valid = HIN - 2 * ty0, cols_valid = HIN - 2 * tx0;
      const float* base = in + ((((size_t)crop * (CIN / 8) + c2 * 2 + half) * HIN + 2 * ty0) * HIN + 2 * tx0) * 8;
      okmask = 0u;
#pragma unroll
      for (int k = 0; k < KI; k++) {
        const int row = rc[k] >> 8, col = rc[k] & 255;
        if (rc[k] >= 0 && row < rows_valid && col < cols_valid) {
          okmask |= 1u << k;
          if (!(dbg & 4)) tc::ldg256(base + (row * HIN + col) * 8, x[k]);
        }
      }
    };
    int item = item_lo, c2 = 0, cnt = 0, cur_crop = -1;
    long long tw = 0, t_start = TRACE_T();
    float xn[KI][8];
    unsigned okn = 0u;
    float ga[8], gb[8];
    float cmean = 0.f, crstd = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) { ga[j] = 0.f; gb[j] = 0.f; }
    if (item < item_hi) load_chunk(item, c2, xn, okn);
    while (item < item_hi) {
      float xc[KI][8];
#pragma unroll
      for (int k = 0; k < KI; k++)
#pragma unroll
        for (int j = 0; j < 8; j++) xc[k][j] = xn[k][j];
      const unsigned okc = okn;
      const int ci = item, cc2 = c2;
      if (++c2 == C2) { c2 = 0; item++; }
      if (item < item_hi) load_chunk(item, c2, xn, okn);
      const int crop = ci / Cfg::TILES;
      const bool new_crop = crop != cur_crop;
      if (new_crop) {
        // GroupNorm statistics of this crop from the CTA's table (no producer-wide barrier, no float64 arithmetic in the loop)
        cur_crop = crop;
        gn_table_get(s_cm, s_cr, in_stats, crop, crop_first, (double)CIN * HIN * HIN, cmean, crstd);
      }
      if (C2 > 1 || new_crop) {
        //  y = relu(x * ga + gb) == relu((x - mean) * rstd * gamma + beta)
#pragma unroll
        for (int j = 0; j < 8; j += 4) {
          const float4 g4 = *reinterpret_cast<const float4*>(&s_gam[cc2 * 16 + half * 8 + j]);
          const float4 b4 = *reinterpret_cast<const float4*>(&s_bet[cc2 * 16 + half * 8 + j]);
          ga[j] = crstd * g4.x; ga[j + 1] = crstd * g4.y; ga[j + 2] = crstd * g4.z; ga[j + 3] = crstd * g4.w;
          gb[j] = fmaf(-cmean, ga[j], b4.x); gb[j + 1] = fmaf(-cmean, ga[j + 1], b4.y);
          gb[j + 2] = fmaf(-cmean, ga[j + 2], b4.z); gb[j + 3] = fmaf(-cmean, ga[j + 3], b4.w);
        }
      }
      const int b = cnt % NBUF;
      const long long tq = TRACE_T();
      tc::mbar_wait(&empty[b], ((cnt / NBUF) & 1) ^ 1);
      tw += TRACE_T() - tq;
      uint8_t* dst = sA + (size_t)b * Cfg::A_BYTES;
      // transform everything first, then store everything: a store's source registers stay locked until the (congested)
      // shared-memory pipe has read them, so interleaving would serialise the arithmetic behind the stores
      uint4 hi[KI], lo[KI];
#pragma unroll
      for (int k = 0; k < KI; k++) {
        const bool okk = (okc >> k) & 1u;
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; j++) y[j] = okk ? fmaxf(fmaf(xc[k][j], ga[j], gb[j]), 0.f) : 0.f;
        tc::split_pack2(y[0], y[1], hi[k].x, lo[k].x);
        tc::split_pack2(y[2], y[3], hi[k].y, lo[k].y);
        tc::split_pack2(y[4], y[5], hi[k].z, lo[k].z);
        tc::split_pack2(y[6], y[7], hi[k].w, lo[k].w);
      }
      if (!(dbg & 2))
#pragma unroll
      for (int k = 0; k < KI; k++) {
        *reinterpret_cast<uint4*>(dst + soff[k]) = hi[k];                            // items beyond the patch land in the dump slot
        *reinterpret_cast<uint4*>(dst + Cfg::A_PREC_BYTES + soff[k]) = lo[k];
      }
      tc::fence_async_smem();
      tc::mbar_arrive(&full[b]);
      cnt++;
    }
    if (tid == 0) { trace_add(TK, 0, tw); trace_add(TK, 1, TRACE_T() - t_start); trace_add(TK, 7, 1); }
  } else if (warp == T2_MMA_WARP) {
    const uint32_t idesc1 = tc::idesc_bf16_f32(128, 2 * NCH), idesc2 = tc::idesc_bf16_f32(128, NCH);
    constexpr uint32_t LBO_A = PH * 2 * PQ * 16, SBO_A = 64 * PQ, LBO_B = 32 * NCH;
    int cnt = 0, it = 0;
    long long twf = 0, twa = 0, t_start = TRACE_T();
    for (int item = item_lo; item < item_hi; item++, it++) {
      const int a = it & 1;
      const long long tq0 = TRACE_T();
      tc::mbar_wait(&acc_empty[a], ((it >> 1) & 1) ^ 1);
      twa += TRACE_T() - tq0;
      const uint32_t d = tm + a * (2 * NCH);
#pragma unroll 1
      for (int c2 = 0; c2 < C2; c2++, cnt++) {
        const int b = cnt % NBUF;
        const long long tq1 = TRACE_T();
        tc::mbar_wait(&full[b], (cnt / NBUF) & 1);
        twf += TRACE_T() - tq1;
        tc::tc_fence_after();
        if (tc::elect_one()) {
          const uint32_t a_lo0 = tc::desc_lo(tc::smem_u32(sA + (size_t)b * Cfg::A_BYTES), LBO_A);
          const uint32_t b_lo0 = tc::desc_lo(tc::smem_u32(sW) + c2 * TAPS * Cfg::TAP_BYTES, LBO_B);
          const uint32_t a_hi = tc::desc_hi(SBO_A), b_hi = tc::desc_hi(128);
#pragma unroll
          for (int ky = 0; ky < KS; ky++) {
#pragma unroll
            for (int kx = 0; kx < KS; kx++) {
              const int tap = ky * KS + kx;
              const uint32_t al0 = a_lo0 + ((((ky * 2 + (kx & 1)) * PQ + (kx >> 1)) * 16) >> 4);
              const uint64_t ah = tc::desc_make(al0, a_hi), al = tc::desc_make(al0 + (Cfg::A_PREC_BYTES >> 4), a_hi);
              const uint64_t bd = tc::desc_make(b_lo0 + ((tap * Cfg::TAP_BYTES) >> 4), b_hi);
              if (!(dbg & 8)) {
                tc::mma_bf16(d, ah, bd, idesc1, (tap > 0 || c2 > 0) ? 1u : 0u);
                tc::mma_bf16(d, al, bd, idesc2, 1u);
              }
            }
          }
          tc::mma_commit(&empty[b]);
          if (c2 == C2 - 1) tc::mma_commit(&acc_full[a]);
        }
        __syncwarp();
      }
    }
    if (lane == 0) { trace_add(TK, 2, twf); trace_add(TK, 3, twa); trace_add(TK, 4, TRACE_T() - t_start); }
  } else {
    // ---------------- epilogue: TMEM -> registers -> (+bias) -> global, no shared memory ----------------
    const int q = warp - T2_EPI_WARP0;
    const int m = q * 32 + lane;
    int it = 0, cur_crop = -1;
    double d1 = 0.0, d2 = 0.0;
    long long twe = 0, t_start = TRACE_T();
    for (int item = item_lo; item < item_hi; item++, it++) {
      const int crop = item / Cfg::TILES, tile = item % Cfg::TILES;
      const int ty0 = (tile / Cfg::TILES_X) * 16, tx0 = (tile % Cfg::TILES_X) * 8;
      const int a = it & 1;
      if (crop != cur_crop) {
        if (cur_crop >= 0) {
          d1 = warp_sum_f64(d1);
          d2 = warp_sum_f64(d2);
          if (lane == 0) {
            atomicAdd(out_stats + (size_t)cur_crop * 2, d1);
            atomicAdd(out_stats + (size_t)cur_crop * 2 + 1, d2);
          }
        }
        cur_crop = crop;
        d1 = 0.0;
        d2 = 0.0;
      }
      const long long tq = TRACE_T();
      tc::mbar_wait(&acc_full[a], (it >> 1) & 1);
      twe += TRACE_T() - tq;
      tc::tc_fence_after();
      const int oy = ty0 + (m >> 3), ox = tx0 + (m & 7);
      const bool ok = oy < HOUT && ox < HOUT;
      const uint32_t tbase = tm + ((uint32_t)(q * 32) << 16) + a * (2 * NCH);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int h = 0; h < NCH / 16; h++) {
        float vh[16], vl[16];
        tc::tmem_ld16(tbase + h * 16, vh);
        tc::tmem_ld16(tbase + NCH + h * 16, vl);
        if (h == NCH / 16 - 1) {
          tc::tc_fence_before();
          tc::mbar_arrive(&acc_empty[a]);
        }
        if (ok) {
#pragma unroll
          for (int c = 0; c < 16; c++) {
            const float bv = (COUT == NCH) ? bias.b[h * 16 + c] : bias.b[nchunk * NCH + h * 16 + c];
            vh[c] = (vh[c] + vl[c]) + bv;
            s1 += vh[c];
            s2 = fmaf(vh[c], vh[c], s2);
          }
          if (!(dbg & 1)) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
              const int ch0 = nchunk * NCH + h * 16 + j * 8;
              // OUT_BLK: channel-blocked [crop][C/8][H][W][8] (a warp stores 4 rows x 256 contiguous bytes per instruction);
              // otherwise plain NHWC for the GEMM-style conv5 kernel
              float* dst = OUT_BLK ? out + ((((size_t)crop * (COUT / 8) + (ch0 >> 3)) * HOUT + oy) * HOUT + ox) * 8
                                   : out + (((size_t)crop * HOUT + oy) * HOUT + ox) * COUT + ch0;
              tc::stg256(dst, vh[j * 8], vh[j * 8 + 1], vh[j * 8 + 2], vh[j * 8 + 3], vh[j * 8 + 4], vh[j * 8 + 5], vh[j * 8 + 6], vh[j * 8 + 7]);
            }
          }
        }
      }
      d1 += (double)s1;
      d2 += (double)s2;
    }
    if (cur_crop >= 0) {
      d1 = warp_sum_f64(d1);
      d2 = warp_sum_f64(d2);
      if (lane == 0) {
        atomicAdd(out_stats + (size_t)cur_crop * 2, d1);
        atomicAdd(out_stats + (size_t)cur_crop * 2 + 1, d2);
      }
    }
    if (q == 0 && lane == 0) { trace_add(TK, 5, twe); trace_add(TK, 6, TRACE_T() - t_start); }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == T2_MMA_WARP) {
    __syncwarp();
    tc::tmem_dealloc(tm, Cfg::TMEM_COLS);
  }
}

// ======================================================================================================
// conv3 (32 -> 64, k5 s2): all 64 output channels in ONE CTA, so the MMAs run at N = 128 ([W_hi | W_lo] stacked, the SS-mode
// rate reaches the math floor at N >= 128: scripts/mma_bench2.cu) and every input tile is staged once instead of once per
// 32-channel chunk.  The hi/lo weights of both 16-channel K chunks are 204.8 KB -- more than shared memory -- so only ONE
// chunk (102.4 KB) is resident and the K chunks are the OUTER loop over a pair of tiles whose accumulators wait in TMEM:
//     pair p:  (t0, c) (t1, c)   [weights -> chunk 1-c]   (t0, 1-c) (t1, 1-c)        c = parity of the GLOBAL pair index
// so the weights are swapped once per pair and the next pair starts with the chunk that is already resident.  The swap is
// two cp.async.bulk halves (taps 0-12 / 13-24), each issued by the loader warp as soon as the MMAs that read the old half
// have completed (tcgen05.commit -> w_free[h]), i.e. the first half streams in while the tensor core still works on taps
// 13-24.  4 accumulators of 128 columns = all 512 TMEM columns: the epilogue of pair p overlaps the MMAs of pair p + 1.
// Roles: warps 0-9 producers, warp 10 weight loader, warp 11 MMA issuer, warps 12-15 epilogue.
// ======================================================================================================
#define T3_NPROD 10
#define T3_PROD_THREADS (T3_NPROD * 32)
#define T3_LOAD_WARP 10
#define T3_NBUF 2
#define T3_WCHUNK (25 * 4096)
#define T3_H0_TAPS 13

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
// adds `bytes` to the pending transaction count of the current phase WITHOUT arriving (the caller arrives later, with everybody else)
__device__ __forceinline__ void mbar_expect_tx_only(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc),
               "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(T2_THREADS) tc_conv3_kernel(const float* __restrict__ in, const double* __restrict__ in_stats,
                                                              const float* __restrict__ gam, const float* __restrict__ bet,
                                                              const uint8_t* __restrict__ wpack, const BiasArg bias,
                                                              float* __restrict__ out, double* __restrict__ out_stats, int n) {
  using Cfg = TcCfg<32, 5, 61, 29, 64, 64, T3_NBUF>;
  constexpr int CIN = 32, KS = 5, HIN = 61, HOUT = 29, COUT = 64, NCH = 64, NBUF = T3_NBUF;
  constexpr int PH = Cfg::PH, PW = Cfg::PW, PQ = Cfg::PQ, TAPS = Cfg::TAPS;
  static_assert(Cfg::TAP_BYTES == 4096 && Cfg::C2 == 2, "conv3 weight chunk layout");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sA = smem + T3_WCHUNK;
  __shared__ __align__(8) uint64_t full[NBUF], empty[NBUF], acc_full[4], acc_empty[4], w_full[2], w_free[2];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(16) float s_gam[CIN], s_bet[CIN];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // Contiguous range of tile PAIRS per CTA.  The K-chunk order of a pair follows the parity of its GLOBAL pair index: a crop has
  // 8 tiles = 4 pairs, so the order in which a tile's two K chunks are accumulated depends only on the tile's place inside its
  // crop -- never on the batch composition or the CTA the pair lands on (bitwise batch invariance of the forward pass).
  static_assert(Cfg::TILES % 4 == 0, "conv3: pairs of a crop must not straddle crops and their parity must repeat per crop");
  const int pairs_total = n * (Cfg::TILES / 2);
  const int pair_lo = (int)(((long long)pairs_total * blockIdx.x) / gridDim.x);
  const int pair_hi = (int)(((long long)pairs_total * (blockIdx.x + 1)) / gridDim.x);
  const int par0 = pair_lo & 1;
  {
    const int4* src = reinterpret_cast<const int4*>(wpack + (size_t)par0 * T3_WCHUNK);     // the first pair's first K chunk is resident first
    for (int i = tid; i < T3_WCHUNK / 16; i += T2_THREADS) reinterpret_cast<int4*>(sW)[i] = __ldg(src + i);
  }
  for (int i = tid; i < CIN; i += T2_THREADS) { s_gam[i] = gam[i]; s_bet[i] = bet[i]; }
  if (tid == 0) {
    for (int b = 0; b < NBUF; b++) { tc::mbar_init(&full[b], T3_PROD_THREADS); tc::mbar_init(&empty[b], 1); }
    for (int a = 0; a < 4; a++) { tc::mbar_init(&acc_full[a], 1); tc::mbar_init(&acc_empty[a], 128); }
    for (int h = 0; h < 2; h++) { tc::mbar_init(&w_full[h], 1); tc::mbar_init(&w_free[h], 1); }
    tc::fence_mbar_init();
  }
  if (warp == T2_MMA_WARP) tc::tmem_alloc(&tmem_base, 512);
  tc::fence_async_smem();
  STRIVE_PDL_TRIGGER();
  STRIVE_PDL_WAIT_PTRS(in, in_stats);          // prologue above: weights / GroupNorm affine (constants), barriers, TMEM; below: the previous kernel's output
  const int item_lo = 2 * pair_lo, item_hi = 2 * pair_hi;
  const int npairs = pair_hi - pair_lo;
  const int crop_first = item_lo / Cfg::TILES;
  __shared__ float s_cm[GN_TAB], s_cr[GN_TAB];
  gn_table_fill(s_cm, s_cr, in_stats, crop_first, (item_hi - 1) / Cfg::TILES, (double)CIN * HIN * HIN, tid);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;

  if (warp < T3_NPROD) {
    // ---------------- producers: same work items as tc_conv_kernel (8 channels of one input pixel), 320 threads ----------------
    constexpr int NPIX = PH * PW, NITEM = NPIX * 2;
    constexpr int KI = (NITEM + T3_PROD_THREADS - 1) / T3_PROD_THREADS;
    constexpr int CG = PH * 2 * PQ * 16;
    const int half = tid & 1;
    int rc[KI], soff[KI];
#pragma unroll
    for (int k = 0; k < KI; k++) {
      const int i = tid + k * T3_PROD_THREADS;
      rc[k] = -1;
      soff[k] = Cfg::DUMP_OFF;
      if (i < NITEM) {
        const int p = i >> 1;
        const int row = p / PW, col = p - row * PW;
        rc[k] = (row << 8) | col;
        soff[k] = ((row * 2 + (col & 1)) * PQ + (col >> 1)) * 16 + half * CG;
      }
    }
    int cur_crop = -1;
    float cmean = 0.f, crstd = 0.f;
    long long tw = 0, t_start = TRACE_T();
    const int njobs = 2 * (item_hi - item_lo);
    // job j of this CTA -> (tile, K chunk): pairs of tiles, K chunk outer (see the header comment)
    auto decode = [&](int j, int& item, int& c2) {
      const int p = j >> 2, r = j & 3;
      const int np = min(2, item_hi - (item_lo + 2 * p));
      const int s2 = np == 2 ? (r >> 1) : r, i2 = np == 2 ? (r & 1) : 0;
      item = item_lo + 2 * p + i2;
      c2 = ((p + par0) & 1) ^ s2;
    };
    auto job_base = [&](int item, int c2, int& rows_valid, int& cols_valid) -> const float* {
      const int crop = item / Cfg::TILES, tile = item - crop * Cfg::TILES;
      const int ty0 = (tile / Cfg::TILES_X) * 16, tx0 = (tile % Cfg::TILES_X) * 8;
      rows_valid = HIN - 2 * ty0;
      cols_valid = HIN - 2 * tx0;
      return in + ((((size_t)crop * (CIN / 8) + c2 * 2 + half) * HIN + 2 * ty0) * HIN + 2 * tx0) * 8;
    };
    for (int cnt = 0; cnt < njobs; cnt++) {
      int item, c2;
      decode(cnt, item, c2);
      const int crop = item / Cfg::TILES;
      {
        {
          int rows_valid, cols_valid;
          const float* base = job_base(item, c2, rows_valid, cols_valid);
          float x[KI][8];
          unsigned ok = 0u;
#pragma unroll
          for (int k = 0; k < KI; k++) {
            const int row = rc[k] >> 8, col = rc[k] & 255;
            if (rc[k] >= 0 && row < rows_valid && col < cols_valid) {
              ok |= 1u << k;
              tc::ldg256(base + (row * HIN + col) * 8, x[k]);
            }
          }
          if (cnt + 1 < njobs) {
            // pull the next job's input towards L2 (no registers, no scoreboard: fence.proxy.async below waits for every
            // outstanding register load, so a register prefetch would serialise behind it)
            int nitem, nc2, nrv, ncv;
            decode(cnt + 1, nitem, nc2);
            const float* nbase = job_base(nitem, nc2, nrv, ncv);
#pragma unroll
            for (int k = 0; k < KI; k++) {
              const int row = rc[k] >> 8, col = rc[k] & 255;
              if (rc[k] >= 0 && row < nrv && col < ncv) asm volatile("prefetch.global.L2 [%0];" ::"l"(nbase + (row * HIN + col) * 8));
            }
          }
          if (crop != cur_crop) {     // statistics from the CTA's table, no producer barrier (see tc_conv_kernel)
            cur_crop = crop;
            gn_table_get(s_cm, s_cr, in_stats, crop, crop_first, (double)CIN * HIN * HIN, cmean, crstd);
          }
          float ga[8], gb[8];
#pragma unroll
          for (int j = 0; j < 8; j += 4) {
            const float4 g4 = *reinterpret_cast<const float4*>(&s_gam[c2 * 16 + half * 8 + j]);
            const float4 b4 = *reinterpret_cast<const float4*>(&s_bet[c2 * 16 + half * 8 + j]);
            ga[j] = crstd * g4.x; ga[j + 1] = crstd * g4.y; ga[j + 2] = crstd * g4.z; ga[j + 3] = crstd * g4.w;
            gb[j] = fmaf(-cmean, ga[j], b4.x); gb[j + 1] = fmaf(-cmean, ga[j + 1], b4.y);
            gb[j + 2] = fmaf(-cmean, ga[j + 2], b4.z); gb[j + 3] = fmaf(-cmean, ga[j + 3], b4.w);
          }
          const int b = cnt % NBUF;
          const long long tq = TRACE_T();
          tc::mbar_wait(&empty[b], ((cnt / NBUF) & 1) ^ 1);
          tw += TRACE_T() - tq;
          uint8_t* dst = sA + (size_t)b * Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < KI; k++) {
            const bool okk = (ok >> k) & 1u;
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; j++) y[j] = okk ? fmaxf(fmaf(x[k][j], ga[j], gb[j]), 0.f) : 0.f;
            uint4 hi, lo;
            tc::split_pack2(y[0], y[1], hi.x, lo.x);
            tc::split_pack2(y[2], y[3], hi.y, lo.y);
            tc::split_pack2(y[4], y[5], hi.z, lo.z);
            tc::split_pack2(y[6], y[7], hi.w, lo.w);
            *reinterpret_cast<uint4*>(dst + soff[k]) = hi;                            // items beyond the patch land in the dump slot
            *reinterpret_cast<uint4*>(dst + Cfg::A_PREC_BYTES + soff[k]) = lo;
          }
          tc::fence_async_smem();
          tc::mbar_arrive(&full[b]);
        }
      }
    }
    if (tid == 0) { trace_add(2, 0, tw); trace_add(2, 1, TRACE_T() - t_start); trace_add(2, 7, 1); }
  } else if (warp == T3_LOAD_WARP) {
    // ---------------- weight loader: one swap per pair, two halves, each as soon as its old contents are dead ----------------
    if (tc::elect_one()) {
      for (int p = 0; p < npairs; p++) {
        const uint8_t* src = wpack + (size_t)(1 - ((p + par0) & 1)) * T3_WCHUNK;
        tc::mbar_wait(&w_free[0], p & 1);
        mbar_expect_tx(&w_full[0], T3_H0_TAPS * 4096);
        for (int t = 0; t < T3_H0_TAPS; t++) bulk_g2s(sW + t * 4096, src + t * 4096, 4096, &w_full[0]);
        tc::mbar_wait(&w_free[1], p & 1);
        mbar_expect_tx(&w_full[1], (TAPS - T3_H0_TAPS) * 4096);
        for (int t = T3_H0_TAPS; t < TAPS; t++) bulk_g2s(sW + t * 4096, src + t * 4096, 4096, &w_full[1]);
      }
    }
    __syncwarp();
  } else if (warp == T2_MMA_WARP) {
    const uint32_t idesc1 = tc::idesc_bf16_f32(128, 2 * NCH), idesc2 = tc::idesc_bf16_f32(128, NCH);
    constexpr uint32_t LBO_A = PH * 2 * PQ * 16, SBO_A = 64 * PQ, LBO_B = 32 * NCH;
    int cnt = 0;
    long long twf = 0, twa = 0, t_start = TRACE_T();
    for (int p = 0; p < npairs; p++) {
      const int np = min(2, item_hi - (item_lo + 2 * p));
      for (int s2 = 0; s2 < 2; s2++) {
        for (int i2 = 0; i2 < np; i2++, cnt++) {
          const int k = (p & 1) * 2 + i2;
          const uint32_t d = tm + k * (2 * NCH);
          if (s2 == 0) {
            const long long tq0 = TRACE_T();
            tc::mbar_wait(&acc_empty[k], ((p >> 1) & 1) ^ 1);
            twa += TRACE_T() - tq0;
          }
          const int b = cnt % NBUF;
          const long long tq1 = TRACE_T();
          tc::mbar_wait(&full[b], (cnt / NBUF) & 1);
          twf += TRACE_T() - tq1;
          const bool after_swap = (s2 == 1 && i2 == 0), before_swap = (s2 == 0 && i2 == np - 1);
          if (after_swap) tc::mbar_wait(&w_full[0], p & 1);
          tc::tc_fence_after();
          const uint32_t a_lo0 = tc::desc_lo(tc::smem_u32(sA + (size_t)b * Cfg::A_BYTES), LBO_A);
          const uint32_t b_lo0 = tc::desc_lo(tc::smem_u32(sW), LBO_B);
          const uint32_t a_hi = tc::desc_hi(SBO_A), b_hi = tc::desc_hi(128);
          if (tc::elect_one()) {
#pragma unroll
            for (int tap = 0; tap < T3_H0_TAPS; tap++) {
              const int ky = tap / KS, kx = tap % KS;
              const uint32_t al0 = a_lo0 + ((((ky * 2 + (kx & 1)) * PQ + (kx >> 1)) * 16) >> 4);
              const uint64_t ah = tc::desc_make(al0, a_hi), al = tc::desc_make(al0 + (Cfg::A_PREC_BYTES >> 4), a_hi);
              const uint64_t bd = tc::desc_make(b_lo0 + ((tap * Cfg::TAP_BYTES) >> 4), b_hi);
              tc::mma_bf16(d, ah, bd, idesc1, (tap > 0 || s2 > 0) ? 1u : 0u);
              tc::mma_bf16(d, al, bd, idesc2, 1u);
            }
            if (before_swap) tc::mma_commit(&w_free[0]);
          }
          __syncwarp();
          if (after_swap) {
            tc::mbar_wait(&w_full[1], p & 1);
            tc::tc_fence_after();
          }
          if (tc::elect_one()) {
#pragma unroll
            for (int tap = T3_H0_TAPS; tap < TAPS; tap++) {
              const int ky = tap / KS, kx = tap % KS;
              const uint32_t al0 = a_lo0 + ((((ky * 2 + (kx & 1)) * PQ + (kx >> 1)) * 16) >> 4);
              const uint64_t ah = tc::desc_make(al0, a_hi), al = tc::desc_make(al0 + (Cfg::A_PREC_BYTES >> 4), a_hi);
              const uint64_t bd = tc::desc_make(b_lo0 + ((tap * Cfg::TAP_BYTES) >> 4), b_hi);
              tc::mma_bf16(d, ah, bd, idesc1, 1u);
              tc::mma_bf16(d, al, bd, idesc2, 1u);
            }
            if (before_swap) tc::mma_commit(&w_free[1]);
            tc::mma_commit(&empty[b]);
            if (s2 == 1) tc::mma_commit(&acc_full[k]);
          }
          __syncwarp();
        }
      }
    }
    if (lane == 0) { trace_add(2, 2, twf); trace_add(2, 3, twa); trace_add(2, 4, TRACE_T() - t_start); }
  } else {
    // ---------------- epilogue: TMEM -> registers -> (+bias) -> global (channel-blocked), fp64 GroupNorm statistics ----------------
    const int q = warp - T2_EPI_WARP0;
    const int m = q * 32 + lane;
    int cur_crop = -1;
    double d1 = 0.0, d2 = 0.0;
    long long twe = 0, t_start = TRACE_T();
    for (int p = 0; p < npairs; p++) {
      const int np = min(2, item_hi - (item_lo + 2 * p));
      for (int i2 = 0; i2 < np; i2++) {
        const int item = item_lo + 2 * p + i2;
        const int crop = item / Cfg::TILES, tile = item % Cfg::TILES;
        const int ty0 = (tile / Cfg::TILES_X) * 16, tx0 = (tile % Cfg::TILES_X) * 8;
        const int k = (p & 1) * 2 + i2;
        if (crop != cur_crop) {
          if (cur_crop >= 0) {
            d1 = warp_sum_f64(d1);
            d2 = warp_sum_f64(d2);
            if (lane == 0) {
              atomicAdd(out_stats + (size_t)cur_crop * 2, d1);
              atomicAdd(out_stats + (size_t)cur_crop * 2 + 1, d2);
            }
          }
          cur_crop = crop;
          d1 = 0.0;
          d2 = 0.0;
        }
        const long long tq = TRACE_T();
        tc::mbar_wait(&acc_full[k], (p >> 1) & 1);
        twe += TRACE_T() - tq;
        tc::tc_fence_after();
        const int oy = ty0 + (m >> 3), ox = tx0 + (m & 7);
        const bool ok = oy < HOUT && ox < HOUT;
        const uint32_t tbase = tm + ((uint32_t)(q * 32) << 16) + k * (2 * NCH);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int h = 0; h < NCH / 16; h++) {
          float vh[16], vl[16];
          tc::tmem_ld16(tbase + h * 16, vh);
          tc::tmem_ld16(tbase + NCH + h * 16, vl);
          if (h == NCH / 16 - 1) {
            tc::tc_fence_before();
            tc::mbar_arrive(&acc_empty[k]);
          }
          if (ok) {
#pragma unroll
            for (int c = 0; c < 16; c++) {
              vh[c] = (vh[c] + vl[c]) + bias.b[h * 16 + c];
              s1 += vh[c];
              s2 = fmaf(vh[c], vh[c], s2);
            }
#pragma unroll
            for (int j = 0; j < 2; j++) {
              const int ch0 = h * 16 + j * 8;
              float* dst = out + ((((size_t)crop * (COUT / 8) + (ch0 >> 3)) * HOUT + oy) * HOUT + ox) * 8;
              tc::stg256(dst, vh[j * 8], vh[j * 8 + 1], vh[j * 8 + 2], vh[j * 8 + 3], vh[j * 8 + 4], vh[j * 8 + 5], vh[j * 8 + 6], vh[j * 8 + 7]);
            }
          }
        }
        d1 += (double)s1;
        d2 += (double)s2;
      }
    }
    if (cur_crop >= 0) {
      d1 = warp_sum_f64(d1);
      d2 = warp_sum_f64(d2);
      if (lane == 0) {
        atomicAdd(out_stats + (size_t)cur_crop * 2, d1);
        atomicAdd(out_stats + (size_t)cur_crop * 2 + 1, d2);
      }
    }
    if (q == 0 && lane == 0) { trace_add(2, 5, twe); trace_add(2, 6, TRACE_T() - t_start); }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == T2_MMA_WARP) {
    __syncwarp();
    tc::tmem_dealloc(tm, 512);
  }
}

// ======================================================================================================
// conv3 on CTA PAIRS (tcgen05 cta_group::2).  A cluster of two CTAs on one TPC works on one crop at a time: the leader (rank 0)
// owns tiles 0-3, the follower tiles 4-7, each as two tile pairs with the K-chunk-outer schedule of tc_conv3_kernel above (the
// order in which a tile accumulates its two K chunks is the same function of the tile's place in its crop: batch-invariant).
// ONE M = 256 MMA covers a leader tile and a follower tile, and every B operand is split across the pair: CTA r stages, per tap and
// K chunk, ONE 64-row block  Z_r = [W_hi rows 32r..32r+31 ; W_lo rows 32r..32r+31]  (2 KB instead of the 4 KB [W_hi ; W_lo]):
//     D[:, 0:128]  += A_hi * [Z_0 ; Z_1]^T                  N = 128: columns  hi(0:32) | lo(0:32) | hi(32:64) | lo(32:64)
//     D[:, 32:96]  += A_lo * [Z_0[0:32] ; Z_1[0:32]]^T      N = 64 on the SAME block: the W_hi rows, landing on lo(0:32) | hi(32:64)
// and the epilogue adds column c to column c + 32 (channels 0-31) resp. 64 + c' to 96 + c' (channels 32-63).  An SM thus fetches
// 3 KB instead of 6 KB of weights per tap and K chunk through its 128 B/clk data path, and a chunk is 50 KB per CTA: taps 0-20 of
// BOTH chunks stay resident, only taps 21-24 (8 KB) are swapped once per tile pair -- behind the 21 resident taps of the next job --
// and three operand buffers fit.  (Small terms are summed apart from the hi x hi products here, so the results differ from the
// single-CTA kernel in the last bits; both are batch-invariant.)
// Only the leader issues MMAs.  Cross-CTA signalling:
//   full[b], acc_empty[k], w_full live in the LEADER: local arrivals + one remote arrival per follower warp / loader
//   empty[b], acc_full[k], w_free are signalled by tcgen05.commit multicast to the same barrier in BOTH CTAs
// ======================================================================================================
#define T3P_NBUF 3
#define T3P_TAP_BYTES 2048
#define T3P_CHUNK_BYTES (25 * T3P_TAP_BYTES)                 // one K chunk of one rank in the weight pack
#define T3P_H0_TAPS 21
#define T3P_R0_BYTES (T3P_H0_TAPS * T3P_TAP_BYTES)           // taps 0..20 of one chunk (resident for both chunks)
#define T3P_R1_BYTES ((25 - T3P_H0_TAPS) * T3P_TAP_BYTES)    // taps 21..24 of the current chunk (swapped)
#define T3P_WBYTES (2 * T3P_R0_BYTES + T3P_R1_BYTES)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrival on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(tc::smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// wait on a local barrier that also receives arrivals from the peer CTA (cluster-scope acquire)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t a = tc::smem_u32(bar);
  uint32_t ok = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(a), "r"(parity), "r"(4000u)
                 : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc),
               "l"(bdesc), "r"(idesc), "r"(accumulate)
               : "memory");
}
// completion of all MMAs issued so far -> one arrival on `bar` in BOTH CTAs of the pair
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(tc::smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T2_THREADS) tc_conv3_pair_kernel(const float* __restrict__ in, const double* __restrict__ in_stats,
                                                                                             const float* __restrict__ gam, const float* __restrict__ bet,
                                                                                             const uint8_t* __restrict__ wpack, const BiasArg bias,
                                                                                             float* __restrict__ out, double* __restrict__ out_stats, int n) {
  using Cfg = TcCfg<32, 5, 61, 29, 64, 64, T3P_NBUF>;
  constexpr int CIN = 32, KS = 5, HIN = 61, HOUT = 29, COUT = 64, NCH = 64, NBUF = T3P_NBUF;
  constexpr int PH = Cfg::PH, PW = Cfg::PW, PQ = Cfg::PQ, TAPS = Cfg::TAPS;
  static_assert(Cfg::TILES == 8, "conv3 pair kernel: a crop is 4 tiles per CTA = 2 tile pairs of alternating parity");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;                          // [taps 0-20 of chunk 0][taps 0-20 of chunk 1][taps 21-24 of the current chunk]
  uint8_t* sW1 = smem + 2 * T3P_R0_BYTES;
  uint8_t* sA = smem + T3P_WBYTES;
  __shared__ __align__(8) uint64_t full[NBUF], empty[NBUF], acc_full[4], acc_empty[4], w_full, w_free;
  __shared__ uint32_t tmem_base;
  __shared__ __align__(16) float s_gam[CIN], s_bet[CIN];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int ncl = (int)(gridDim.x >> 1), cl = (int)(blockIdx.x >> 1);
  const int crop_lo = (int)(((long long)n * cl) / ncl), crop_hi = (int)(((long long)n * (cl + 1)) / ncl);
  const int npairs = 2 * (crop_hi - crop_lo);            // local pair p: crop crop_lo + p / 2, tiles 4 rank + 2 (p & 1) + {0, 1}; parity p & 1
  const uint8_t* wrank = wpack + (size_t)rank * 2 * T3P_CHUNK_BYTES;
  {
    // resident: taps 0-20 of both chunks; the first pair starts with K chunk 0: its taps 21-24
    for (int i = tid; i < T3P_WBYTES / 16; i += T2_THREADS) {
      const int byte = i * 16;
      const int src = byte < T3P_R0_BYTES ? byte : (byte < 2 * T3P_R0_BYTES ? T3P_CHUNK_BYTES + (byte - T3P_R0_BYTES) : T3P_R0_BYTES + (byte - 2 * T3P_R0_BYTES));
      reinterpret_cast<int4*>(sW)[i] = __ldg(reinterpret_cast<const int4*>(wrank + src));
    }
  }
  for (int i = tid; i < CIN; i += T2_THREADS) { s_gam[i] = gam[i]; s_bet[i] = bet[i]; }
  if (tid == 0) {
    for (int b = 0; b < NBUF; b++) { tc::mbar_init(&full[b], leader ? T3_PROD_THREADS / 2 + T3_NPROD / 2 : 1); tc::mbar_init(&empty[b], 1); }
    for (int a = 0; a < 4; a++) { tc::mbar_init(&acc_full[a], 1); tc::mbar_init(&acc_empty[a], leader ? 128 + 4 : 1); }
    tc::mbar_init(&w_full, leader ? 2 : 1);
    tc::mbar_init(&w_free, 1);
    tc::fence_mbar_init();
  }
  if (warp == T2_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc::fence_async_smem();
  STRIVE_PDL_TRIGGER();
  STRIVE_PDL_WAIT_PTRS(in, in_stats);          // prologue above: weights / GroupNorm affine (constants), barriers, TMEM; below: the previous kernel's output
  __shared__ float s_cm[GN_TAB], s_cr[GN_TAB];
  gn_table_fill(s_cm, s_cr, in_stats, crop_lo, crop_hi - 1, (double)CIN * HIN * HIN, tid);
  tc::tc_fence_before();
  __syncthreads();
  cluster_sync_all();                          // the peer's barriers are initialised before anything arrives on them
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;
  auto item_of = [&](int p, int i2) { return (crop_lo + (p >> 1)) * Cfg::TILES + 4 * (int)rank + 2 * (p & 1) + i2; };

  if (warp < T3_NPROD) {
    // ---------------- producers: TWO groups of five warps take the jobs alternately ----------------
    // A job is "load the patch, wait for it, transform, store, fence" -- the proxy fence waits for every outstanding load of the
    // thread, so one group cannot prefetch across jobs; with two groups one job's load latency hides behind the other job's
    // arithmetic (with the B bytes halved the MMAs no longer cover for it: one group left the tensor core starved 30 % of the time).
    // Work item = 8 channels (one channel block, 32 bytes) of one input pixel, 9 item slots per thread in three batches.
    constexpr int GT = T3_PROD_THREADS / 2;                                    // 160 threads per group
    constexpr int NPIX = PH * PW, NITEM = NPIX * 2;
    constexpr int KI = (NITEM + GT - 1) / GT;                                  // 9
    constexpr int KB = 3;                                                      // item slots per load batch
    static_assert(KI == 3 * KB, "conv3 pair producers: three batches of three item slots");
    constexpr int CG = PH * 2 * PQ * 16;
    const int grp = tid / GT, gt = tid - grp * GT;
    const int half = tid & 1;
    // one register per item slot: (operand-buffer offset << 12) | (patch row << 6) | patch column; row 63 = no item (never valid)
    static_assert(PH < 63 && PW < 64 && Cfg::A_PREC_BYTES < (1 << 19), "item slot packing");
    uint32_t pk[KI];
#pragma unroll
    for (int k = 0; k < KI; k++) {
      const int i = gt + k * GT;
      pk[k] = ((uint32_t)Cfg::DUMP_OFF << 12) | (63u << 6);
      if (i < NITEM) {
        const int p = i >> 1;
        const int row = p / PW, col = p - row * PW;
        pk[k] = ((uint32_t)(((row * 2 + (col & 1)) * PQ + (col >> 1)) * 16 + half * CG) << 12) | ((uint32_t)row << 6) | (uint32_t)col;
      }
    }
    int cur_crop = -1;
    float cmean = 0.f, crstd = 0.f;
    const int njobs = 4 * npairs;
    auto decode = [&](int j, int& item, int& c2) {       // job j -> (tile, K chunk): pairs of tiles, K chunk outer
      const int p = j >> 2, r = j & 3;
      item = item_of(p, r & 1);
      c2 = (p & 1) ^ (r >> 1);
    };
    auto job_base = [&](int item, int c2, int& rows_valid, int& cols_valid) -> const float* {
      const int crop = item / Cfg::TILES, tile = item - crop * Cfg::TILES;
      const int ty0 = (tile / Cfg::TILES_X) * 16, tx0 = (tile % Cfg::TILES_X) * 8;
      rows_valid = HIN - 2 * ty0;
      cols_valid = HIN - 2 * tx0;
      return in + ((((size_t)crop * (CIN / 8) + c2 * 2 + half) * HIN + 2 * ty0) * HIN + 2 * tx0) * 8;
    };
    for (int cnt = grp; cnt < njobs; cnt += 2) {
      int item, c2;
      decode(cnt, item, c2);
      const int crop = item / Cfg::TILES;
      int rows_valid, cols_valid;
      const float* base = job_base(item, c2, rows_valid, cols_valid);
      if (crop != cur_crop) {
        cur_crop = crop;
        gn_table_get(s_cm, s_cr, in_stats, crop, crop_lo, (double)CIN * HIN * HIN, cmean, crstd);
      }
      float ga[8], gb[8];
#pragma unroll
      for (int j = 0; j < 8; j += 4) {
        const float4 g4 = *reinterpret_cast<const float4*>(&s_gam[c2 * 16 + half * 8 + j]);
        const float4 b4 = *reinterpret_cast<const float4*>(&s_bet[c2 * 16 + half * 8 + j]);
        ga[j] = crstd * g4.x; ga[j + 1] = crstd * g4.y; ga[j + 2] = crstd * g4.z; ga[j + 3] = crstd * g4.w;
        gb[j] = fmaf(-cmean, ga[j], b4.x); gb[j + 1] = fmaf(-cmean, ga[j + 1], b4.y);
        gb[j + 2] = fmaf(-cmean, ga[j + 2], b4.z); gb[j + 3] = fmaf(-cmean, ga[j + 3], b4.w);
      }
      const int b = cnt % NBUF;
      uint8_t* dst = sA + (size_t)b * Cfg::A_BYTES;
      // three batches of three item slots: the loads of batch i + 1 are in flight while batch i is transformed (inside a job nothing
      // fences; only the end-of-job proxy fence waits for outstanding loads)
      float xa[KB][8], xb[KB][8];          // two rotating batch buffers
      unsigned ok = 0u;
      auto load_batch = [&](int bt, float (&x)[KB][8]) {
#pragma unroll
        for (int k = 0; k < KB; k++) {
          const uint32_t e = pk[bt * KB + k];
          const int row = (int)((e >> 6) & 63u), col = (int)(e & 63u);
          if (row < rows_valid && col < cols_valid) {
            ok |= 1u << (bt * KB + k);
            tc::ldg256(base + (row * HIN + col) * 8, x[k]);
          }
        }
      };
      auto transform_batch = [&](int bt, const float (&x)[KB][8]) {
#pragma unroll
        for (int k = 0; k < KB; k++) {
          const uint32_t e = pk[bt * KB + k];
          const bool okk = (ok >> (bt * KB + k)) & 1u;
          float y[8];
#pragma unroll
          for (int j = 0; j < 8; j++) y[j] = okk ? fmaxf(fmaf(x[k][j], ga[j], gb[j]), 0.f) : 0.f;
          uint4 hi, lo;
          tc::split_pack2(y[0], y[1], hi.x, lo.x);
          tc::split_pack2(y[2], y[3], hi.y, lo.y);
          tc::split_pack2(y[4], y[5], hi.z, lo.z);
          tc::split_pack2(y[6], y[7], hi.w, lo.w);
          *reinterpret_cast<uint4*>(dst + (e >> 12)) = hi;                          // items beyond the patch land in the dump slot
          *reinterpret_cast<uint4*>(dst + Cfg::A_PREC_BYTES + (e >> 12)) = lo;
        }
      };
      load_batch(0, xa);
      load_batch(1, xb);
      if (cnt + 2 < njobs) {      // this group's next job towards L2
        int nitem, nc2, nrv, ncv;
        decode(cnt + 2, nitem, nc2);
        const float* nbase = job_base(nitem, nc2, nrv, ncv);
#pragma unroll
        for (int k = 0; k < KI; k++) {
          const int row = (int)((pk[k] >> 6) & 63u), col = (int)(pk[k] & 63u);
          if (row < nrv && col < ncv) asm volatile("prefetch.global.L2 [%0];" ::"l"(nbase + (row * HIN + col) * 8));
        }
      }
      tc::mbar_wait(&empty[b], ((cnt / NBUF) & 1) ^ 1);
      transform_batch(0, xa);
      load_batch(2, xa);
      transform_batch(1, xb);
      transform_batch(2, xa);
      tc::fence_async_smem();
      if (leader) {
        tc::mbar_arrive(&full[b]);
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&full[b], 0);
      }
    }
    if (tid == 0 && leader) trace_add(2, 7, 1);
  } else if (warp == T3_LOAD_WARP) {
    // ---------------- weight loader: taps 21-24 of this CTA's half of the other K chunk, once per pair ----------------
    if (tc::elect_one()) {
      for (int p = 0; p < npairs; p++) {
        const uint8_t* src = wrank + (size_t)(1 - (p & 1)) * T3P_CHUNK_BYTES + T3P_R0_BYTES;
        tc::mbar_wait(&w_free, p & 1);
        mbar_expect_tx(&w_full, T3P_R1_BYTES);
        bulk_g2s(sW1, src, T3P_R1_BYTES, &w_full);
        if (!leader) {
          tc::mbar_wait(&w_full, p & 1);
          mbar_arrive_cluster(&w_full, 0);
        }
      }
    }
    __syncwarp();
  } else if (warp == T2_MMA_WARP) {
    if (leader) {
      const uint32_t idesc1 = tc::idesc_bf16_f32(256, 2 * NCH), idesc2 = tc::idesc_bf16_f32(256, NCH);
      constexpr uint32_t LBO_A = PH * 2 * PQ * 16, SBO_A = 64 * PQ;
      int cnt = 0;
      long long twf = 0, twa = 0, tww = 0, t_start = TRACE_T();
      for (int p = 0; p < npairs; p++) {
        for (int s2 = 0; s2 < 2; s2++) {
          const int cc = (p & 1) ^ s2;                     // K chunk of these two jobs
          for (int i2 = 0; i2 < 2; i2++, cnt++) {
            const int k = (p & 1) * 2 + i2;
            const uint32_t d = tm + k * (2 * NCH);
            const long long tq0 = TRACE_T();
            if (s2 == 0) mbar_wait_cluster(&acc_empty[k], ((p >> 1) & 1) ^ 1);
            const int b = cnt % NBUF;
            const long long tq1 = TRACE_T();
            mbar_wait_cluster(&full[b], (cnt / NBUF) & 1);
            twa += tq1 - tq0; twf += TRACE_T() - tq1;
            const bool after_swap = (s2 == 1 && i2 == 0), before_swap = (s2 == 0 && i2 == 1);
            tc::tc_fence_after();
            const uint32_t a_lo0 = tc::desc_lo(tc::smem_u32(sA + (size_t)b * Cfg::A_BYTES), LBO_A);
            const uint32_t b0_lo0 = tc::desc_lo(tc::smem_u32(sW) + cc * T3P_R0_BYTES, 1024), b1_lo0 = tc::desc_lo(tc::smem_u32(sW1), 1024);
            const uint32_t a_hi = tc::desc_hi(SBO_A), b_hi = tc::desc_hi(128);
            if (tc::elect_one()) {
#pragma unroll
              for (int tap = 0; tap < T3P_H0_TAPS; tap++) {
                const int ky = tap / KS, kx = tap % KS;
                const uint32_t al0 = a_lo0 + ((((ky * 2 + (kx & 1)) * PQ + (kx >> 1)) * 16) >> 4);
                const uint64_t ah = tc::desc_make(al0, a_hi), al = tc::desc_make(al0 + (Cfg::A_PREC_BYTES >> 4), a_hi);
                const uint64_t bd = tc::desc_make(b0_lo0 + ((tap * T3P_TAP_BYTES) >> 4), b_hi);
                mma_bf16_pair(d, ah, bd, idesc1, (tap > 0 || s2 > 0) ? 1u : 0u);
                mma_bf16_pair(d + 32, al, bd, idesc2, 1u);
              }
            }
            __syncwarp();
            if (after_swap) {
              const long long tq3 = TRACE_T();
              mbar_wait_cluster(&w_full, p & 1);
              tww += TRACE_T() - tq3;
              tc::tc_fence_after();
            }
            if (tc::elect_one()) {
#pragma unroll
              for (int tap = T3P_H0_TAPS; tap < TAPS; tap++) {
                const int ky = tap / KS, kx = tap % KS;
                const uint32_t al0 = a_lo0 + ((((ky * 2 + (kx & 1)) * PQ + (kx >> 1)) * 16) >> 4);
                const uint64_t ah = tc::desc_make(al0, a_hi), al = tc::desc_make(al0 + (Cfg::A_PREC_BYTES >> 4), a_hi);
                const uint64_t bd = tc::desc_make(b1_lo0 + (((tap - T3P_H0_TAPS) * T3P_TAP_BYTES) >> 4), b_hi);
                mma_bf16_pair(d, ah, bd, idesc1, 1u);
                mma_bf16_pair(d + 32, al, bd, idesc2, 1u);
              }
              if (before_swap) mma_commit_pair(&w_free);
              mma_commit_pair(&empty[b]);
              if (s2 == 1) mma_commit_pair(&acc_full[k]);
            }
            __syncwarp();
          }
        }
      }
      if (lane == 0) { trace_add(2, 2, twf); trace_add(2, 3, twa); trace_add(2, 4, TRACE_T() - t_start); trace_add(2, 5, tww); }
    }
  } else if (warp >= T2_EPI_WARP0) {
    // ---------------- epilogue: as in tc_conv3_kernel; the follower reports each drained accumulator to the leader ----------------
    const int q = warp - T2_EPI_WARP0;
    const int m = q * 32 + lane;
    int cur_crop = -1;
    double d1 = 0.0, d2 = 0.0;
    for (int p = 0; p < npairs; p++) {
      for (int i2 = 0; i2 < 2; i2++) {
        const int item = item_of(p, i2);
        const int crop = item / Cfg::TILES, tile = item % Cfg::TILES;
        const int ty0 = (tile / Cfg::TILES_X) * 16, tx0 = (tile % Cfg::TILES_X) * 8;
        const int k = (p & 1) * 2 + i2;
        if (crop != cur_crop) {
          if (cur_crop >= 0) {
            d1 = warp_sum_f64(d1);
            d2 = warp_sum_f64(d2);
            if (lane == 0) {
              atomicAdd(out_stats + (size_t)cur_crop * 2, d1);
              atomicAdd(out_stats + (size_t)cur_crop * 2 + 1, d2);
            }
          }
          cur_crop = crop;
          d1 = 0.0;
          d2 = 0.0;
        }
        tc::mbar_wait(&acc_full[k], (p >> 1) & 1);
        tc::tc_fence_after();
        const int oy = ty0 + (m >> 3), ox = tx0 + (m & 7);
        const bool ok = oy < HOUT && ox < HOUT;
        const uint32_t tbase = tm + ((uint32_t)(q * 32) << 16) + k * (2 * NCH);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int h = 0; h < NCH / 16; h++) {
          float vh[16], vl[16];
          const int colh = h * 16 < 32 ? h * 16 : 32 + h * 16;     // channel block -> column of its hi x hi sums; the small terms sit 32 columns on
          tc::tmem_ld16(tbase + colh, vh);
          tc::tmem_ld16(tbase + colh + 32, vl);
          if (h == NCH / 16 - 1) {
            tc::tc_fence_before();
            if (leader) {
              tc::mbar_arrive(&acc_empty[k]);
            } else {
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster(&acc_empty[k], 0);
            }
          }
          if (ok) {
#pragma unroll
            for (int c = 0; c < 16; c++) {
              vh[c] = (vh[c] + vl[c]) + bias.b[h * 16 + c];
              s1 += vh[c];
              s2 = fmaf(vh[c], vh[c], s2);
            }
#pragma unroll
            for (int j = 0; j < 2; j++) {
              const int ch0 = h * 16 + j * 8;
              float* dst = out + ((((size_t)crop * (COUT / 8) + (ch0 >> 3)) * HOUT + oy) * HOUT + ox) * 8;
              tc::stg256(dst, vh[j * 8], vh[j * 8 + 1], vh[j * 8 + 2], vh[j * 8 + 3], vh[j * 8 + 4], vh[j * 8 + 5], vh[j * 8 + 6], vh[j * 8 + 7]);
            }
          }
        }
        d1 += (double)s1;
        d2 += (double)s2;
      }
    }
    if (cur_crop >= 0) {
      d1 = warp_sum_f64(d1);
      d2 = warp_sum_f64(d2);
      if (lane == 0) {
        atomicAdd(out_stats + (size_t)cur_crop * 2, d1);
        atomicAdd(out_stats + (size_t)cur_crop * 2 + 1, d2);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // no CTA leaves (or frees tensor memory) while the peer may still signal its barriers or its MMAs write here
  if (warp == T2_MMA_WARP) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
  }
}

// ======================================================================================================
// conv5 / conv6 / fc: small spatial extent (6x6, 2x2, 1x1 outputs) -> rows of many crops are packed into M = 128 tiles and
// the A operand is gathered explicitly (im2col rows written straight into the canonical K-major layout, 64-wide K chunks).
// GEMM:  out[m][n] = sum_k relu(GN(in))[row m, tap(k), c(k)] * W[n][k],   k = tap * CIN + c,  NHWC in/out.
// ======================================================================================================
template <int CIN, int KS, int HIN, int HOUT, int COUT, bool FINAL>
struct Tc3Cfg {
  static constexpr int PIX = HOUT * HOUT;
  static constexpr int K = KS * KS * CIN;
  static constexpr int NCH = K / 64;
  static constexpr int A_PREC = 128 * 64 * 2;
  static constexpr int W_PREC = COUT * 64 * 2;
  static constexpr int STAGE = 2 * A_PREC + 2 * W_PREC;
  static constexpr int NBUF = 3;
  static constexpr size_t SMEM = (size_t)NBUF * STAGE;
};

template <int CIN, int KS, int HIN, int HOUT, int COUT, bool FINAL>
__global__ void __launch_bounds__(TC_THREADS) tc_gemm_kernel(const float* __restrict__ in, const double* __restrict__ in_stats,
                                                             const float* __restrict__ gam, const float* __restrict__ bet,
                                                             const uint8_t* __restrict__ wpack, const float* __restrict__ bias,
                                                             float* __restrict__ out, double* __restrict__ out_stats, int n, int dbg_arg) {
  const int dbg = STRIVE_TC_DEBUG ? dbg_arg : 0;     // timing-experiment flags: compiled out of the product build
  using Cfg = Tc3Cfg<CIN, KS, HIN, HOUT, COUT, FINAL>;
  constexpr int PIX = Cfg::PIX, NCH = Cfg::NCH, NBUF = Cfg::NBUF;
  static_assert(CIN % 64 == 0 && COUT % 32 == 0 && 2 * COUT <= 512, "tc_gemm tiling");
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[NBUF], empty[NBUF], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base;
  __shared__ float s_gam[CIN], s_bet[CIN], s_bias[COUT];
  __shared__ float s_mean[2][128], s_rstd[2][128];
  __shared__ __align__(16) float s_stage[4][32 * 36];
  __shared__ int s_off[2][128];     // element offset of the row's (2oy, 2ox) input pixel, -1 = row beyond the batch
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < CIN; i += TC_THREADS) { s_gam[i] = gam[i]; s_bet[i] = bet[i]; }
  for (int i = tid; i < COUT; i += TC_THREADS) s_bias[i] = bias[i];
  if (tid == 0) {
    for (int b = 0; b < NBUF; b++) { tc::mbar_init(&full[b], TC_PROD_THREADS); tc::mbar_init(&empty[b], 1); }
    for (int a = 0; a < 2; a++) { tc::mbar_init(&acc_full[a], 1); tc::mbar_init(&acc_empty[a], 128); }
    tc::fence_mbar_init();
  }
  constexpr uint32_t TCOLS = 2 * COUT;   // 256 or 128: a power of two >= 32
  if (warp == TC_MMA_WARP) tc::tmem_alloc(&tmem_base, TCOLS);
  STRIVE_PDL_TRIGGER();
  STRIVE_PDL_WAIT_PTRS(in, in_stats);          // prologue above: weights / GroupNorm affine (constants), barriers, TMEM; below: the previous kernel's output
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;
  const long long rows_total = (long long)n * PIX;
  const int tiles = (int)((rows_total + 127) / 128);

  if (warp < TC_NPROD) {
    int cnt = 0, tpar = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, tpar ^= 1) {
      if (tid < 128) {
        const long long gm = (long long)tile * 128 + tid;
        float mean = 0.f, rstd = 0.f;
        int off = -1;
        if (gm < rows_total) {
          const int crop = (int)(gm / PIX), pix = (int)(gm % PIX);
          gn_stats(in_stats, crop, (double)CIN * HIN * HIN, mean, rstd);
          off = ((crop * HIN + 2 * (pix / HOUT)) * HIN + 2 * (pix % HOUT)) * CIN;
        }
        s_mean[tpar][tid] = mean; s_rstd[tpar][tid] = rstd; s_off[tpar][tid] = off;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(TC_PROD_THREADS) : "memory");   // producers only
#pragma unroll 1
      for (int kc = 0; kc < NCH; kc++, cnt++) {
        const int b = cnt % NBUF;
        uint8_t* sA = smem + (size_t)b * Cfg::STAGE;
        uint8_t* sWt = sA + 2 * Cfg::A_PREC;
        const int k0 = kc * 64;
        const int tap = k0 / CIN, c0 = k0 % CIN;
        const int ky = tap / KS, kx = tap % KS;
        // every global load of the chunk is issued before anything waits on one (the old load -> use loops paid ~9 L2 round
        // trips per chunk and made these kernels 10x slower than their MMAs): the im2col rows go to registers now, the weight
        // chunk goes through cp.async once the ring slot is free
        constexpr int NIT = (128 * 8 + TC_PROD_THREADS - 1) / TC_PROD_THREADS;     // 3 (row, 8-channel group) items per thread
        float4 t0[NIT], t1[NIT];
        int offs[NIT];
#pragma unroll
        for (int j = 0; j < NIT; j++) {
          const int idx = tid + j * TC_PROD_THREADS;
          offs[j] = idx < 128 * 8 ? s_off[tpar][idx & 127] : -2;
          if (offs[j] >= 0) {
            const float* src = in + (size_t)offs[j] + (ky * HIN + kx) * CIN + c0 + (idx >> 7) * 8;
            if (dbg & 4) { t0[j] = make_float4(1.f, 1.f, 1.f, 1.f); t1[j] = t0[j]; continue; }
            t0[j] = __ldg(reinterpret_cast<const float4*>(src));
            t1[j] = __ldg(reinterpret_cast<const float4*>(src + 4));
          }
        }
        tc::mbar_wait(&empty[b], ((cnt / NBUF) & 1) ^ 1);
        if (tid == 0 && !(dbg & 32)) {
          // the chunk's hi/lo weights are contiguous in the pack: ONE bulk copy whose bytes are counted on the stage's `full`
          // barrier (the per-thread cp.async loop it replaces held a quarter of this kernel's stall samples)
          mbar_expect_tx_only(&full[b], 2 * Cfg::W_PREC);
          bulk_g2s(sWt, wpack + (size_t)kc * 2 * Cfg::W_PREC, 2 * Cfg::W_PREC, &full[b]);
        }
#pragma unroll
        for (int j = 0; j < NIT; j++) {
          const int idx = tid + j * TC_PROD_THREADS;
          if (offs[j] == -2) continue;
          const int m = idx & 127, kg = idx >> 7;
          uint32_t hi[4] = {0u, 0u, 0u, 0u}, lo[4] = {0u, 0u, 0u, 0u};
          if (offs[j] >= 0) {
            const float mean = s_mean[tpar][m], rstd = s_rstd[tpar][m];
            const float x[8] = {t0[j].x, t0[j].y, t0[j].z, t0[j].w, t1[j].x, t1[j].y, t1[j].z, t1[j].w};
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
              const int ch = c0 + kg * 8 + c;
              const float a0 = rstd * s_gam[ch], a1 = rstd * s_gam[ch + 1];
              const float y0 = fmaxf(fmaf(x[c], a0, fmaf(-mean, a0, s_bet[ch])), 0.f);
              const float y1 = fmaxf(fmaf(x[c + 1], a1, fmaf(-mean, a1, s_bet[ch + 1])), 0.f);
              tc::split_pack2(y0, y1, hi[c >> 1], lo[c >> 1]);
            }
          }
          const int unit = (kg * 16 + (m >> 3)) * 8 + (m & 7);
          if (dbg & 2) continue;
          *reinterpret_cast<uint4*>(sA + (size_t)unit * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(sA + Cfg::A_PREC + (size_t)unit * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        tc::fence_async_smem();
        tc::mbar_arrive(&full[b]);
      }
    }
  } else if (warp == TC_MMA_WARP) {
    {
      const uint32_t idesc = tc::idesc_bf16_f32(128, COUT);
      constexpr uint32_t LBO_W = (COUT / 8) * 128;
      int cnt = 0, it = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, it++) {
        const int a = it & 1;
        tc::mbar_wait(&acc_empty[a], ((it >> 1) & 1) ^ 1);
        const uint32_t d = tm + a * COUT;
#pragma unroll 1
        for (int kc = 0; kc < NCH; kc++, cnt++) {
          const int b = cnt % NBUF;
          tc::mbar_wait(&full[b], (cnt / NBUF) & 1);
          tc::tc_fence_after();
          if (tc::elect_one()) {
          const uint32_t abase = tc::smem_u32(smem + (size_t)b * Cfg::STAGE);
          const uint32_t a_lo0 = tc::desc_lo(abase, 2048), b_lo0 = tc::desc_lo(abase + 2 * Cfg::A_PREC, LBO_W);
          const uint32_t a_hi = tc::desc_hi(128), b_hi = tc::desc_hi(128);
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint64_t ah = tc::desc_make(a_lo0 + ((j * 2 * 2048) >> 4), a_hi);
            const uint64_t al = tc::desc_make(a_lo0 + ((Cfg::A_PREC + j * 2 * 2048) >> 4), a_hi);
            const uint64_t bh = tc::desc_make(b_lo0 + ((j * 2 * LBO_W) >> 4), b_hi);
            const uint64_t bl = tc::desc_make(b_lo0 + ((Cfg::W_PREC + j * 2 * LBO_W) >> 4), b_hi);
            if (dbg & 8) continue;
            tc::mma_bf16(d, ah, bh, idesc, (j > 0 || kc > 0) ? 1u : 0u);
            tc::mma_bf16(d, al, bh, idesc, 1u);
            tc::mma_bf16(d, ah, bl, idesc, 1u);
          }
          tc::mma_commit(&empty[b]);
          if (kc == NCH - 1) tc::mma_commit(&acc_full[a]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    const int q = warp - TC_EPI_WARP0;
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, it++) {
      const int a = it & 1;
      const long long gm = (long long)tile * 128 + q * 32 + lane;
      const bool valid = gm < rows_total;
      tc::mbar_wait(&acc_full[a], (it >> 1) & 1);
      tc::tc_fence_after();
      float s1 = 0.f, s2 = 0.f;
      float* stg = s_stage[q];
      const long long gm0 = (long long)tile * 128 + q * 32;     // first row of this warp
#pragma unroll 1
      for (int h = 0; h < COUT / 32; h++) {
        float v0[16], v1[16];
        tc::tmem_ld16(tm + ((uint32_t)(q * 32) << 16) + a * COUT + h * 32, v0);
        tc::tmem_ld16(tm + ((uint32_t)(q * 32) << 16) + a * COUT + h * 32 + 16, v1);
        if (h == COUT / 32 - 1) {
          tc::tc_fence_before();
          tc::mbar_arrive(&acc_empty[a]);
        }
#pragma unroll
        for (int c = 0; c < 16; c++) {
          v0[c] += s_bias[h * 32 + c];
          v1[c] += s_bias[h * 32 + 16 + c];
          if (valid) {
            s1 += v0[c] + v1[c];
            s2 = fmaf(v0[c], v0[c], s2);
            s2 = fmaf(v1[c], v1[c], s2);
          }
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
          *reinterpret_cast<float4*>(stg + lane * 36 + c) = make_float4(v0[c], v0[c + 1], v0[c + 2], v0[c + 3]);
          *reinterpret_cast<float4*>(stg + lane * 36 + 16 + c) = make_float4(v1[c], v1[c + 1], v1[c + 2], v1[c + 3]);
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int e = lane + 32 * j;
          const int rl = e >> 3, ch = (e & 7) * 4;
          if (gm0 + rl < rows_total)
            *reinterpret_cast<float4*>(out + (size_t)(gm0 + rl) * COUT + h * 32 + ch) = *reinterpret_cast<const float4*>(stg + rl * 36 + ch);
        }
      }
      if (!FINAL && valid) {
        const int crop = (int)(gm / PIX);
        atomicAdd(out_stats + (size_t)crop * 2, (double)s1);
        atomicAdd(out_stats + (size_t)crop * 2 + 1, (double)s2);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP) {
    __syncwarp();
    tc::tmem_dealloc(tm, TCOLS);
  }
}

// ------------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------------
// test hook (strive_map_crop): the production crop_pack kernel, unpacked to the reference's (N,4,256,256) uint8 layout
__global__ void crop_unpack_kernel(const uint8_t* __restrict__ packed_crop, uint8_t* __restrict__ out, int n, int C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * 65536) return;
  const size_t crop = i >> 16, pix = i & 65535;
  const unsigned b = packed_crop[i];
  for (int c = 0; c < C; c++) out[(crop * C + c) * 65536 + pix] = (b >> c) & 1u;
}
int tc_crop_pack_unpacked(const StriveMap* map, const float* pose, const int32_t* map_of, int n, uint8_t* out, cudaStream_t stream) {
  STRIVE_CHECK(map->packed != nullptr && map->C <= 4, STRIVE_EINVAL, "crop_pack needs StriveMap.packed and <= 4 layers");
  STRIVE_CHECK(map->packed_pitch >= map->W && (map->packed_pitch & 15) == 0 && ((uintptr_t)map->packed & 15) == 0, STRIVE_EINVAL, "StriveMap.packed_pitch must be >= W and a multiple of 16, packed 16-byte aligned");
  uint8_t* tmp = nullptr;
  STRIVE_CUDA(cudaMallocAsync((void**)&tmp, (size_t)n * 65536, stream));
  dim3 gp(16, n);
  KPROF("crop_pack", stream, STRIVE_CUDA_LAUNCH(crop_pack_kernel, gp, 256, 0, stream, *map, pose, map_of, tmp, n));
  STRIVE_LAUNCH_CHECK();
  const size_t tot = (size_t)n * 65536;
  crop_unpack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(tmp, out, n, map->C);
  STRIVE_LAUNCH_CHECK();
  STRIVE_CUDA(cudaFreeAsync(tmp, stream));
  return 0;
}

static BiasArg make_bias(const float* h_bias, int nb) {
  BiasArg b;
  for (int i = 0; i < 64; i++) b.b[i] = i < nb ? h_bias[i] : 0.f;
  return b;
}

int tc_launch_conv1(const StriveMap* map, const float* pose, const int32_t* map_of, const uint8_t* wpack, const float* h_bias, float* out,
                    double* out_stats, uint8_t* packed_crop, int n, cudaStream_t stream) {
  static unsigned attr = 0;
  const size_t smem = T1_WBYTES + T1_NBUF * T1_PATCH_BYTES;
  if (strive_first_use_on_device(&attr)) {
    STRIVE_CUDA(cudaFuncSetAttribute(tc_conv1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  STRIVE_CHECK(map->packed != nullptr, STRIVE_EINVAL, "tensor-core map encoder needs StriveMap.packed");
  STRIVE_CHECK(map->packed_pitch >= map->W && (map->packed_pitch & 15) == 0 && ((uintptr_t)map->packed & 15) == 0, STRIVE_EINVAL, "StriveMap.packed_pitch must be >= W and a multiple of 16, packed 16-byte aligned");
  dim3 gp(16, n);
  KPROF("crop_pack", stream, STRIVE_CUDA_LAUNCH(crop_pack_kernel, gp, 256, 0, stream, *map, pose, map_of, packed_crop, n));
  STRIVE_LAUNCH_CHECK();
  const int items = n * T1_RB * T1_CB;
  const int grid = items < num_sms() ? items : num_sms();     // the kernel owns all 512 TMEM columns: one CTA per SM
  const BiasArg bias = make_bias(h_bias, 16);
  KPROF("tc_conv1", stream, STRIVE_CUDA_LAUNCH(tc_conv1_kernel, grid, TC_THREADS, smem, stream, packed_crop, wpack, bias, out, out_stats, n));
  STRIVE_LAUNCH_CHECK();
  return 0;
}

template <int CIN, int KS, int HIN, int HOUT, int COUT, int NCH, int NBUF, bool OUT_BLK>
static int tc_launch(const char* name, const float* in, const double* in_stats, const float* gam, const float* bet, const uint8_t* wpack,
                     const float* h_bias, float* out, double* out_stats, int n, cudaStream_t stream) {
  using Cfg = TcCfg<CIN, KS, HIN, HOUT, COUT, NCH, NBUF>;
  static_assert(COUT % NCH == 0 && (NCH == 32 || NCH == 64) && CIN % 16 == 0, "tc conv tiling");
  static_assert(Cfg::SMEM <= 225 * 1024, "tc conv shared memory");
  auto kern = tc_conv_kernel<CIN, KS, HIN, HOUT, COUT, NCH, NBUF, OUT_BLK>;
  static unsigned attr = 0;
  if (strive_first_use_on_device(&attr)) {
    STRIVE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  }
  const int items = n * Cfg::TILES;
  int gx = num_sms() / (COUT / NCH);
  if (gx < 1) gx = 1;
  if (gx > items) gx = items;
  dim3 grid(gx, COUT / NCH);
  const BiasArg bias = make_bias(h_bias, COUT);
  KPROF(name, stream, STRIVE_CUDA_LAUNCH(kern, grid, T2_THREADS, Cfg::SMEM, stream, in, in_stats, gam, bet, wpack, bias, out, out_stats, n, g_tc_dbg));
  STRIVE_LAUNCH_CHECK();
  return 0;
}

// conv3 kernel choice (strive_mapenc_set_pair): 1 = CTA-pair kernel (tcgen05 cta_group::2, default), 0 = single-CTA kernel.
// (conv2 was built the same way -- Z_r of 32 rows, four operand buffers, four accumulators, outputs within 1e-5 of the single-CTA
// kernel -- and measured 825 us against 765 us per 2048 crops: with the B bytes halved its MMAs need 1880 cycles per tile, but the
// producers + the epilogue of that layer need ~3100 whatever the MMA side does.  Removed.)
static int g_pair_mask = 1;
extern "C" int strive_mapenc_set_pair(int on) {
  g_pair_mask = on == 2 ? 3 : (on ? 1 : 0);     // 2 = pairs even when fewer than ~74 of them fit at once (tests under tools that limit residency)
  return 0;
}
// how many CTA pairs of `kern` can be resident at once (a GPC with an odd number of usable SMs leaves one of them without a partner)
template <typename K>
static int max_resident_pairs(K kern, size_t smem) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * num_sms());
  cfg.blockDim = dim3(T2_THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int nc = 0;
  if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess) {
    (void)cudaGetLastError();
    nc = 0;
  }
  return nc;
}

int tc_launch_conv2(const float* in, const double* in_stats, const float* gam, const float* bet, const uint8_t* wpack, const float* h_bias,
                    float* out, double* out_stats, int n, cudaStream_t stream) {
  return tc_launch<16, 5, 125, 61, 32, 32, 3, true>("tc_conv2", in, in_stats, gam, bet, wpack, h_bias, out, out_stats, n, stream);
}
int tc_launch_conv3(const float* in, const double* in_stats, const float* gam, const float* bet, const uint8_t* wpack, const uint8_t* wpack_pair,
                    const float* h_bias, float* out, double* out_stats, int n, cudaStream_t stream) {
  using Cfg = TcCfg<32, 5, 61, 29, 64, 64, T3_NBUF>;
  constexpr size_t SMEM = (size_t)T3_WCHUNK + (size_t)T3_NBUF * Cfg::A_BYTES;
  constexpr size_t SMEM_PAIR = (size_t)T3P_WBYTES + (size_t)T3P_NBUF * Cfg::A_BYTES;
  static_assert(SMEM <= 225 * 1024 && SMEM_PAIR <= 225 * 1024, "conv3 shared memory");
  static unsigned attr = 0;
  static int max_clusters_dev[32] = {};
  int dev = 0;
  if (strive_first_use_on_device(&attr, &dev)) {
    int max_clusters = 0;
    STRIVE_CUDA(cudaFuncSetAttribute(tc_conv3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    STRIVE_CUDA(cudaFuncSetAttribute(tc_conv3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_PAIR));
    max_clusters = max_resident_pairs(tc_conv3_pair_kernel, SMEM_PAIR);
    if (getenv("STRIVE_TC_VERBOSE")) fprintf(stderr, "strive_b200: conv3 pair kernel: %d CTA pairs resident on %d SMs\n", max_clusters, num_sms());
    max_clusters_dev[dev] = max_clusters;
  }
  const int max_clusters = max_clusters_dev[dev];
  const BiasArg bias = make_bias(h_bias, 64);
  if ((g_pair_mask & 1) && wpack_pair != nullptr && (2 * max_clusters >= num_sms() - 8 || ((g_pair_mask & 2) && max_clusters >= 1))) {
    int ncl = max_clusters < n ? max_clusters : n;       // one crop is the unit of work of a pair
    KPROF("tc_conv3", stream, STRIVE_CUDA_LAUNCH(tc_conv3_pair_kernel, 2 * ncl, T2_THREADS, SMEM_PAIR, stream, in, in_stats, gam, bet, wpack_pair, bias, out, out_stats, n));
    STRIVE_LAUNCH_CHECK();
    return 0;
  }
  const int items = n * Cfg::TILES;
  int gx = num_sms();
  if (gx > (items + 1) / 2) gx = (items + 1) / 2;
  KPROF("tc_conv3", stream, STRIVE_CUDA_LAUNCH(tc_conv3_kernel, gx, T2_THREADS, SMEM, stream, in, in_stats, gam, bet, wpack, bias, out, out_stats, n));
  STRIVE_LAUNCH_CHECK();
  return 0;
}
int tc_launch_conv4(const float* in, const double* in_stats, const float* gam, const float* bet, const uint8_t* wpack, const float* h_bias,
                    float* out, double* out_stats, int n, cudaStream_t stream) {
  return tc_launch<64, 3, 29, 14, 64, 64, 2, false>("tc_conv4", in, in_stats, gam, bet, wpack, h_bias, out, out_stats, n, stream);
}

template <int CIN, int KS, int HIN, int HOUT, int COUT, bool FINAL>
static int tc3_launch(const char* name, const float* in, const double* in_stats, const float* gam, const float* bet, const uint8_t* wpack,
                      const float* bias, float* out, double* out_stats, int n, cudaStream_t stream) {
  using Cfg = Tc3Cfg<CIN, KS, HIN, HOUT, COUT, FINAL>;
  static_assert(Cfg::SMEM <= 226 * 1024, "tc gemm shared memory");
  auto kern = tc_gemm_kernel<CIN, KS, HIN, HOUT, COUT, FINAL>;
  static unsigned attr = 0;
  if (strive_first_use_on_device(&attr)) {
    STRIVE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  }
  const long long rows = (long long)n * Cfg::PIX;
  const int tiles = (int)((rows + 127) / 128);
  const int gx = tiles < num_sms() ? tiles : num_sms();
  KPROF(name, stream, STRIVE_CUDA_LAUNCH(kern, gx, TC_THREADS, Cfg::SMEM, stream, in, in_stats, gam, bet, wpack, bias, out, out_stats, n, g_tc_dbg));
  STRIVE_LAUNCH_CHECK();
  return 0;
}

int tc_launch_conv5(const float* in, const double* in_stats, const float* gam, const float* bet, const uint8_t* wpack, const float* bias,
                    float* out, double* out_stats, int n, cudaStream_t stream) {
  return tc3_launch<64, 3, 14, 6, 128, false>("tc_conv5", in, in_stats, gam, bet, wpack, bias, out, out_stats, n, stream);
}
int tc_launch_conv6(const float* in, const double* in_stats, const float* gam, const float* bet, const uint8_t* wpack, const float* bias,
                    float* out, double* out_stats, int n, cudaStream_t stream) {
  return tc3_launch<128, 3, 6, 2, 128, false>("tc_conv6", in, in_stats, gam, bet, wpack, bias, out, out_stats, n, stream);
}
int tc_launch_fc(const float* in, const double* in_stats, const float* gam, const float* bet, const uint8_t* wpack, const float* bias,
                 float* out, int n, cudaStream_t stream) {
  return tc3_launch<128, 2, 2, 1, 64, true>("tc_fc", in, in_stats, gam, bet, wpack, bias, out, nullptr, n, stream);
}
