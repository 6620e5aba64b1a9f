// Tensor-core (tcgen05 + TMEM) implicit-GEMM convolutions of the map encoder.
//
// Numerics: bf16 operand SPLITTING with fp32 accumulation in TMEM.  x = hi + lo (hi = bf16(x), lo = bf16(x - hi)).
//   conv1: the input is the binary crop (exact in bf16); weights are split -> 2 MMAs per K step (error 2^-17 relative).
//   conv2..: activations and weights are both split -> hi*hi + lo*hi + hi*lo (3 MMAs, dropped lo*lo term 2^-18).
// This keeps the encoder at fp32-level accuracy (tests: <= 1e-4 abs on O(1) features; measured ~1e-5) at 1/3 of the
// dense bf16 tensor peak; see DESIGN.md.
//
// Operand addressing ("shifted window"): the GroupNorm'ed/ReLU'ed input tile is written to shared memory ONCE, with the
// columns de-interleaved by parity (stride-2 convolution -> consecutive output pixels are consecutive 16-byte rows of a
// parity plane).  Every filter tap is then just a different start address / the same LBO,SBO in the K-major no-swizzle
// matrix descriptor, so there is no im2col expansion in shared memory at all.
#include "common.cuh"
#include "tc.cuh"

// ======================================================================================================
// conv1: 4 -> 16, k7 s2, input gathered from the raster.  CTA = 32x32 outputs = 8 MMA sub-tiles (16 rows x 8 cols).
// A tile: [row][col][4 ch] bf16 (8 B / pixel); one K=16 MMA = 4 taps (kx..kx+3) x 4 channels at fixed ky.
// ======================================================================================================
#define T1_PH 69
#define T1_PW 70
#define T1_WBYTES (7 * 2 * 2 * 512)
#define T1_PATCH_BYTES (T1_PH * T1_PW * 8)
#define T1_THREADS 256
#define T1_SUPER 4   // 4 x 4 super-tiles of 32 x 32 outputs cover 125 x 125

struct CropFrameTc {
  float px, py, hc, hs;
  double dx0, dx1, inv0, inv1;
  const uint8_t* base;
};

// round-half-even(g / dx) exactly as torch.round(float64 quotient): multiply by the reciprocal and fall back to the true
// division only when the product lands within 1e-6 of a .5 boundary (the only case where the two could round differently).
__device__ __forceinline__ long long round_div_exact(float g, double dx, double inv) {
  const double q = (double)g * inv;
  const double fr = q - floor(q);
  if (fabs(fr - 0.5) < 1e-6) return __double2ll_rn((double)g / dx);
  return __double2ll_rn(q);
}

__device__ __forceinline__ void crop_pixel_tc(const CropFrameTc& f, float l, float w, int H, int W, long long& xp, long long& yp) {
  // exact restatement of get_map_obs (reference datasets/nuscenes_utils.py:248-263); see mapenc.cu crop_pixel
  float gx = __fadd_rn(__fsub_rn(__fmul_rn(l, f.hc), __fmul_rn(w, f.hs)), f.px);
  float gy = __fadd_rn(__fadd_rn(__fmul_rn(l, f.hs), __fmul_rn(w, f.hc)), f.py);
  if (isnan(gx)) gx = 0.f;
  if (isnan(gy)) gy = 0.f;
  xp = round_div_exact(gx, f.dx0, f.inv0);
  yp = round_div_exact(gy, f.dx1, f.inv1);
  if (yp < 0 || yp >= H || xp < 0 || xp >= W) { xp = 0; yp = 0; }
}

__global__ void __launch_bounds__(T1_THREADS) tc_conv1_kernel(StriveMap map, const float* __restrict__ pose,
                                                              const int32_t* __restrict__ map_of, const uint8_t* __restrict__ wpack,
                                                              const float* __restrict__ bias, float* __restrict__ out,
                                                              double* __restrict__ out_stats, int n) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sP = smem + T1_WBYTES;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  __shared__ float s_bias[16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < T1_WBYTES / 16; i += T1_THREADS) reinterpret_cast<int4*>(sW)[i] = __ldg(reinterpret_cast<const int4*>(wpack) + i);
  if (tid < 16) s_bias[tid] = bias[tid];
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base, 128);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;
  const uint32_t idesc = tc::idesc_bf16_f32(128, 16);
  uint32_t phase = 0;
  const int items = n * T1_SUPER * T1_SUPER;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int crop = item / (T1_SUPER * T1_SUPER);
    const int st = item % (T1_SUPER * T1_SUPER);
    const int oy0 = (st / T1_SUPER) * 32, ox0 = (st % T1_SUPER) * 32;
    {
      const int m = map_of[crop];
      CropFrameTc f;
      f.px = pose[crop * 4 + 0]; f.py = pose[crop * 4 + 1]; f.hc = pose[crop * 4 + 2]; f.hs = pose[crop * 4 + 3];
      f.dx0 = map.dx[m * 2 + 0]; f.dx1 = map.dx[m * 2 + 1];
      f.inv0 = 1.0 / f.dx0; f.inv1 = 1.0 / f.dx1;
      f.base = map.packed + (size_t)m * map.H * map.W;     // bit c of a byte = layer c (binary raster)
      for (int i = tid; i < T1_PH * T1_PW; i += T1_THREADS) {
        const int r = i / T1_PW, c = i % T1_PW;
        const int iy = oy0 * 2 + r, ix = ox0 * 2 + c;
        uint32_t lo = 0, hi = 0;
        if (iy < 256 && ix < 256) {
          long long xp, yp;
          crop_pixel_tc(f, __ldg(map.lin_l + iy), __ldg(map.lin_w + ix), map.H, map.W, xp, yp);
          const uint32_t bits = __ldg(f.base + (size_t)yp * map.W + xp);
          const uint32_t one = 0x3F80u;   // bf16(1.0)
          lo = ((bits & 1u) ? one : 0u) | ((bits & 2u) ? (one << 16) : 0u);
          hi = ((bits & 4u) ? one : 0u) | ((bits & 8u) ? (one << 16) : 0u);
        }
        *reinterpret_cast<uint2*>(sP + (size_t)i * 8) = make_uint2(lo, hi);
      }
    }
    tc::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      const uint32_t pbase = tc::smem_u32(sP), wbase = tc::smem_u32(sW);
#pragma unroll 1
      for (int sub = 0; sub < 8; sub++) {
        const int sy = sub >> 2, sx = sub & 3;
        const uint32_t abase = pbase + ((sy * 32) * T1_PW + sx * 16) * 8;
        uint32_t acc = 0;
#pragma unroll 1
        for (int ky = 0; ky < 7; ky++) {
#pragma unroll
          for (int kq = 0; kq < 2; kq++) {
            const uint64_t ad = tc::smem_desc(abase + (ky * T1_PW + 4 * kq) * 8, 16, 2 * T1_PW * 8);
            const uint32_t wb = wbase + ((ky * 2 + kq) * 2) * 512;
            tc::mma_bf16(tm + sub * 16, ad, tc::smem_desc(wb, 256, 128), idesc, acc);
            tc::mma_bf16(tm + sub * 16, ad, tc::smem_desc(wb + 512, 256, 128), idesc, 1);
            acc = 1;
          }
        }
      }
      tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, phase);
    phase ^= 1;
    tc::tc_fence_after();
    // epilogue: warp w handles TMEM lanes (w%4)*32.. of sub-tiles (w/4)*4 .. +3
    float s1 = 0.f, s2 = 0.f;
    {
      const int q = warp & 3;
      const int m = q * 32 + lane;              // row of the 128-row sub-tile = oy_l*8 + ox_l
      const int oyl = m >> 3, oxl = m & 7;
#pragma unroll 1
      for (int k = 0; k < 4; k++) {
        const int sub = (warp >> 2) * 4 + k;
        const int sy = sub >> 2, sx = sub & 3;
        float v[16];
        tc::tmem_ld16(tm + ((uint32_t)(q * 32) << 16) + sub * 16, v);
        const int oy = oy0 + sy * 16 + oyl, ox = ox0 + sx * 8 + oxl;
        if (oy < 125 && ox < 125) {
          float* o = out + (((size_t)crop * 125 + oy) * 125 + ox) * 16;
#pragma unroll
          for (int c = 0; c < 16; c++) {
            v[c] += s_bias[c];
            s1 += v[c];
            s2 = fmaf(v[c], v[c], s2);
          }
#pragma unroll
          for (int c = 0; c < 16; c += 4) *reinterpret_cast<float4*>(o + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        }
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      atomicAdd(out_stats + (size_t)crop * 2, (double)s1);
      atomicAdd(out_stats + (size_t)crop * 2 + 1, (double)s2);
    }
    tc::tc_fence_before();
    __syncthreads();   // TMEM + patch are free again
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, 128);
}

// ======================================================================================================
// conv2..4: stride-2 kxk conv, CIN multiple of 16, output-channel chunks of N=32 (blockIdx.y), NHWC fp32 in/out.
// CTA tile = 16 x 8 outputs (M = 128).  Weights of the chunk stay resident in shared memory; the input tile is staged
// per 16-channel chunk into an NBUF-deep ring so staging of chunk i+1 overlaps the (asynchronous) MMAs of chunk i.
// ======================================================================================================
template <int CIN, int KS, int HIN, int HOUT, int COUT, int NBUF>
struct TcCfg {
  static constexpr int N = 32;
  static constexpr int C2 = CIN / 16;
  static constexpr int TAPS = KS * KS;
  static constexpr int PH = 30 + KS;
  static constexpr int PW = 14 + KS;
  static constexpr int PQ = 8 + (KS - 1) / 2;
  static constexpr int A_PREC_BYTES = 2 * PH * 2 * PQ * 16;
  static constexpr int A_BYTES = 2 * A_PREC_BYTES;
  static constexpr int W_BYTES = C2 * TAPS * 2 * 1024;
  static constexpr int TILES_Y = (HOUT + 15) / 16;
  static constexpr int TILES_X = (HOUT + 7) / 8;
  static constexpr int TILES = TILES_Y * TILES_X;
  static constexpr size_t SMEM = (size_t)W_BYTES + (size_t)NBUF * A_BYTES;
};

#define T2_THREADS 512
template <int CIN, int KS, int HIN, int HOUT, int COUT, int NBUF>
__global__ void __launch_bounds__(T2_THREADS) tc_conv_kernel(const float* __restrict__ in, const double* __restrict__ in_stats,
                                                      const float* __restrict__ gam, const float* __restrict__ bet,
                                                      const uint8_t* __restrict__ wpack, const float* __restrict__ bias,
                                                      float* __restrict__ out, double* __restrict__ out_stats, int n) {
  using Cfg = TcCfg<CIN, KS, HIN, HOUT, COUT, NBUF>;
  constexpr int PH = Cfg::PH, PW = Cfg::PW, PQ = Cfg::PQ, C2 = Cfg::C2, TAPS = Cfg::TAPS;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sA = smem + Cfg::W_BYTES;
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base;
  __shared__ float s_gam[CIN], s_bet[CIN], s_bias[32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunk = blockIdx.y;
  {
    const int4* src = reinterpret_cast<const int4*>(wpack + (size_t)nchunk * Cfg::W_BYTES);
    for (int i = tid; i < Cfg::W_BYTES / 16; i += T2_THREADS) reinterpret_cast<int4*>(sW)[i] = __ldg(src + i);
  }
  for (int i = tid; i < CIN; i += T2_THREADS) { s_gam[i] = gam[i]; s_bet[i] = bet[i]; }
  if (tid < 32) s_bias[tid] = bias[nchunk * 32 + tid];
  if (tid == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init(&bars[1], 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base, 32);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;
  const uint32_t idesc = tc::idesc_bf16_f32(128, 32);
  int counter = 0;
  bool pend[2] = {false, false};
  uint32_t ph[2] = {0, 0};
  bool have_prev = false;
  int p_crop = 0, p_ty0 = 0, p_tx0 = 0, p_b = 0;

  auto wait_buf = [&](int b) {
    if (pend[b]) {
      tc::mbar_wait(&bars[b], ph[b]);
      ph[b] ^= 1;
      pend[b] = false;
    }
  };
  auto epilogue = [&](int crop, int ty0, int tx0) {
    if (warp >= 4) return;   // TMEM lanes 0..127 are read by warps 0..3
    const int m = warp * 32 + lane;
    const int oy = ty0 + (m >> 3), ox = tx0 + (m & 7);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      float v[16];
      tc::tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + h * 16, v);
      if (oy < HOUT && ox < HOUT) {
        float* o = out + (((size_t)crop * HOUT + oy) * HOUT + ox) * COUT + nchunk * 32 + h * 16;
#pragma unroll
        for (int c = 0; c < 16; c++) {
          v[c] += s_bias[h * 16 + c];
          s1 += v[c];
          s2 = fmaf(v[c], v[c], s2);
        }
#pragma unroll
        for (int c = 0; c < 16; c += 4) *reinterpret_cast<float4*>(o + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      atomicAdd(out_stats + (size_t)crop * 2, (double)s1);
      atomicAdd(out_stats + (size_t)crop * 2 + 1, (double)s2);
    }
  };

  const int items = n * Cfg::TILES;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int crop = item / Cfg::TILES, tile = item % Cfg::TILES;
    const int ty0 = (tile / Cfg::TILES_X) * 16, tx0 = (tile % Cfg::TILES_X) * 8;
    float mean, rstd;
    {
      const double cnt = (double)CIN * HIN * HIN;
      const double mu = in_stats[(size_t)crop * 2] / cnt;
      double var = in_stats[(size_t)crop * 2 + 1] / cnt - mu * mu;
      if (var < 0.0) var = 0.0;
      mean = (float)mu;
      rstd = (float)(1.0 / sqrt(var + 1e-5));
    }
#pragma unroll 1
    for (int c2 = 0; c2 < C2; c2++) {
      const int b = counter % NBUF;
      wait_buf(b);
      uint8_t* dst = sA + (size_t)b * Cfg::A_BYTES;
      const float* src = in + (size_t)crop * HIN * HIN * CIN + c2 * 16;
      for (int p = tid; p < PH * PW; p += T2_THREADS) {
        const int row = p / PW, col = p % PW;
        const int iy = 2 * ty0 + row, ix = 2 * tx0 + col;
        uint32_t hi[8], lo[8];
        if (iy < HIN && ix < HIN) {
          const float4* s4 = reinterpret_cast<const float4*>(src + ((size_t)iy * HIN + ix) * CIN);
          float x[16];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const float4 t = __ldg(s4 + q);
            x[q * 4] = t.x; x[q * 4 + 1] = t.y; x[q * 4 + 2] = t.z; x[q * 4 + 3] = t.w;
          }
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            float h0, l0, h1, l1;
            const float y0 = fmaxf(fmaf((x[c] - mean) * rstd, s_gam[c2 * 16 + c], s_bet[c2 * 16 + c]), 0.f);
            const float y1 = fmaxf(fmaf((x[c + 1] - mean) * rstd, s_gam[c2 * 16 + c + 1], s_bet[c2 * 16 + c + 1]), 0.f);
            tc::split_bf16(y0, h0, l0);
            tc::split_bf16(y1, h1, l1);
            hi[c >> 1] = tc::pack_bf16(h0, h1);
            lo[c >> 1] = tc::pack_bf16(l0, l1);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 8; c++) { hi[c] = 0u; lo[c] = 0u; }
        }
        const int u0 = ((row * 2 + (col & 1)) * PQ + (col >> 1));   // 16-byte unit inside channel-group 0
        uint8_t* d0 = dst + (size_t)u0 * 16;
        constexpr int CG = PH * 2 * PQ * 16;
        *reinterpret_cast<uint4*>(d0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(d0 + CG) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        *reinterpret_cast<uint4*>(d0 + Cfg::A_PREC_BYTES) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(d0 + Cfg::A_PREC_BYTES + CG) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      }
      tc::fence_async_smem();
      if (c2 == 0 && have_prev) {
        wait_buf(p_b);
        tc::tc_fence_after();
        epilogue(p_crop, p_ty0, p_tx0);
        tc::tc_fence_before();
      }
      __syncthreads();
      if (tid == 0) {
        tc::tc_fence_after();
        const uint32_t abase = tc::smem_u32(dst);
        const uint32_t wbase = tc::smem_u32(sW) + c2 * TAPS * 2048;
        constexpr uint32_t LBO_A = PH * 2 * PQ * 16, SBO_A = 64 * PQ;
        uint32_t acc = (c2 > 0) ? 1u : 0u;
#pragma unroll 1
        for (int tap = 0; tap < TAPS; tap++) {
          const int ky = tap / KS, kx = tap % KS;
          const uint32_t aoff = ((ky * 2 + (kx & 1)) * PQ + (kx >> 1)) * 16;
          const uint64_t ah = tc::smem_desc(abase + aoff, LBO_A, SBO_A);
          const uint64_t al = tc::smem_desc(abase + Cfg::A_PREC_BYTES + aoff, LBO_A, SBO_A);
          const uint64_t bh = tc::smem_desc(wbase + tap * 2048, 512, 128);
          const uint64_t bl = tc::smem_desc(wbase + tap * 2048 + 1024, 512, 128);
          tc::mma_bf16(tm, ah, bh, idesc, acc);
          tc::mma_bf16(tm, al, bh, idesc, 1);
          tc::mma_bf16(tm, ah, bl, idesc, 1);
          acc = 1;
        }
        tc::mma_commit(&bars[b]);
      }
      pend[b] = true;
      counter++;
    }
    have_prev = true;
    p_crop = crop; p_ty0 = ty0; p_tx0 = tx0;
    p_b = (counter - 1) % NBUF;
  }
  if (have_prev) {
    wait_buf(p_b);
    tc::tc_fence_after();
    epilogue(p_crop, p_ty0, p_tx0);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, 32);
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
  static constexpr size_t SMEM = 2 * (size_t)STAGE;
};

template <int CIN, int KS, int HIN, int HOUT, int COUT, bool FINAL>
__global__ void __launch_bounds__(T2_THREADS) tc_gemm_kernel(const float* __restrict__ in, const double* __restrict__ in_stats,
                                                             const float* __restrict__ gam, const float* __restrict__ bet,
                                                             const uint8_t* __restrict__ wpack, const float* __restrict__ bias,
                                                             float* __restrict__ out, double* __restrict__ out_stats, int n) {
  using Cfg = Tc3Cfg<CIN, KS, HIN, HOUT, COUT, FINAL>;
  constexpr int PIX = Cfg::PIX, NCH = Cfg::NCH;
  static_assert(CIN % 64 == 0 && COUT % 16 == 0, "tc_gemm tiling");
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base;
  __shared__ float s_gam[CIN], s_bet[CIN], s_bias[COUT];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < CIN; i += T2_THREADS) { s_gam[i] = gam[i]; s_bet[i] = bet[i]; }
  for (int i = tid; i < COUT; i += T2_THREADS) s_bias[i] = bias[i];
  if (tid == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init(&bars[1], 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base, COUT < 32 ? 32 : COUT);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = tmem_base;
  const uint32_t idesc = tc::idesc_bf16_f32(128, COUT);
  int counter = 0;
  bool pend[2] = {false, false};
  uint32_t ph[2] = {0, 0};
  bool have_prev = false;
  int p_tile = 0, p_b = 0;
  auto wait_buf = [&](int b) {
    if (pend[b]) {
      tc::mbar_wait(&bars[b], ph[b]);
      ph[b] ^= 1;
      pend[b] = false;
    }
  };
  const long long rows_total = (long long)n * PIX;
  auto epilogue = [&](int tile) {
    if (warp >= 4) return;
    const long long gm = (long long)tile * 128 + warp * 32 + lane;
    const bool valid = gm < rows_total;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
    for (int h = 0; h < COUT / 16; h++) {
      float v[16];
      tc::tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + h * 16, v);
      if (valid) {
        float* o = out + (size_t)gm * COUT + h * 16;
#pragma unroll
        for (int c = 0; c < 16; c++) {
          v[c] += s_bias[h * 16 + c];
          s1 += v[c];
          s2 = fmaf(v[c], v[c], s2);
        }
#pragma unroll
        for (int c = 0; c < 16; c += 4) *reinterpret_cast<float4*>(o + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
      }
    }
    if (!FINAL && valid) {
      const int crop = (int)(gm / PIX);
      atomicAdd(out_stats + (size_t)crop * 2, (double)s1);
      atomicAdd(out_stats + (size_t)crop * 2 + 1, (double)s2);
    }
  };

  const int tiles = (int)((rows_total + 127) / 128);
  const int m = tid & 127;          // every thread stages a fixed row of the tile, k-groups (tid>>7) and (tid>>7)+4
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long gm = (long long)tile * 128 + m;
    const bool valid = gm < rows_total;
    int crop = 0, oy = 0, ox = 0;
    float mean = 0.f, rstd = 0.f;
    if (valid) {
      crop = (int)(gm / PIX);
      const int pix = (int)(gm % PIX);
      oy = pix / HOUT; ox = pix % HOUT;
      const double cnt = (double)CIN * HIN * HIN;
      const double mu = in_stats[(size_t)crop * 2] / cnt;
      double var = in_stats[(size_t)crop * 2 + 1] / cnt - mu * mu;
      if (var < 0.0) var = 0.0;
      mean = (float)mu;
      rstd = (float)(1.0 / sqrt(var + 1e-5));
    }
#pragma unroll 1
    for (int kc = 0; kc < NCH; kc++) {
      const int b = counter & 1;
      wait_buf(b);
      uint8_t* sA = smem + (size_t)b * Cfg::STAGE;
      uint8_t* sWt = sA + 2 * Cfg::A_PREC;
      {
        const int4* src = reinterpret_cast<const int4*>(wpack + (size_t)kc * 2 * Cfg::W_PREC);
        for (int i = tid; i < 2 * Cfg::W_PREC / 16; i += T2_THREADS) reinterpret_cast<int4*>(sWt)[i] = __ldg(src + i);
      }
      const int k0 = kc * 64;
      const int tap = k0 / CIN, c0 = k0 % CIN;
      const int ky = tap / KS, kx = tap % KS;
      const float* src = in + (((size_t)crop * HIN + 2 * oy + ky) * HIN + 2 * ox + kx) * CIN + c0;
#pragma unroll
      for (int half = 0; half < 2; half++) {
        const int kg = (tid >> 7) + half * 4;
        uint32_t hi[4] = {0u, 0u, 0u, 0u}, lo[4] = {0u, 0u, 0u, 0u};
        if (valid) {
          const float4 t0 = __ldg(reinterpret_cast<const float4*>(src + kg * 8));
          const float4 t1 = __ldg(reinterpret_cast<const float4*>(src + kg * 8 + 4));
          const float x[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
          for (int c = 0; c < 8; c += 2) {
            const int ch = c0 + kg * 8 + c;
            float h0, l0, h1, l1;
            const float y0 = fmaxf(fmaf((x[c] - mean) * rstd, s_gam[ch], s_bet[ch]), 0.f);
            const float y1 = fmaxf(fmaf((x[c + 1] - mean) * rstd, s_gam[ch + 1], s_bet[ch + 1]), 0.f);
            tc::split_bf16(y0, h0, l0);
            tc::split_bf16(y1, h1, l1);
            hi[c >> 1] = tc::pack_bf16(h0, h1);
            lo[c >> 1] = tc::pack_bf16(l0, l1);
          }
        }
        const int unit = (kg * 16 + (m >> 3)) * 8 + (m & 7);
        *reinterpret_cast<uint4*>(sA + (size_t)unit * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(sA + Cfg::A_PREC + (size_t)unit * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      tc::fence_async_smem();
      if (kc == 0 && have_prev) {
        wait_buf(p_b);
        tc::tc_fence_after();
        epilogue(p_tile);
        tc::tc_fence_before();
      }
      __syncthreads();
      if (tid == 0) {
        tc::tc_fence_after();
        const uint32_t abase = tc::smem_u32(sA), wbase = tc::smem_u32(sWt);
        constexpr uint32_t LBO_W = (COUT / 8) * 128;
        uint32_t acc = (kc > 0) ? 1u : 0u;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const uint64_t ah = tc::smem_desc(abase + j * 2 * 2048, 2048, 128);
          const uint64_t al = tc::smem_desc(abase + Cfg::A_PREC + j * 2 * 2048, 2048, 128);
          const uint64_t bh = tc::smem_desc(wbase + j * 2 * LBO_W, LBO_W, 128);
          const uint64_t bl = tc::smem_desc(wbase + Cfg::W_PREC + j * 2 * LBO_W, LBO_W, 128);
          tc::mma_bf16(tm, ah, bh, idesc, acc);
          tc::mma_bf16(tm, al, bh, idesc, 1);
          tc::mma_bf16(tm, ah, bl, idesc, 1);
          acc = 1;
        }
        tc::mma_commit(&bars[b]);
      }
      pend[b] = true;
      counter++;
    }
    have_prev = true;
    p_tile = tile;
    p_b = (counter - 1) & 1;
  }
  if (have_prev) {
    wait_buf(p_b);
    tc::tc_fence_after();
    epilogue(p_tile);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tm, COUT < 32 ? 32 : COUT);
}

// ------------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------------
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

int tc_launch_conv1(const StriveMap* map, const float* pose, const int32_t* map_of, const uint8_t* wpack, const float* bias, float* out,
                    double* out_stats, int n, cudaStream_t stream) {
  static bool attr = false;
  const size_t smem = T1_WBYTES + T1_PATCH_BYTES;
  if (!attr) {
    STRIVE_CUDA(cudaFuncSetAttribute(tc_conv1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int items = n * T1_SUPER * T1_SUPER;
  const int grid = items < num_sms() * 3 ? items : num_sms() * 3;
  KPROF("tc_conv1", stream, tc_conv1_kernel<<<grid, T1_THREADS, smem, stream>>>(*map, pose, map_of, wpack, bias, out, out_stats, n));
  STRIVE_LAUNCH_CHECK();
  return 0;
}

template <int CIN, int KS, int HIN, int HOUT, int COUT, int NBUF>
static int tc_launch(const char* name, const float* in, const double* in_stats, const float* gam, const float* bet, const uint8_t* wpack,
                     const float* bias, float* out, double* out_stats, int n, cudaStream_t stream) {
  using Cfg = TcCfg<CIN, KS, HIN, HOUT, COUT, NBUF>;
  static_assert(COUT % 32 == 0 && CIN % 16 == 0, "tc conv tiling");
  static_assert(Cfg::SMEM <= 227 * 1024, "tc conv shared memory");
  auto kern = tc_conv_kernel<CIN, KS, HIN, HOUT, COUT, NBUF>;
  static bool attr = false;
  if (!attr) {
    STRIVE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    attr = true;
  }
  const int items = n * Cfg::TILES;
  const int per_sm = (Cfg::SMEM + 2048) * 2 <= 227 * 1024 ? 2 : 1;
  int gx = num_sms() * per_sm / (COUT / 32);
  if (gx < 1) gx = 1;
  if (gx > items) gx = items;
  dim3 grid(gx, COUT / 32);
  KPROF(name, stream, kern<<<grid, T2_THREADS, Cfg::SMEM, stream>>>(in, in_stats, gam, bet, wpack, bias, out, out_stats, n));
  STRIVE_LAUNCH_CHECK();
  return 0;
}

int tc_launch_conv2(const float* in, const double* in_stats, const float* gam, const float* bet, const uint8_t* wpack, const float* bias,
                    float* out, double* out_stats, int n, cudaStream_t stream) {
  return tc_launch<16, 5, 125, 61, 32, 2>("tc_conv2", in, in_stats, gam, bet, wpack, bias, out, out_stats, n, stream);
}
int tc_launch_conv3(const float* in, const double* in_stats, const float* gam, const float* bet, const uint8_t* wpack, const float* bias,
                    float* out, double* out_stats, int n, cudaStream_t stream) {
  return tc_launch<32, 5, 61, 29, 64, 2>("tc_conv3", in, in_stats, gam, bet, wpack, bias, out, out_stats, n, stream);
}
int tc_launch_conv4(const float* in, const double* in_stats, const float* gam, const float* bet, const uint8_t* wpack, const float* bias,
                    float* out, double* out_stats, int n, cudaStream_t stream) {
  return tc_launch<64, 3, 29, 14, 64, 2>("tc_conv4", in, in_stats, gam, bet, wpack, bias, out, out_stats, n, stream);
}

template <int CIN, int KS, int HIN, int HOUT, int COUT, bool FINAL>
static int tc3_launch(const char* name, const float* in, const double* in_stats, const float* gam, const float* bet, const uint8_t* wpack,
                      const float* bias, float* out, double* out_stats, int n, cudaStream_t stream) {
  using Cfg = Tc3Cfg<CIN, KS, HIN, HOUT, COUT, FINAL>;
  static_assert(Cfg::SMEM <= 227 * 1024, "tc gemm shared memory");
  auto kern = tc_gemm_kernel<CIN, KS, HIN, HOUT, COUT, FINAL>;
  static bool attr = false;
  if (!attr) {
    STRIVE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    attr = true;
  }
  const long long rows = (long long)n * Cfg::PIX;
  const int tiles = (int)((rows + 127) / 128);
  const int gx = tiles < num_sms() ? tiles : num_sms();
  KPROF(name, stream, kern<<<gx, T2_THREADS, Cfg::SMEM, stream>>>(in, in_stats, gam, bet, wpack, bias, out, out_stats, n));
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
