// Edge phase of the decoder GNN on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
// Included by rollout.cu after edge_mma.cuh (same translation unit: StepArgs / ModelDev / Tape, tc.cuh, wpipe.cuh).
//
// reference: src/models/interaction_net.py:139-184 (message(): edge_mlp on [x_i, x_j, sem_i, sem_j, rel], max aggregation at
// the target :92, zeros for edge-less nodes :187-188); MLP = Linear, LayerNorm, ReLU, Linear, LayerNorm, ReLU, Linear
// (models/common.py:26-44).
//
// Tile = 128 EDGES (rows) x the 128 / 64 output columns of the two dense layers.  The first layer is factorised (rollout.cu:
// h1 = P_i + Q_j + W_rel rel) and costs 6 flops per element, so the rows are built by the CUDA cores straight into the
// canonical K-major shared-memory operand layout; the two dense contractions
//     D1[128 x 128] = relu(LN(h1)) . W3^T        D2[128 x 64] = relu(LN(D1 + b)) . W6^T
// run as tcgen05.mma kind::f16 with an FP16 hi/lo split of BOTH operands (a.w ~= a_hi.w_hi + a_lo.w_hi + a_hi.w_lo, dropped
// a_lo.w_lo ~ 2^-22; fp16 pairs carry 11 + 11 mantissa bits = the TF32 x 3 fidelity of the mma.sync kernels; bf16 pairs (8 + 8) measured
// 1e-5 on the aggregated messages, 4x worse, and flipped arg-max routings), accumulating in fp32 in TMEM.  A thread owns HALF a row
// (64 columns) of the tile in every element-wise phase -- tcgen05.ld 32x32b hands lane l of warp w row 32 (w % 4) + l, and warps
// w, w + 4 read the two column halves of the same rows -- so LayerNorm needs one 2-float exchange per row instead of shuffles
// and no activation ever round-trips through shared memory except as the next MMA's operand.
// The edge list of a scene is padded so that a target's edges never straddle a tile: slot = round_up(n - 1, 8) rows per target,
// 128 / slot targets per tile; the max / arg-max over a target's edges is then a tile-local scan of the D2 tile in shared memory.
// Edge-MLP weights (W3, W6 hi / lo in canonical layout, 96 KB) stay resident in shared memory; one CTA per SM, persistent over
// the tiles.  Scenes with more than 129 agents keep the mma.sync kernels (edge_mma.cuh).
#pragma once

#define ET_CPT 32                         // columns of a row per thread (see edge_fwd_tc_kernel)
#define ET_THREADS (128 * (128 / ET_CPT))
#define ET_W3_BYTES 32768                 // 128 (n) x 128 (k) bf16, canonical K-major: ((k>>3)*16 + (n>>3))*128 + (n&7)*16 + (k&7)*2
#define ET_W6_BYTES 16384                 // 64 (n) x 128 (k)
#define ET_W3H_OFF 0
#define ET_W3L_OFF (ET_W3_BYTES)
#define ET_W6H_OFF (2 * ET_W3_BYTES)
#define ET_W6L_OFF (2 * ET_W3_BYTES + ET_W6_BYTES)
#define ET_PACK_BYTES (2 * ET_W3_BYTES + 2 * ET_W6_BYTES)      // 98304
#define ET_A_BYTES 32768                  // one precision of a 128 x 128 activation tile
#define ET_MB_LD 68                       // row pitch (floats) of the D2 staging tile: conflict-free float4 stores at 272 B
#define ET_SMEM (ET_PACK_BYTES + 2 * ET_A_BYTES)
#define ETB_CPT 32                        // columns per thread of the backward kernel
#define ET_MAX_N 129                      // scenes up to 129 agents (128 edges per target = one tile)

// W: native [N][K] row-major fp32 (nn.Linear weight) -> fp16 hi and lo parts of ET_WSCALE * W in the canonical K-major operand
// layout.  The power-of-two scale keeps the lo parts of weights of magnitude 1e-2..1e-1 in the NORMAL fp16 range (exact to undo:
// the accumulators are multiplied by 1 / ET_WSCALE when they leave TMEM).
#define ET_WSCALE 16.0f
__global__ void edge_tc_pack_kernel(const float* __restrict__ W, int N, int K, uint8_t* __restrict__ out_hi, uint8_t* __restrict__ out_lo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * K) return;
  const int n = idx / K, k = idx % K;
  const float w = W[idx] * ET_WSCALE;
  const __half h = __float2half_rn(w);
  const __half l = __float2half_rn(w - __half2float(h));
  const size_t off = ((size_t)((k >> 3) * (N / 8) + (n >> 3)) * 8 + (n & 7)) * 16 + (size_t)(k & 7) * 2;
  *reinterpret_cast<__half*>(out_hi + off) = h;
  *reinterpret_cast<__half*>(out_lo + off) = l;
}

// Tile table of one scene batch: tiles[2 q] = scene, tiles[2 q + 1] = first (local) target of tile q; *ntiles = number of tiles.
// One block; scenes are scanned in chunks of blockDim.x.
__device__ __forceinline__ int et_slot(int n) { const int ne = n - 1; return ne <= 8 ? 8 : ((ne + 7) & ~7); }
__global__ void __launch_bounds__(1024) edge_tc_tiles_kernel(const int32_t* __restrict__ ptr, int S, int32_t* __restrict__ tiles,
                                                             int32_t* __restrict__ ntiles) {
  __shared__ int scan[1024];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int s0 = 0; s0 < S; s0 += 1024) {
    const int s = s0 + threadIdx.x;
    int n = 0, tps = 1, cnt = 0;
    if (s < S) {
      n = ptr[s + 1] - ptr[s];
      tps = max(1, 128 / et_slot(n));                 // (scenes above ET_MAX_N agents never reach this path)
      cnt = (n + tps - 1) / tps;
    }
    scan[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int v = threadIdx.x >= o ? scan[threadIdx.x - o] : 0;
      __syncthreads();
      scan[threadIdx.x] += v;
      __syncthreads();
    }
    const int off = base + scan[threadIdx.x] - cnt;
    for (int k = 0; k < cnt; k++) {
      tiles[2 * (off + k)] = s;
      tiles[2 * (off + k) + 1] = k * tps;
    }
    __syncthreads();
    if (threadIdx.x == 1023) base += scan[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *ntiles = base;
}

struct EtRow {
  int i, j, lj;        // global target / source agent, local source index
  bool valid;
};

// (scene, first target) of tile q -> the edge this thread's row stands for
__device__ __forceinline__ EtRow et_row(const StepArgs& a, const int32_t* __restrict__ tiles, int q, int row, int& p0, int& n, int& slot, int& first) {
  const int s = tiles[2 * q];
  first = tiles[2 * q + 1];
  p0 = a.ptr[s];
  n = a.ptr[s + 1] - p0;
  slot = et_slot(n);
  const int k = row / slot, e = row - k * slot;
  const int li = first + k;
  EtRow r;
  r.valid = (k < 128 / slot) && (li < n) && (e < n - 1);
  const int lis = r.valid ? li : 0, es = r.valid ? e : 0;
  r.lj = es + (es >= lis ? 1 : 0);
  r.i = p0 + lis;
  r.j = r.valid ? p0 + r.lj : r.i;
  return r;
}

// v (this thread's CPT columns of a 128-wide row) <- relu(LayerNorm(v) * gam + bet); the other parts of the row live in the
// threads 128, 256, ... places away: partial sums meet in s_part.  Two-pass statistics as nn.LayerNorm (eps 1e-5).
template <int CPT>
__device__ __forceinline__ void et_ln_relu(float (&v)[CPT], const float* __restrict__ s_gam, const float* __restrict__ s_bet, float* s_part, int row,
                                           int part) {
  constexpr int PARTS = 128 / CPT;
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CPT; c++) s += v[c];
  s_part[part * 128 + row] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int p = 0; p < PARTS; p++) tot += s_part[p * 128 + row];
  const float mean = tot * (1.0f / 128.0f);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < CPT; c++) { const float d = v[c] - mean; q = fmaf(d, d, q); }
  s_part[512 + part * 128 + row] = q;
  __syncthreads();
  float qt = 0.f;
#pragma unroll
  for (int p = 0; p < PARTS; p++) qt += s_part[512 + p * 128 + row];
  const float rstd = 1.0f / sqrtf(qt * (1.0f / 128.0f) + LN_EPS);
#pragma unroll
  for (int c = 0; c < CPT; c++) v[c] = fmaxf(fmaf((v[c] - mean) * rstd, s_gam[part * CPT + c], s_bet[part * CPT + c]), 0.f);
}

// this thread's CPT columns -> fp16 hi / lo operand rows (canonical K-major, 128 rows): CPT / 8 k-groups of 16 bytes each
template <int CPT>
__device__ __forceinline__ void et_store_operand(const float (&v)[CPT], bool valid, uint8_t* sA, int row, int part) {
#pragma unroll
  for (int g = 0; g < CPT / 8; g++) {
    uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = make_uint4(0u, 0u, 0u, 0u);
    if (valid) {
      tc::split_pack2_f16(v[8 * g + 0], v[8 * g + 1], hi.x, lo.x);
      tc::split_pack2_f16(v[8 * g + 2], v[8 * g + 3], hi.y, lo.y);
      tc::split_pack2_f16(v[8 * g + 4], v[8 * g + 5], hi.z, lo.z);
      tc::split_pack2_f16(v[8 * g + 6], v[8 * g + 7], hi.w, lo.w);
    }
    const int unit = ((part * (CPT / 8) + g) * 16 + (row >> 3)) * 8 + (row & 7);
    *reinterpret_cast<uint4*>(sA + (size_t)unit * 16) = hi;
    *reinterpret_cast<uint4*>(sA + ET_A_BYTES + (size_t)unit * 16) = lo;
  }
}

// D[tmem] = A (128 x 128, hi/lo in sA) . B^T (N x 128, hi/lo packs): 8 K steps x 3 split terms, issued by one thread
__device__ __forceinline__ void et_issue_gemm(uint32_t d_tmem, uint32_t sA_addr, uint32_t sBh_addr, uint32_t sBl_addr, int N, uint64_t* bar) {
  const uint32_t idesc = tc::idesc_f16_f32(128, N);
  const uint32_t lbo_b = (uint32_t)(N / 8) * 128u;
  const uint32_t a_hi = tc::desc_hi(128), b_hi = tc::desc_hi(128);
  const uint32_t ah0 = tc::desc_lo(sA_addr, 2048), al0 = tc::desc_lo(sA_addr + ET_A_BYTES, 2048);
  const uint32_t bh0 = tc::desc_lo(sBh_addr, lbo_b), bl0 = tc::desc_lo(sBl_addr, lbo_b);
#pragma unroll
  for (int ks = 0; ks < 8; ks++) {
    const uint64_t ah = tc::desc_make(ah0 + ((ks * 2 * 2048) >> 4), a_hi), al = tc::desc_make(al0 + ((ks * 2 * 2048) >> 4), a_hi);
    const uint64_t bh = tc::desc_make(bh0 + ((ks * 2 * lbo_b) >> 4), b_hi), bl = tc::desc_make(bl0 + ((ks * 2 * lbo_b) >> 4), b_hi);
    tc::mma_bf16(d_tmem, ah, bh, idesc, ks > 0 ? 1u : 0u);
    tc::mma_bf16(d_tmem, al, bh, idesc, 1u);
    tc::mma_bf16(d_tmem, ah, bl, idesc, 1u);
  }
  if (bar != nullptr) tc::mma_commit(bar);
}

// one tile's worth of this thread's first-layer inputs, loaded ahead of use (the loads of tile q + grid fly while the tensor
// core works on tile q): pq = P_i + Q_j for the thread's columns, rel = transform2frame(pos_i, pos_j)
template <int CPT>
struct EtPre {
  EtRow r;
  int p0, n, slot, first;
  float rel[4];
  float pq[CPT];
};

template <int CPT>
__device__ __forceinline__ void et_prefetch(const StepArgs& a, const int32_t* __restrict__ tiles, int q, int row, int part, EtPre<CPT>& o) {
  const int NA = a.NA;
  o.r = et_row(a, tiles, q, row, o.p0, o.n, o.slot, o.first);
  const float* posg = a.tp.pos + (size_t)a.t * NA * 4;
  const float4 a4 = __ldg(reinterpret_cast<const float4*>(posg + (size_t)o.r.i * 4));
  const float4 b4 = __ldg(reinterpret_cast<const float4*>(posg + (size_t)o.r.j * 4));
  const float4* Pi = reinterpret_cast<const float4*>(a.tp.P + ((size_t)a.t * NA + o.r.i) * 128 + part * CPT);
  const float4* Qj = reinterpret_cast<const float4*>(a.tp.Q + ((size_t)a.t * NA + o.r.j) * 128 + part * CPT);
#pragma unroll
  for (int c4 = 0; c4 < CPT / 4; c4++) {
    const float4 p = __ldg(Pi + c4), qv = __ldg(Qj + c4);
    o.pq[c4 * 4] = p.x + qv.x; o.pq[c4 * 4 + 1] = p.y + qv.y; o.pq[c4 * 4 + 2] = p.z + qv.z; o.pq[c4 * 4 + 3] = p.w + qv.w;
  }
  const float pi[4] = {a4.x, a4.y, a4.z, a4.w}, pj[4] = {b4.x, b4.y, b4.z, b4.w};
  t2f_fwd(pi, pj, o.rel);
#pragma unroll
  for (int d = 0; d < 4; d++)
    if (isnan(o.rel[d])) o.rel[d] = 0.f;                            // interaction_net.py:162
}

// CPT columns per thread: 128 / CPT threads share a row -> 128 * 128 / CPT threads.  CPT = 32 (512 threads, 16 warps) keeps twice
// the warps in flight of CPT = 64 for the same registers per SM: the element-wise phases are latency-, not issue-bound.
template <int CPT>
__global__ void __launch_bounds__(128 * (128 / CPT), 1) edge_fwd_tc_kernel(ModelDev M, StepArgs a, const uint8_t* __restrict__ pack,
                                                                          const int32_t* __restrict__ tiles, const int32_t* __restrict__ ntiles_p) {
  constexpr int PARTS = 128 / CPT, NT = 128 * PARTS;
  STRIVE_PDL_TRIGGER();
  extern __shared__ __align__(1024) uint8_t esm[];
  uint8_t* sW = esm;
  uint8_t* sA = esm + ET_PACK_BYTES;
  float* mbuf = reinterpret_cast<float*>(sA);                     // D2 staging tile: aliases the operand buffer once the MMAs are done
  __shared__ __align__(8) uint64_t bar_w, bar_mma;
  __shared__ uint32_t tmem_base;
  __shared__ float s_g1[128], s_b1[128], s_bias3[128], s_g4[128], s_b4[128], s_bias6[64];
  __shared__ __align__(16) float s_wrel[4 * 128];
  __shared__ float s_part[1024];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = (warp & 3) * 32 + lane, part = warp >> 2;
  // ---- prologue: model constants only (weights -> shared memory, LayerNorm / bias vectors, barriers, TMEM) -- runs while the
  // previous kernel of the stream drains; nothing the predecessor writes is touched before the wait below
  if (tid < 128) {
    s_g1[tid] = M.seg[S_E_LN1_G][tid]; s_b1[tid] = M.seg[S_E_LN1_B][tid]; s_bias3[tid] = M.seg[S_E3_B][tid];
    s_g4[tid] = M.seg[S_E_LN4_G][tid]; s_b4[tid] = M.seg[S_E_LN4_B][tid];
    if (tid < 64) s_bias6[tid] = M.seg[S_E6_B][tid];
  }
  for (int k = tid; k < 512; k += NT) s_wrel[k] = M.seg[S_E0_T_REL][k];
  if (tid == 0) {
    tc::mbar_init(&bar_w, 1);
    tc::mbar_init(&bar_mma, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base, 256);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (tid == 0) {
    wp_mbar_expect_tx(&bar_w, ET_PACK_BYTES);
    for (uint32_t off = 0; off < ET_PACK_BYTES; off += 32768u) wp_bulk_g2s(sW + off, pack + off, 32768u, &bar_w);
  }
  const uint32_t tm = tmem_base;
  const uint32_t t_row = tm + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t sA_addr = tc::smem_u32(sA), sW_addr = tc::smem_u32(sW);
  STRIVE_PDL_WAIT_PTRS(tiles, ntiles_p);
  const int ntiles = *ntiles_p;
  const int NA = a.NA;
  uint32_t ph = 0;
  bool w_ready = false;
  EtPre<CPT> cur;
  if ((int)blockIdx.x < ntiles) et_prefetch<CPT>(a, tiles, blockIdx.x, row, part, cur);
  for (int q = blockIdx.x; q < ntiles; q += gridDim.x) {
    const EtRow r = cur.r;
    const int p0 = cur.p0, n = cur.n, slot = cur.slot, first = cur.first;
    // ---- phase A: h1 = P_i + Q_j + W_rel rel (same fmaf order as the SIMT kernel), LN, ReLU -> operand tile
    float v[CPT];
#pragma unroll
    for (int c4 = 0; c4 < CPT / 4; c4++) {
      float x[4] = {cur.pq[c4 * 4], cur.pq[c4 * 4 + 1], cur.pq[c4 * 4 + 2], cur.pq[c4 * 4 + 3]};
#pragma unroll
      for (int d = 0; d < 4; d++) {
        const float4 w = *reinterpret_cast<const float4*>(&s_wrel[d * 128 + part * CPT + c4 * 4]);
        x[0] = fmaf(cur.rel[d], w.x, x[0]); x[1] = fmaf(cur.rel[d], w.y, x[1]); x[2] = fmaf(cur.rel[d], w.z, x[2]); x[3] = fmaf(cur.rel[d], w.w, x[3]);
      }
      v[c4 * 4] = x[0]; v[c4 * 4 + 1] = x[1]; v[c4 * 4 + 2] = x[2]; v[c4 * 4 + 3] = x[3];
    }
    et_ln_relu<CPT>(v, s_g1, s_b1, s_part, row, part);
    et_store_operand<CPT>(v, r.valid, sA, row, part);
    tc::fence_async_smem();
    __syncthreads();
    if (!w_ready) { wp_wait(&bar_w, 0); w_ready = true; }
    if (tid == 0) {
      tc::tc_fence_after();
      et_issue_gemm(tm, sA_addr, sW_addr + ET_W3H_OFF, sW_addr + ET_W3L_OFF, 128, &bar_mma);
    }
    tc::mbar_wait(&bar_mma, ph);
    ph ^= 1u;
    tc::tc_fence_after();
    // ---- phase B: D1 / scale + bias, LN, ReLU -> operand tile (the first GEMM has completed: its operand buffer is free)
#pragma unroll
    for (int cc = 0; cc < CPT / 16; cc++) {
      float t16[16];
      tc::tmem_ld16(t_row + part * CPT + cc * 16, t16);
#pragma unroll
      for (int c = 0; c < 16; c++) v[cc * 16 + c] = fmaf(t16[c], 1.0f / ET_WSCALE, s_bias3[part * CPT + cc * 16 + c]);
    }
    et_ln_relu<CPT>(v, s_g4, s_b4, s_part, row, part);
    et_store_operand<CPT>(v, r.valid, sA, row, part);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      et_issue_gemm(tm + 128, sA_addr, sW_addr + ET_W6H_OFF, sW_addr + ET_W6L_OFF, 64, &bar_mma);
    }
    // the next tile's first-layer inputs leave L2 while the tensor core runs the second GEMM
    if (q + (int)gridDim.x < ntiles) et_prefetch<CPT>(a, tiles, q + gridDim.x, row, part, cur);
    tc::mbar_wait(&bar_mma, ph);
    ph ^= 1u;
    tc::tc_fence_after();
    // ---- phase C: D2 / scale + bias -> staging tile; max / arg-max over each target's rows (smaller source index wins ties, NaN never wins)
    constexpr int C6 = 64 / PARTS;                                  // D2 columns per thread: 32 or 16
#pragma unroll
    for (int cc = 0; cc < C6 / 16; cc++) {
      float t16[16];
      tc::tmem_ld16(t_row + 128 + part * C6 + cc * 16, t16);
#pragma unroll
      for (int c = 0; c < 16; c += 4) {
        const int col = part * C6 + cc * 16 + c;
        *reinterpret_cast<float4*>(mbuf + (size_t)row * ET_MB_LD + col) =
            make_float4(fmaf(t16[c], 1.0f / ET_WSCALE, s_bias6[col]), fmaf(t16[c + 1], 1.0f / ET_WSCALE, s_bias6[col + 1]),
                        fmaf(t16[c + 2], 1.0f / ET_WSCALE, s_bias6[col + 2]), fmaf(t16[c + 3], 1.0f / ET_WSCALE, s_bias6[col + 3]));
      }
    }
    tc::tc_fence_before();
    __syncthreads();
    {
      const int ch = tid & 63;
      const int tps = 128 / slot;
      for (int k = tid >> 6; k < tps; k += NT / 64) {
        const int li = first + k;
        if (li >= n) break;
        float best = -INFINITY;
        int bi = 255;
        const float* col = mbuf + (size_t)(k * slot) * ET_MB_LD + ch;
        for (int e = 0; e < n - 1; e++) {
          const float m = col[(size_t)e * ET_MB_LD];
          if (m > best) { best = m; bi = e + (e >= li ? 1 : 0); }
        }
        a.tp.aggr[((size_t)a.t * NA + p0 + li) * 64 + ch] = bi == 255 ? 0.f : best;
        a.tp.arg[((size_t)a.t * NA + p0 + li) * 64 + ch] = (uint8_t)bi;
      }
    }
    __syncthreads();                     // the staging tile aliases the operand buffer of the next tile
  }
  if (!w_ready) wp_wait(&bar_w, 0);      // never leave with a bulk copy in flight
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc::tmem_dealloc(tm, 256);
  }
}

// ======================================================================================================
// Edge backward on tcgen05.  Per tile of 128 edges:
//   recompute   a1 = relu(LN(h1)) -> D1 = a1 . W3^T (h2 before bias)                 [GEMM1, same operands as the forward kernel]
//   route       dm[r][c] = d_aggr[i][c] if arg[i][c] == source(r) else 0             (aggr = 'max': the winner gets the gradient)
//               D3 = dm . W6       (d a2, before the LayerNorm / ReLU adjoint)        [GEMM3, K = 64]
//   adjoint     d h2 = LN-ReLU adjoint(D3; h2 = D1 / s + b) -> D4 = d h2 . W3        [GEMM4]
//               d h1 = LN-ReLU adjoint(D4; h1 recomputed from P_i + Q_j + W_rel rel)
//   scatter     dP_i = sum over the target's rows (tile-local scan), dQ_j += d h1 (atomics), d rel -> transform2frame adjoint ->
//               g_pos_j (atomics) and g_pos_i (tile-local sum, then atomics)
// GEMM3 / GEMM4 multiply by the TRANSPOSED weight matrices: they read the SAME shared-memory packs as the forward GEMMs through
// MN-major operand descriptors (instruction-descriptor bit 16; SBO = stride of the core matrices along the pack's K axis, LBO =
// 128 B along its row axis), so the 96 KB of resident weights serve both directions and no transposed copy exists.
// No per-edge activation is stored between forward and backward (it would be 4 GB per rollout step at BASELINE configs[4]).
// ======================================================================================================
#define ETB_DM_BYTES 16384                // one precision of the 128 x 64 routed-gradient operand tile
#define ETB_SMEM (ET_PACK_BYTES + 2 * ET_A_BYTES + 2 * ETB_DM_BYTES)
#define ETB_ST_LD 132                     // row pitch (floats) of the d h1 staging tile (conflict-free float4 stores); 128 x 132 x 4 B <= 96 KB

// D[tmem] (+)= A (128 x 16*ksteps, K-major hi/lo at sA_h / sA_l) . pack^T, pack = [R rows][128] K-major hi/lo (R = 16*ksteps): MN-major B view
__device__ __forceinline__ void et_issue_gemm_t(uint32_t d_tmem, uint32_t sA_h, uint32_t sA_l, uint32_t sBh_addr, uint32_t sBl_addr, int R, int ksteps) {
  const uint32_t idesc = tc::idesc_f16_f32(128, 128) | (1u << 16);          // b_major = MN
  const uint32_t sbo_b = (uint32_t)(R / 8) * 128u;
  const uint32_t a_hi = tc::desc_hi(128), b_hi = tc::desc_hi(sbo_b);
  const uint32_t ah0 = tc::desc_lo(sA_h, 2048), al0 = tc::desc_lo(sA_l, 2048);
  const uint32_t bh0 = tc::desc_lo(sBh_addr, 128), bl0 = tc::desc_lo(sBl_addr, 128);
  for (int ks = 0; ks < ksteps; ks++) {
    const uint64_t ah = tc::desc_make(ah0 + ((ks * 2 * 2048) >> 4), a_hi), al = tc::desc_make(al0 + ((ks * 2 * 2048) >> 4), a_hi);
    const uint64_t bh = tc::desc_make(bh0 + ((ks * 256) >> 4), b_hi), bl = tc::desc_make(bl0 + ((ks * 256) >> 4), b_hi);
    tc::mma_bf16(d_tmem, ah, bh, idesc, ks > 0 ? 1u : 0u);
    tc::mma_bf16(d_tmem, al, bh, idesc, 1u);
    tc::mma_bf16(d_tmem, ah, bl, idesc, 1u);
  }
}

// row statistics of a 128-wide row split over PARTS threads (two-pass, as et_ln_relu)
template <int CPT>
__device__ __forceinline__ void et_ln_stats(const float (&v)[CPT], float* s_part, int row, int part, float& mean, float& rstd) {
  constexpr int PARTS = 128 / CPT;
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CPT; c++) s += v[c];
  s_part[part * 128 + row] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int p = 0; p < PARTS; p++) tot += s_part[p * 128 + row];
  mean = tot * (1.0f / 128.0f);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < CPT; c++) { const float d = v[c] - mean; q = fmaf(d, d, q); }
  s_part[512 + part * 128 + row] = q;
  __syncthreads();
  float qt = 0.f;
#pragma unroll
  for (int p = 0; p < PARTS; p++) qt += s_part[512 + p * 128 + row];
  rstd = 1.0f / sqrtf(qt * (1.0f / 128.0f) + LN_EPS);
}

// adjoint of h = relu(LN(a) * gam + bet): dh (in: d h, out: d a); a = pre-LN activations, (mean, rstd) their row statistics
template <int CPT>
__device__ __forceinline__ void et_ln_relu_bwd(float (&dh)[CPT], const float (&a)[CPT], float mean, float rstd, const float* __restrict__ s_gam,
                                               const float* __restrict__ s_bet, float* s_red, int row, int part) {
  constexpr int PARTS = 128 / CPT;
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int c = 0; c < CPT; c++) {
    const float g = s_gam[part * CPT + c];
    const float xh = (a[c] - mean) * rstd;
    const float y = fmaf(xh, g, s_bet[part * CPT + c]);
    const float dg = (y > 0.f) ? dh[c] * g : 0.f;
    dh[c] = dg;
    s1 += dg;
    s2 = fmaf(dg, xh, s2);
  }
  s_red[part * 128 + row] = s1;
  s_red[512 + part * 128 + row] = s2;
  __syncthreads();
  float t1 = 0.f, t2 = 0.f;
#pragma unroll
  for (int p = 0; p < PARTS; p++) { t1 += s_red[p * 128 + row]; t2 += s_red[512 + p * 128 + row]; }
  t1 *= (1.0f / 128.0f);
  t2 *= (1.0f / 128.0f);
#pragma unroll
  for (int c = 0; c < CPT; c++) {
    const float xh = (a[c] - mean) * rstd;
    dh[c] = rstd * (dh[c] - t1 - xh * t2);
  }
}

template <int CPT>
__device__ __forceinline__ void et_h1(const StepArgs& a, const EtRow& r, const float (&rel)[4], const float* __restrict__ s_wrel, int part, float (&v)[CPT]) {
  const int NA = a.NA;
  const float4* Pi = reinterpret_cast<const float4*>(a.tp.P + ((size_t)a.t * NA + r.i) * 128 + part * CPT);
  const float4* Qj = reinterpret_cast<const float4*>(a.tp.Q + ((size_t)a.t * NA + r.j) * 128 + part * CPT);
#pragma unroll
  for (int c4 = 0; c4 < CPT / 4; c4++) {
    const float4 p = __ldg(Pi + c4), qv = __ldg(Qj + c4);
    float x[4] = {p.x + qv.x, p.y + qv.y, p.z + qv.z, p.w + qv.w};
#pragma unroll
    for (int d = 0; d < 4; d++) {
      const float4 w = *reinterpret_cast<const float4*>(&s_wrel[d * 128 + part * CPT + c4 * 4]);
      x[0] = fmaf(rel[d], w.x, x[0]); x[1] = fmaf(rel[d], w.y, x[1]); x[2] = fmaf(rel[d], w.z, x[2]); x[3] = fmaf(rel[d], w.w, x[3]);
    }
    v[c4 * 4] = x[0]; v[c4 * 4 + 1] = x[1]; v[c4 * 4 + 2] = x[2]; v[c4 * 4 + 3] = x[3];
  }
}

template <int CPT>
__global__ void __launch_bounds__(128 * (128 / CPT), 1) edge_bwd_tc_kernel(ModelDev M, StepArgs a, const uint8_t* __restrict__ pack,
                                                                          const int32_t* __restrict__ tiles, const int32_t* __restrict__ ntiles_p) {
  constexpr int PARTS = 128 / CPT, NT = 128 * PARTS, C6 = 64 / PARTS;
  STRIVE_PDL_TRIGGER();
  extern __shared__ __align__(1024) uint8_t esm[];
  uint8_t* sW = esm;
  uint8_t* sA = esm + ET_PACK_BYTES;
  uint8_t* sDM = sA + 2 * ET_A_BYTES;
  float* stg = reinterpret_cast<float*>(sA);                      // d h1 staging tile: operand buffers are dead by then (sA + sDM = 96 KB)
  __shared__ __align__(8) uint64_t bar_w, bar_mma;
  __shared__ uint32_t tmem_base;
  __shared__ float s_g1[128], s_b1[128], s_bias3[128], s_g4[128], s_b4[128];
  __shared__ __align__(16) float s_wrel[4 * 128];
  __shared__ float s_part[1024], s_red[1024], s_drel[4 * 512];
  __shared__ float s_dpi[16 * 4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = (warp & 3) * 32 + lane, part = warp >> 2;
  if (tid < 128) {
    s_g1[tid] = M.seg[S_E_LN1_G][tid]; s_b1[tid] = M.seg[S_E_LN1_B][tid]; s_bias3[tid] = M.seg[S_E3_B][tid];
    s_g4[tid] = M.seg[S_E_LN4_G][tid]; s_b4[tid] = M.seg[S_E_LN4_B][tid];
  }
  for (int k = tid; k < 512; k += NT) s_wrel[k] = M.seg[S_E0_T_REL][k];
  if (tid < 64) s_dpi[tid] = 0.f;
  if (tid == 0) {
    tc::mbar_init(&bar_w, 1);
    tc::mbar_init(&bar_mma, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base, 256);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (tid == 0) {
    wp_mbar_expect_tx(&bar_w, ET_PACK_BYTES);
    for (uint32_t off = 0; off < ET_PACK_BYTES; off += 32768u) wp_bulk_g2s(sW + off, pack + off, 32768u, &bar_w);
  }
  const uint32_t tm = tmem_base;
  const uint32_t t_row = tm + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t sA_addr = tc::smem_u32(sA), sDM_addr = tc::smem_u32(sDM), sW_addr = tc::smem_u32(sW);
  STRIVE_PDL_WAIT_PTRS(tiles, ntiles_p);
  const int ntiles = *ntiles_p;
  const int NA = a.NA;
  const float* posg = a.tp.pos + (size_t)a.t * NA * 4;
  uint32_t ph = 0;
  bool w_ready = false;
  for (int q = blockIdx.x; q < ntiles; q += gridDim.x) {
    int p0, n, slot, first;
    const EtRow r = et_row(a, tiles, q, row, p0, n, slot, first);
    float pi[4], pj[4], rel[4];
    unsigned nanmask = 0u;
    {
      const float4 a4 = __ldg(reinterpret_cast<const float4*>(posg + (size_t)r.i * 4));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(posg + (size_t)r.j * 4));
      pi[0] = a4.x; pi[1] = a4.y; pi[2] = a4.z; pi[3] = a4.w;
      pj[0] = b4.x; pj[1] = b4.y; pj[2] = b4.z; pj[3] = b4.w;
      t2f_fwd(pi, pj, rel);
#pragma unroll
      for (int d = 0; d < 4; d++)
        if (isnan(rel[d])) { rel[d] = 0.f; nanmask |= 1u << d; }     // interaction_net.py:162
    }
    // ---- phase A: recompute a1 -> operand tile; routed gradient dm -> its operand tile
    float mean1, rstd1;
    {
      float v[CPT];
      et_h1<CPT>(a, r, rel, s_wrel, part, v);
      et_ln_stats<CPT>(v, s_part, row, part, mean1, rstd1);
#pragma unroll
      for (int c = 0; c < CPT; c++) v[c] = fmaxf(fmaf((v[c] - mean1) * rstd1, s_g1[part * CPT + c], s_b1[part * CPT + c]), 0.f);
      et_store_operand<CPT>(v, r.valid, sA, row, part);
    }
    {
      // this thread's C6 of the 64 message channels: one k-group of 8 columns per 16-byte operand row
      const uint8_t* argp = a.tp.arg + ((size_t)a.t * NA + r.i) * 64 + part * C6;
      const float* dagp = a.tp.d_aggr + (size_t)r.i * 64 + part * C6;
#pragma unroll
      for (int g = 0; g < C6 / 8; g++) {
        const uint2 ab = __ldg(reinterpret_cast<const uint2*>(argp + 8 * g));
        const float4 d0 = __ldg(reinterpret_cast<const float4*>(dagp + 8 * g)), d1 = __ldg(reinterpret_cast<const float4*>(dagp + 8 * g + 4));
        const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        float m8[8];
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const unsigned am = ((c < 4 ? ab.x : ab.y) >> (8 * (c & 3))) & 0xffu;
          m8[c] = (r.valid && (int)am == r.lj) ? dv[c] : 0.f;
        }
        uint4 hi, lo;
        tc::split_pack2_f16(m8[0], m8[1], hi.x, lo.x);
        tc::split_pack2_f16(m8[2], m8[3], hi.y, lo.y);
        tc::split_pack2_f16(m8[4], m8[5], hi.z, lo.z);
        tc::split_pack2_f16(m8[6], m8[7], hi.w, lo.w);
        const int unit = ((part * (C6 / 8) + g) * 16 + (row >> 3)) * 8 + (row & 7);
        *reinterpret_cast<uint4*>(sDM + (size_t)unit * 16) = hi;
        *reinterpret_cast<uint4*>(sDM + ETB_DM_BYTES + (size_t)unit * 16) = lo;
      }
    }
    tc::fence_async_smem();
    __syncthreads();
    if (!w_ready) { wp_wait(&bar_w, 0); w_ready = true; }
    if (tid == 0) {
      tc::tc_fence_after();
      et_issue_gemm(tm, sA_addr, sW_addr + ET_W3H_OFF, sW_addr + ET_W3L_OFF, 128, nullptr);                        // D1: cols 0..127
      et_issue_gemm_t(tm + 128, sDM_addr, sDM_addr + ETB_DM_BYTES, sW_addr + ET_W6H_OFF, sW_addr + ET_W6L_OFF, 64, 4);   // D3: cols 128..255
      tc::mma_commit(&bar_mma);
    }
    tc::mbar_wait(&bar_mma, ph);
    ph ^= 1u;
    tc::tc_fence_after();
    // ---- phase C: h2 = D1 / s + b, its LayerNorm statistics; d a2 = D3 / s; LayerNorm-ReLU adjoint -> d h2 operand tile
    {
      float h2[CPT], dh[CPT];
#pragma unroll
      for (int cc = 0; cc < CPT / 16; cc++) {
        float t16[16];
        tc::tmem_ld16(t_row + part * CPT + cc * 16, t16);
#pragma unroll
        for (int c = 0; c < 16; c++) h2[cc * 16 + c] = fmaf(t16[c], 1.0f / ET_WSCALE, s_bias3[part * CPT + cc * 16 + c]);
        tc::tmem_ld16(t_row + 128 + part * CPT + cc * 16, t16);
#pragma unroll
        for (int c = 0; c < 16; c++) dh[cc * 16 + c] = t16[c] * (1.0f / ET_WSCALE);
      }
      float mean2, rstd2;
      et_ln_stats<CPT>(h2, s_part, row, part, mean2, rstd2);
      et_ln_relu_bwd<CPT>(dh, h2, mean2, rstd2, s_g4, s_b4, s_red, row, part);
      et_store_operand<CPT>(dh, r.valid, sA, row, part);
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      et_issue_gemm_t(tm, sA_addr, sA_addr + ET_A_BYTES, sW_addr + ET_W3H_OFF, sW_addr + ET_W3L_OFF, 128, 8);      // D4 over D1
      tc::mma_commit(&bar_mma);
    }
    // ---- phase E: d h1 = LayerNorm-ReLU adjoint(D4 / s; h1 recomputed); scatter.  The P / Q rows of the recompute leave L2 while the
    // tensor core runs GEMM4.
    {
      float h1[CPT], d1[CPT];
      et_h1<CPT>(a, r, rel, s_wrel, part, h1);
      tc::mbar_wait(&bar_mma, ph);
      ph ^= 1u;
      tc::tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < CPT / 16; cc++) {
        float t16[16];
        tc::tmem_ld16(t_row + part * CPT + cc * 16, t16);
#pragma unroll
        for (int c = 0; c < 16; c++) d1[cc * 16 + c] = t16[c] * (1.0f / ET_WSCALE);
      }
      tc::tc_fence_before();
      et_ln_relu_bwd<CPT>(d1, h1, mean1, rstd1, s_g1, s_b1, s_red, row, part);
      if (!r.valid) {
#pragma unroll
        for (int c = 0; c < CPT; c++) d1[c] = 0.f;                     // padding rows carry exact zeros anyway (dm = 0); keep NaNs of unused lanes out
      }
      // d rel = W_rel d h1 (partial over this thread's columns)
      float dr[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < CPT; c++) {
#pragma unroll
        for (int d = 0; d < 4; d++) dr[d] = fmaf(d1[c], s_wrel[d * 128 + part * CPT + c], dr[d]);
      }
#pragma unroll
      for (int d = 0; d < 4; d++) s_drel[d * 512 + part * 128 + row] = dr[d];
      // staging tile for the per-target column sums (all MMAs of this tile have completed: the operand buffers are free)
#pragma unroll
      for (int c = 0; c < CPT; c += 4)
        *reinterpret_cast<float4*>(stg + (size_t)row * ETB_ST_LD + part * CPT + c) = make_float4(d1[c], d1[c + 1], d1[c + 2], d1[c + 3]);
    }
    __syncthreads();
    if (part == 0) {                                               // warps 0..3: one thread per row
      float dpi[4] = {0.f, 0.f, 0.f, 0.f}, dpj[4] = {0.f, 0.f, 0.f, 0.f};
      if (r.valid) {
        float drel[4];
#pragma unroll
        for (int d = 0; d < 4; d++) {
          float s = 0.f;
#pragma unroll
          for (int p = 0; p < PARTS; p++) s += s_drel[d * 512 + p * 128 + row];
          drel[d] = ((nanmask >> d) & 1u) ? 0.f : s;
        }
        t2f_bwd(pi, pj, drel, dpi, dpj);
#pragma unroll
        for (int d = 0; d < 4; d++) atomicAdd(a.tp.g_pos + (size_t)r.j * 4 + d, dpj[d]);
      }
      // position adjoint of the TARGET: the rows of one target are consecutive, so a warp holds at most 32 / slot + 1 targets --
      // reduce per target with shuffles and let one lane add the sum (32 lanes hitting one shared-memory word serialise)
      const int k = row / slot;
      unsigned todo = 0xffffffffu;
      while (todo) {
        const int leader = __ffs(todo) - 1;
        const int kk = __shfl_sync(0xffffffffu, k, leader);
        const bool mine = (k == kk);
        float v[4];
#pragma unroll
        for (int d = 0; d < 4; d++) {
          v[d] = mine ? dpi[d] : 0.f;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v[d] += __shfl_xor_sync(0xffffffffu, v[d], o);
        }
        if (lane == leader) {
#pragma unroll
          for (int d = 0; d < 4; d++) atomicAdd(&s_dpi[kk * 4 + d], v[d]);
        }
        todo &= ~__ballot_sync(0xffffffffu, mine);
      }
    }
    {
      const int col = tid & 127;
      const int tps = 128 / slot;
      // dP_i = sum over the target's rows
      for (int k = tid >> 7; k < tps; k += NT / 128) {
        const int li = first + k;
        if (li >= n) break;
        const float* cp = stg + (size_t)(k * slot) * ETB_ST_LD + col;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;                 // four independent chains: the scan is shared-memory latency bound
        int e = 0;
        for (; e + 4 <= n - 1; e += 4) {
          s0 += cp[(size_t)e * ETB_ST_LD];
          s1 += cp[(size_t)(e + 1) * ETB_ST_LD];
          s2 += cp[(size_t)(e + 2) * ETB_ST_LD];
          s3 += cp[(size_t)(e + 3) * ETB_ST_LD];
        }
        for (; e < n - 1; e++) s0 += cp[(size_t)e * ETB_ST_LD];
        a.tp.dP[(size_t)(p0 + li) * 128 + col] = (s0 + s1) + (s2 + s3);
      }
      // dQ_j += sum over the tile's targets of the row that has j as its source: ONE atomic per (source, column) and tile instead of
      // one per edge (the 8.1 M per-edge float atomics of a rollout step bound this kernel and its mma.sync predecessor alike:
      // ncu stall_lg 29 %, lts RED requests 8.16 M)
      for (int lj = tid >> 7; lj < n; lj += NT / 128) {
        float s = 0.f;
        const int kmax = min(tps, n - first);
#pragma unroll 4
        for (int k = 0; k < kmax; k++) {
          const int li = first + k;
          const int e = max(min(lj - (lj > li ? 1 : 0), n - 2), 0);    // (clamped: the li == lj term is masked below)
          const float v = stg[(k * slot + e) * ETB_ST_LD + col];
          s += (li != lj) ? v : 0.f;
        }
        atomicAdd(a.tp.dQ + (size_t)(p0 + lj) * 128 + col, s);
      }
    }
    __syncthreads();
    if (tid < 64) {
      const int k = tid >> 2;
      if (k < 128 / slot && first + k < n) atomicAdd(a.tp.g_pos + (size_t)(p0 + first + k) * 4 + (tid & 3), s_dpi[tid]);
      s_dpi[tid] = 0.f;
    }
    __syncthreads();                     // staging tile / s_dpi are reused by the next tile
  }
  if (!w_ready) wp_wait(&bar_w, 0);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc::tmem_dealloc(tm, 256);
  }
}
