// Edge phase of the decoder GNN on the warp-level tensor path (mma.sync m16n8k8 TF32, fp32 accumulate).
// Included by rollout.cu after StepArgs / ModelDev / EdgeCtx (same translation unit).
//
// reference: src/models/interaction_net.py:139-184 (message: edge_mlp on [x_i, x_j, sem_i, sem_j, rel], max aggregation :92,
// zeros for edge-less nodes :187-188); MLP = Linear, LayerNorm, ReLU, Linear, LayerNorm, ReLU, Linear (models/common.py:26-44).
//
// Why: the fp32 SIMT edge kernels were the largest non-encoder item of an iteration (edge_fwd 129 us + edge_bwd 270 us per
// rollout step at BASELINE configs[1]) and sat at a third of the FFMA peak; measured on this B200 (scripts/mma_sync_bench.cu)
// mma.sync TF32 sustains 512 MAC/clk/SM = 4x the FFMA rate.  fp32 fidelity is kept with the 3-term split
//     x = hi + lo (hi = tf32 truncation of x, lo = x - hi exactly):   x*w ~= lo*w_hi + hi*w_lo + hi*w_hi   (dropped: lo*lo ~ 2^-20)
// A warp owns a tile of 16 edges of ONE target agent and keeps every activation in registers in the accumulator ("C")
// fragment layout: lane (g = lane/4, t = lane%4) holds rows g, g+8 and, of each 8-column block j, columns 8j+2t, 8j+2t+1:
//     v[j] = { (g, 8j+2t), (g, 8j+2t+1), (g+8, 8j+2t), (g+8, 8j+2t+1) }
// The same registers are the A operand of the next GEMM when the k index inside a block is permuted (k = t <-> column 2t,
// k = t+4 <-> column 2t+1); the weight fragments are packed once per model with that permutation, so LayerNorm/ReLU between
// the layers never leaves the register file and there is no shared-memory activation buffer at all.
// Weight fragments: block (j, nt) = 32 lanes x {b0_hi, b1_hi, b0_lo, b1_lo},  b0 = W[8j+2t][8nt+g], b1 = W[8j+2t+1][8nt+g]
//   forward : fp32 words (tf32 hi / tf32 lo), 192 KB resident in shared memory (E3_T, E6_T)
//   backward: bf16 hi / bf16 lo pairs (exact as tf32 operands; 2^-17 weight precision is ample for the adjoint), 160 KB resident
#pragma once

#ifndef EM_WARPS
#define EM_WARPS 8
#endif
#ifndef EM_NG
#define EM_NG 4      // column blocks per group: dependency distance of an accumulator's three split terms
#endif
#define EM_THREADS (EM_WARPS * 32)
#define EM_F3_OFF 0                       // forward pack of E3_T  [16][16][32][4] fp32
#define EM_F6_OFF 131072                  // forward pack of E6_T  [16][8][32][4] fp32
#define EM_B3T_OFF 196608                 // backward pack of E3_T [16][16][32][4] bf16
#define EM_B6N_OFF 262144                 // backward pack of E6_N [8][16][32][4] bf16
#define EM_B3N_OFF 294912                 // backward pack of E3_N [16][16][32][4] bf16
#define EM_FRAG_BYTES 360448
#define EM_FWD_SMEM (EM_B3T_OFF)
#define EM_BWD_SMEM (EM_FRAG_BYTES - EM_B3T_OFF)

__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// W: [K][N] row-major.  One thread per (j, nt, lane).
__global__ void edge_frag_pack_kernel(const float* __restrict__ W, int K, int N, float* __restrict__ out_f32, uint16_t* __restrict__ out_bf16) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int NT = N / 8;
  if (idx >= (K / 8) * NT * 32) return;
  const int lane = idx & 31, nt = (idx >> 5) % NT, j = (idx >> 5) / NT;
  const int g = lane >> 2, t = lane & 3;
  const float w0 = W[(size_t)(8 * j + 2 * t) * N + 8 * nt + g];
  const float w1 = W[(size_t)(8 * j + 2 * t + 1) * N + 8 * nt + g];
  if (out_f32 != nullptr) {
    const float h0 = tf32_trunc(w0), h1 = tf32_trunc(w1);
    out_f32[(size_t)idx * 4 + 0] = h0;
    out_f32[(size_t)idx * 4 + 1] = h1;
    out_f32[(size_t)idx * 4 + 2] = tf32_trunc(w0 - h0);
    out_f32[(size_t)idx * 4 + 3] = tf32_trunc(w1 - h1);
  } else {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(w0), h1 = __float2bfloat16_rn(w1);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(w0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(w1 - __bfloat162float(h1));
    out_bf16[(size_t)idx * 4 + 0] = __bfloat16_as_ushort(h0);
    out_bf16[(size_t)idx * 4 + 1] = __bfloat16_as_ushort(h1);
    out_bf16[(size_t)idx * 4 + 2] = __bfloat16_as_ushort(l0);
    out_bf16[(size_t)idx * 4 + 3] = __bfloat16_as_ushort(l1);
  }
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// acc[nt] += h (16 x 8*JK, C layout) . W (8*JK x 8*NT), W given as a fragment pack in shared memory
template <int JK, int NT, bool BF16>
__device__ __forceinline__ void frag_gemm(const float (&h)[JK][4], float (&acc)[NT][4], const uint8_t* __restrict__ pack, int lane) {
#pragma unroll
  for (int j = 0; j < JK; j++) {
    uint32_t ahi[4], alo[4];
    const float av[4] = {h[j][0], h[j][2], h[j][1], h[j][3]};     // a0 (g,k=t) a1 (g+8,k=t) a2 (g,k=t+4) a3 (g+8,k=t+4)
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const float hi = tf32_trunc(av[q]);
      ahi[q] = __float_as_uint(hi);
      alo[q] = __float_as_uint(av[q] - hi);
    }
    // groups of EM_NG column blocks: the three split terms of one accumulator are EM_NG MMAs apart
#pragma unroll
    for (int n0 = 0; n0 < NT; n0 += EM_NG) {
      uint32_t bh[EM_NG][2], bl[EM_NG][2];
#pragma unroll
      for (int u = 0; u < EM_NG; u++) {
        const int nt = n0 + u;
        if constexpr (BF16) {
          const uint2 w = *reinterpret_cast<const uint2*>(pack + ((size_t)(j * NT + nt) * 32 + lane) * 8);
          bh[u][0] = w.x << 16; bh[u][1] = w.x & 0xffff0000u; bl[u][0] = w.y << 16; bl[u][1] = w.y & 0xffff0000u;
        } else {
          const uint4 w = *reinterpret_cast<const uint4*>(pack + ((size_t)(j * NT + nt) * 32 + lane) * 16);
          bh[u][0] = w.x; bh[u][1] = w.y; bl[u][0] = w.z; bl[u][1] = w.w;
        }
      }
#pragma unroll
      for (int u = 0; u < EM_NG; u++) mma_tf32(acc[n0 + u], alo, bh[u][0], bh[u][1]);
#pragma unroll
      for (int u = 0; u < EM_NG; u++) mma_tf32(acc[n0 + u], ahi, bl[u][0], bl[u][1]);
#pragma unroll
      for (int u = 0; u < EM_NG; u++) mma_tf32(acc[n0 + u], ahi, bh[u][0], bh[u][1]);
    }
  }
}

__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// per-row LayerNorm statistics of a 128-wide C-layout tensor: index 0 = row g, 1 = row g+8
__device__ __forceinline__ void frag_ln_stats(const float (&v)[16][4], float (&mean)[2], float (&rstd)[2]) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 16; j++) { s0 += v[j][0] + v[j][1]; s1 += v[j][2] + v[j][3]; }
  mean[0] = quad_sum(s0) * (1.0f / 128.0f);
  mean[1] = quad_sum(s1) * (1.0f / 128.0f);
  float q0 = 0.f, q1 = 0.f;
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const float d0 = v[j][0] - mean[0], d1 = v[j][1] - mean[0], d2 = v[j][2] - mean[1], d3 = v[j][3] - mean[1];
    q0 = fmaf(d0, d0, q0); q0 = fmaf(d1, d1, q0);
    q1 = fmaf(d2, d2, q1); q1 = fmaf(d3, d3, q1);
  }
  rstd[0] = 1.0f / sqrtf(quad_sum(q0) * (1.0f / 128.0f) + LN_EPS);
  rstd[1] = 1.0f / sqrtf(quad_sum(q1) * (1.0f / 128.0f) + LN_EPS);
}

// v <- relu(LayerNorm(v) * gam + bet) in place (models/common.py:33-35)
__device__ __forceinline__ void frag_ln_relu(float (&v)[16][4], const float* __restrict__ gam, const float* __restrict__ bet, int t) {
  float mean[2], rstd[2];
  frag_ln_stats(v, mean, rstd);
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const float2 g2 = __ldg(reinterpret_cast<const float2*>(gam + 8 * j + 2 * t));
    const float2 b2 = __ldg(reinterpret_cast<const float2*>(bet + 8 * j + 2 * t));
    v[j][0] = fmaxf(fmaf((v[j][0] - mean[0]) * rstd[0], g2.x, b2.x), 0.f);
    v[j][1] = fmaxf(fmaf((v[j][1] - mean[0]) * rstd[0], g2.y, b2.y), 0.f);
    v[j][2] = fmaxf(fmaf((v[j][2] - mean[1]) * rstd[1], g2.x, b2.x), 0.f);
    v[j][3] = fmaxf(fmaf((v[j][3] - mean[1]) * rstd[1], g2.y, b2.y), 0.f);
  }
}

// backward of h = relu(LN(a) * gam + bet): dh (in/out: becomes da), a = pre-LN activations
__device__ __forceinline__ void frag_ln_relu_bwd(float (&dh)[16][4], const float (&a)[16][4], const float* __restrict__ gam,
                                                 const float* __restrict__ bet, int t) {
  float mean[2], rstd[2];
  frag_ln_stats(a, mean, rstd);
  float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const float2 g2 = __ldg(reinterpret_cast<const float2*>(gam + 8 * j + 2 * t));
    const float2 b2 = __ldg(reinterpret_cast<const float2*>(bet + 8 * j + 2 * t));
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int r = e >> 1;
      const float gg = (e & 1) ? g2.y : g2.x, bb = (e & 1) ? b2.y : b2.x;
      const float xh = (a[j][e] - mean[r]) * rstd[r];
      const float y = fmaf(xh, gg, bb);
      const float dg = (y > 0.f) ? dh[j][e] * gg : 0.f;
      dh[j][e] = dg;
      s1[r] += dg;
      s2[r] = fmaf(dg, xh, s2[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < 2; r++) {
    s1[r] = quad_sum(s1[r]) * (1.0f / 128.0f);
    s2[r] = quad_sum(s2[r]) * (1.0f / 128.0f);
  }
#pragma unroll
  for (int j = 0; j < 16; j++) {
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int r = e >> 1;
      const float xh = (a[j][e] - mean[r]) * rstd[r];
      dh[j][e] = rstd[r] * (dh[j][e] - s1[r] - xh * s2[r]);
    }
  }
}

template <int NT>
__device__ __forceinline__ void frag_bias(float (&acc)[NT][4], const float* __restrict__ bias, int t) {
#pragma unroll
  for (int nt = 0; nt < NT; nt++) {
    const float2 b2 = __ldg(reinterpret_cast<const float2*>(bias + 8 * nt + 2 * t));
    acc[nt][0] = b2.x; acc[nt][1] = b2.y; acc[nt][2] = b2.x; acc[nt][3] = b2.y;
  }
}

struct EdgeTile {
  int lj[2], j[2];      // local / global index of the source agent of rows g, g+8
  bool valid[2];
  float rel[2][4];
  float pj[2][4];
};

// rows of tile q of target agent c.i, first edge layer h1pre = P_i + Q_j + W_rel rel in C layout (same fmaf order as edge_layer0)
__device__ __forceinline__ void edge_tile_h1pre(const ModelDev& M, const StepArgs& a, const EdgeCtx& c, int q, int g, int t, EdgeTile& tl,
                                                float (&h)[16][4]) {
  const int NA = a.NA;
  const float* posg = a.tp.pos + (size_t)a.t * NA * 4;
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int e = q * 16 + g + 8 * r;
    tl.valid[r] = e < c.ne;
    const int ee = tl.valid[r] ? e : 0;
    tl.lj[r] = ee + (ee >= c.li ? 1 : 0);
    tl.j[r] = tl.valid[r] ? c.p0 + tl.lj[r] : c.i;
    const float4 p4 = __ldg(reinterpret_cast<const float4*>(posg + (size_t)tl.j[r] * 4));
    tl.pj[r][0] = p4.x; tl.pj[r][1] = p4.y; tl.pj[r][2] = p4.z; tl.pj[r][3] = p4.w;
    t2f_fwd(c.pos_i, tl.pj[r], tl.rel[r]);
#pragma unroll
    for (int d = 0; d < 4; d++)
      if (isnan(tl.rel[r][d])) tl.rel[r][d] = 0.f;          // interaction_net.py:162
  }
  const float* Pi = a.tp.P + ((size_t)a.t * NA + c.i) * 128;
  const float* Q0 = a.tp.Q + ((size_t)a.t * NA + tl.j[0]) * 128;
  const float* Q1 = a.tp.Q + ((size_t)a.t * NA + tl.j[1]) * 128;
  const float* Wr = M.seg[S_E0_T_REL];
#pragma unroll
  for (int j = 0; j < 16; j++) {
    const int col = 8 * j + 2 * t;
    const float2 p2 = __ldg(reinterpret_cast<const float2*>(Pi + col));
    const float2 q0 = __ldg(reinterpret_cast<const float2*>(Q0 + col));
    const float2 q1 = __ldg(reinterpret_cast<const float2*>(Q1 + col));
    float v0 = p2.x + q0.x, v1 = p2.y + q0.y, v2 = p2.x + q1.x, v3 = p2.y + q1.y;
#pragma unroll
    for (int d = 0; d < 4; d++) {
      const float2 w2 = __ldg(reinterpret_cast<const float2*>(Wr + d * 128 + col));
      v0 = fmaf(tl.rel[0][d], w2.x, v0);
      v1 = fmaf(tl.rel[0][d], w2.y, v1);
      v2 = fmaf(tl.rel[1][d], w2.x, v2);
      v3 = fmaf(tl.rel[1][d], w2.y, v3);
    }
    h[j][0] = v0; h[j][1] = v1; h[j][2] = v2; h[j][3] = v3;
  }
}

__device__ __forceinline__ void em_agent_ctx(const StepArgs& a, int i, EdgeCtx& c) {
  c.i = i;
  const int s = a.scene_of[i];
  c.p0 = a.ptr[s];
  c.n = a.ptr[s + 1] - c.p0;
  c.ne = c.n - 1;
  c.li = i - c.p0;
  const float4 p = __ldg(reinterpret_cast<const float4*>(a.tp.pos + ((size_t)a.t * a.NA + i) * 4));
  c.pos_i[0] = p.x; c.pos_i[1] = p.y; c.pos_i[2] = p.z; c.pos_i[3] = p.w;
}

// one elected thread streams `bytes` of fragment packs into shared memory; every consumer waits on `bar` (phase 0) once
__device__ __forceinline__ void em_load_packs(uint8_t* dst, const uint8_t* src, uint32_t bytes, uint64_t* bar) {
  wp_mbar_expect_tx(bar, bytes);
  for (uint32_t off = 0; off < bytes; off += 32768u) wp_bulk_g2s(dst + off, src + off, min(32768u, bytes - off), bar);
}

__device__ __forceinline__ bool em_better(float v, int ix, float bv, int bi) { return v > bv || (v == bv && ix < bi); }

__global__ void __launch_bounds__(EM_THREADS, 1) edge_fwd_mma_kernel(ModelDev M, StepArgs a, const uint8_t* __restrict__ frags) {
  STRIVE_PDL_TRIGGER();
  STRIVE_PDL_WAIT();
  extern __shared__ __align__(128) uint8_t esm[];
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) em_load_packs(esm, frags + EM_F3_OFF, EM_FWD_SMEM, &bar);
  bool loaded = false;
  const int NA = a.NA;
  for (int i = blockIdx.x * EM_WARPS + warp; i < NA; i += gridDim.x * EM_WARPS) {
    EdgeCtx c;
    em_agent_ctx(a, i, c);
    float bval[8][2];
    int bidx[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) { bval[nt][0] = bval[nt][1] = -INFINITY; bidx[nt][0] = bidx[nt][1] = 255; }
    const int ntiles = (c.ne + 15) >> 4;
    for (int q = 0; q < ntiles; q++) {
      EdgeTile tl;
      float h[16][4];
      edge_tile_h1pre(M, a, c, q, g, t, tl, h);
      frag_ln_relu(h, M.seg[S_E_LN1_G], M.seg[S_E_LN1_B], t);
      if (!loaded) { wp_wait(&bar, 0); loaded = true; }
      float acc1[16][4];
      frag_bias<16>(acc1, M.seg[S_E3_B], t);
      frag_gemm<16, 16, false>(h, acc1, esm + EM_F3_OFF, lane);
      frag_ln_relu(acc1, M.seg[S_E_LN4_G], M.seg[S_E_LN4_B], t);
      float acc2[8][4];
      frag_bias<8>(acc2, M.seg[S_E6_B], t);
      frag_gemm<16, 8, false>(acc1, acc2, esm + EM_F6_OFF - EM_F3_OFF, lane);
      // running arg-max over the rows this lane holds (smaller source index wins ties; NaN never wins, as `m > best` in the SIMT
      // kernel); the reduction across the 8 row groups happens once per agent, after its last tile
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
#pragma unroll
          for (int r = 0; r < 2; r++) {
            const float m = acc2[nt][2 * r + e];
            if (tl.valid[r] && m == m && em_better(m, tl.lj[r], bval[nt][e], bidx[nt][e])) { bval[nt][e] = m; bidx[nt][e] = tl.lj[r]; }
          }
        }
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bval[nt][e], o);
          const int oi = __shfl_xor_sync(0xffffffffu, bidx[nt][e], o);
          if (em_better(ov, oi, bval[nt][e], bidx[nt][e])) { bval[nt][e] = ov; bidx[nt][e] = oi; }
        }
      }
    }
    if (g == 0) {
      float* ag = a.tp.aggr + ((size_t)a.t * NA + i) * 64;
      uint8_t* ar = a.tp.arg + ((size_t)a.t * NA + i) * 64;
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        const float v0 = bidx[nt][0] == 255 ? 0.f : bval[nt][0], v1 = bidx[nt][1] == 255 ? 0.f : bval[nt][1];
        *reinterpret_cast<float2*>(ag + 8 * nt + 2 * t) = make_float2(v0, v1);
        *reinterpret_cast<uchar2*>(ar + 8 * nt + 2 * t) = make_uchar2((unsigned char)bidx[nt][0], (unsigned char)bidx[nt][1]);
      }
    }
  }
  if (!loaded) wp_wait(&bar, 0);     // never leave with the bulk copies in flight
}

__global__ void __launch_bounds__(EM_THREADS, 1) edge_bwd_mma_kernel(ModelDev M, StepArgs a, const uint8_t* __restrict__ frags) {
  STRIVE_PDL_TRIGGER();
  STRIVE_PDL_WAIT();
  extern __shared__ __align__(128) uint8_t esm[];
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) em_load_packs(esm, frags + EM_B3T_OFF, EM_BWD_SMEM, &bar);
  bool loaded = false;
  const int NA = a.NA;
  const uint8_t* pk3t = esm;
  const uint8_t* pk6n = esm + (EM_B6N_OFF - EM_B3T_OFF);
  const uint8_t* pk3n = esm + (EM_B3N_OFF - EM_B3T_OFF);
  const float* Wr = M.seg[S_E0_T_REL];
  for (int i = blockIdx.x * EM_WARPS + warp; i < NA; i += gridDim.x * EM_WARPS) {
    EdgeCtx c;
    em_agent_ctx(a, i, c);
    float dagg[8][2];
    int am[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      const float2 d2 = *reinterpret_cast<const float2*>(a.tp.d_aggr + (size_t)i * 64 + 8 * nt + 2 * t);
      const uchar2 m2 = *reinterpret_cast<const uchar2*>(a.tp.arg + ((size_t)a.t * NA + i) * 64 + 8 * nt + 2 * t);
      dagg[nt][0] = d2.x; dagg[nt][1] = d2.y;
      am[nt][0] = m2.x; am[nt][1] = m2.y;
    }
    float dPacc[16][2];
#pragma unroll
    for (int j = 0; j < 16; j++) dPacc[j][0] = dPacc[j][1] = 0.f;
    float dposi[4] = {0.f, 0.f, 0.f, 0.f};
    const int ntiles = (c.ne + 15) >> 4;
    for (int q = 0; q < ntiles; q++) {
      EdgeTile tl;
      float dh[16][4];
      {
        float acc1[16][4];
        {
          float h[16][4];
          edge_tile_h1pre(M, a, c, q, g, t, tl, h);
          frag_ln_relu(h, M.seg[S_E_LN1_G], M.seg[S_E_LN1_B], t);
          if (!loaded) { wp_wait(&bar, 0); loaded = true; }
          frag_bias<16>(acc1, M.seg[S_E3_B], t);
          frag_gemm<16, 16, true>(h, acc1, pk3t, lane);            // a2 = pre-LN activations of layer 2
        }
        // d_m routed to the arg-max edges (interaction_net.py aggr='max'), then back through the last Linear
        float dm[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int r = e >> 1, cc = e & 1;
            dm[nt][e] = (tl.valid[r] && am[nt][cc] == tl.lj[r]) ? dagg[nt][cc] : 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < 16; j++) dh[j][0] = dh[j][1] = dh[j][2] = dh[j][3] = 0.f;
        frag_gemm<8, 16, true>(dm, dh, pk6n, lane);
        frag_ln_relu_bwd(dh, acc1, M.seg[S_E_LN4_G], M.seg[S_E_LN4_B], t);     // dh = d a2
      }
      float d1[16][4];
#pragma unroll
      for (int j = 0; j < 16; j++) d1[j][0] = d1[j][1] = d1[j][2] = d1[j][3] = 0.f;
      frag_gemm<16, 16, true>(dh, d1, pk3n, lane);                              // d h1
      {
        float h[16][4];
        EdgeTile t2;
        edge_tile_h1pre(M, a, c, q, g, t, t2, h);                               // recomputed instead of kept: 64 registers
        frag_ln_relu_bwd(d1, h, M.seg[S_E_LN1_G], M.seg[S_E_LN1_B], t);         // d1 = d h1pre = d(P_i + Q_j + W_rel rel)
      }
      // rows that are padding have dm = 0, hence exact zeros all the way down
      float drel[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int j = 0; j < 16; j++) {
        const int col = 8 * j + 2 * t;
        dPacc[j][0] += d1[j][0] + d1[j][2];
        dPacc[j][1] += d1[j][1] + d1[j][3];
#pragma unroll
        for (int d = 0; d < 4; d++) {
          const float2 w2 = __ldg(reinterpret_cast<const float2*>(Wr + d * 128 + col));
          drel[0][d] = fmaf(d1[j][0], w2.x, fmaf(d1[j][1], w2.y, drel[0][d]));
          drel[1][d] = fmaf(d1[j][2], w2.x, fmaf(d1[j][3], w2.y, drel[1][d]));
        }
#pragma unroll
        for (int r = 0; r < 2; r++) {
          if (tl.valid[r]) {
            float* dq = a.tp.dQ + (size_t)tl.j[r] * 128 + col;
            atomicAdd(dq, d1[j][2 * r]);
            atomicAdd(dq + 1, d1[j][2 * r + 1]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 2; r++) {
#pragma unroll
        for (int d = 0; d < 4; d++) drel[r][d] = quad_sum(drel[r][d]);
        if (tl.valid[r] && t == 0) {
          float relchk[4];
          t2f_fwd(c.pos_i, tl.pj[r], relchk);
#pragma unroll
          for (int d = 0; d < 4; d++)
            if (isnan(relchk[d])) drel[r][d] = 0.f;
          float dpj[4] = {0.f, 0.f, 0.f, 0.f};
          t2f_bwd(c.pos_i, tl.pj[r], drel[r], dposi, dpj);
#pragma unroll
          for (int k = 0; k < 4; k++) atomicAdd(a.tp.g_pos + (size_t)tl.j[r] * 4 + k, dpj[k]);
        }
      }
    }
    // column sums over the rows held by the 8 row groups, position adjoint of the target agent
#pragma unroll
    for (int j = 0; j < 16; j++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        float v = dPacc[j][e];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        dPacc[j][e] = v;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float v = dposi[k];                      // non-zero in lanes t == 0 only
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      dposi[k] = v;
    }
    if (g == 0) {
#pragma unroll
      for (int j = 0; j < 16; j++) *reinterpret_cast<float2*>(a.tp.dP + (size_t)i * 128 + 8 * j + 2 * t) = make_float2(dPacc[j][0], dPacc[j][1]);
    }
    if (lane < 4) atomicAdd(a.tp.g_pos + (size_t)i * 4 + lane, lane == 0 ? dposi[0] : lane == 1 ? dposi[1] : lane == 2 ? dposi[2] : dposi[3]);
  }
  if (!loaded) wp_wait(&bar, 0);
}
