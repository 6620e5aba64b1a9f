// Decoder rollout (forward + BPTT w.r.t. z) of the STRIVE traffic prior.
//
// Restates reference src/models/traffic_model.py:589-704 (autoregressive_decoder), :714-733 (sim_traj),
// src/models/interaction_net.py:52-218 (SceneInteractionNet / AgentInteractionConv with max aggregation),
// src/models/common.py:8-67 (MLP, car_dynamics), nn.GRU(4,64,3) one step, utils/transforms.py:78-139.
//
// Per rollout step t (forward):   node_fwd -> edge_fwd -> post_fwd -> gru_fwd -> [map encoder, mapenc.cu]
// Per rollout step t (backward):  gru_bwd -> post_bwd -> edge_bwd -> node_bwd      (t = FT-1 .. 0)
//
// Work decomposition: one warp owns R rows (agents, or edges of one target agent); a row's activations live
// in per-warp shared memory (broadcast reads), each lane owns OUT/32 output columns, weights are read
// coalesced from global ([in][out] layout) and are L1/L2 resident (<1 MB for the whole decoder).
// The first edge-MLP layer is factorised:  W1 [x_i|x_j|sem_i|sem_j|rel] = P_i + Q_j + W_rel rel  (node terms
// P,Q computed once per node), which removes 40 % of the per-edge MACs.
//
// Backward recomputes edge/node activations from the tape (node-level tensors only) instead of storing
// per-edge activations (E x 320 floats per step would be 4 GB at BASELINE config 5).
#define STRIVE_PDL_CLASS 1   // bit of strive_set_pdl() that enables programmatic dependent launch for this file's kernels
#include "common.cuh"
#include "wpipe.cuh"
#include <cuda_bf16.h>

// ------------------------------------------------------------------------------------------------------
// tape layout (floats unless noted), all [t][agent][...]
// ------------------------------------------------------------------------------------------------------
struct Tape {
  float* pastfeat;  // [FT][NA][64]   past_feat input of step t
  float* mapfeat;   // [FT][NA][64]   map_feat input of step t
  float* mem;       // [FT][NA][3][64] GRU hidden at the start of step t
  float* prev;      // [FT][NA][6]    prev_state input of step t (normalised)
  float* pos;       // [FT][NA][4]    pos used for edge transforms at step t (normalised)
  float* loc;       // [FT][NA][4]    local-frame delta fed to the GRU at step t
  float* x;         // [FT][NA][64]   mlp_in output
  float* P;         // [FT][NA][128]
  float* Q;         // [FT][NA][128]
  float* aggr;      // [FT][NA][64]
  uint8_t* arg;     // [FT][NA][64]   local index (within scene) of the arg-max source, 255 = none
  float* z;         // [NA][32]       copy of the latent the forward pass ran on (node recompute in backward)
  // backward carries / scratch, [NA][...]
  float* g_prev;    // [NA][6]
  float* g_pos;     // [NA][4]
  float* g_pf;      // [NA][64]
  float* g_mem;     // [NA][3][64]
  float* d_loc;     // [NA][4]
  float* d_xupd;    // [NA][64]
  float* d_aggr;    // [NA][64]
  float* dP;        // [NA][128]
  float* dQ;        // [NA][128]
  float* pose;      // [NA][4]  unnormalised crop pose for the map encoder
  int32_t* map_of;  // [NA]
  int32_t* et_tiles;   // [NA][2] edge-tile table of the tcgen05 edge kernels (edge_tc.cuh): (scene, first local target)
  int32_t* et_ntiles;  // [1]
  void* mapenc_ws;  // map encoder workspace
  int64_t mapenc_ws_bytes;
  float* alt[9];    // second set of the nine backward carries above: the sweep strive_decode_bwd_pair runs on its side stream
};

// switches a tape view to the second carry set (forward tensors stay shared)
static inline void tape_use_alt(Tape& tp) {
  tp.g_prev = tp.alt[0]; tp.g_pos = tp.alt[1]; tp.g_pf = tp.alt[2]; tp.g_mem = tp.alt[3]; tp.d_loc = tp.alt[4];
  tp.d_xupd = tp.alt[5]; tp.d_aggr = tp.alt[6]; tp.dP = tp.alt[7]; tp.dQ = tp.alt[8];
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static int64_t tape_carve(Tape* tp, char* base, int NA, int FT) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align256(bytes);
    return p;
  };
  const size_t n = (size_t)NA, T = (size_t)FT;
  tp->pastfeat = (float*)take(T * n * 64 * 4);
  tp->mapfeat = (float*)take(T * n * 64 * 4);
  tp->mem = (float*)take(T * n * 192 * 4);
  tp->prev = (float*)take(T * n * 6 * 4);
  tp->pos = (float*)take(T * n * 4 * 4);
  tp->loc = (float*)take(T * n * 4 * 4);
  tp->x = (float*)take(T * n * 64 * 4);
  tp->P = (float*)take(T * n * 128 * 4);
  tp->Q = (float*)take(T * n * 128 * 4);
  tp->aggr = (float*)take(T * n * 64 * 4);
  tp->arg = (uint8_t*)take(T * n * 64);
  tp->z = (float*)take(n * 32 * 4);
  tp->g_prev = (float*)take(n * 6 * 4);
  tp->g_pos = (float*)take(n * 4 * 4);
  tp->g_pf = (float*)take(n * 64 * 4);
  tp->g_mem = (float*)take(n * 192 * 4);
  tp->d_loc = (float*)take(n * 4 * 4);
  tp->d_xupd = (float*)take(n * 64 * 4);
  tp->d_aggr = (float*)take(n * 64 * 4);
  tp->dP = (float*)take(n * 128 * 4);
  tp->dQ = (float*)take(n * 128 * 4);
  tp->pose = (float*)take(n * 4 * 4);
  tp->map_of = (int32_t*)take(n * 4);
  tp->et_tiles = (int32_t*)take(n * 2 * 4);
  tp->et_ntiles = (int32_t*)take(256);
  {
    const size_t w[9] = {6, 4, 64, 192, 4, 64, 64, 128, 128};
    for (int i = 0; i < 9; i++) tp->alt[i] = (float*)take(n * w[i] * 4);
  }
  tp->mapenc_ws_bytes = strive_mapenc_workspace_bytes(NA);
  tp->mapenc_ws = (void*)take((size_t)tp->mapenc_ws_bytes);
  return (int64_t)off;
}

extern "C" int64_t strive_decode_tape_bytes(int32_t num_agents, int32_t ft) {
  Tape tp;
  return tape_carve(&tp, nullptr, num_agents, ft);
}

// ------------------------------------------------------------------------------------------------------
// kernel parameter blocks
// ------------------------------------------------------------------------------------------------------
struct ModelDev {
  const float* seg[S_COUNT];
  int nc, in0_rows, u0_rows;
};

struct StepArgs {
  int NA, t, FT, NC;
  const int32_t* ptr;
  const int32_t* scene_of;
  const float* lw;
  const float* sem;
  const float* z;
  const float* ext;     // (S,FT,4) or null
  float* traj;          // (NA,FT,4)
  const float* d_traj;  // (NA,FT,4)
  float* d_z;           // (NA,32)
  Tape tp;
};

// node-level kernels: NODE_WARPS consumer warps x NODE_R rows each + one producer warp that streams the weights through
// shared memory (wpipe.cuh).  History, measured on B200 at NA = 2048 (scripts/prof_step.py) with per-warp weight reads from
// L1/L2: R=4/unroll 2: 9.5 ms of node-level kernels per iteration, R=2/unroll 4: 7.8 ms (L2-latency bound, 0.3 % of the FMA peak).
#ifndef NODE_R
#define NODE_R 4
#endif
#ifndef NODE_WARPS
#define NODE_WARPS 4
#endif
#define NODE_THREADS ((NODE_WARPS + 1) * 32)
#define EDGE_R 8
#define EDGE_WARPS 4
#define LDA 172   // >= 168, multiple of 4
#define LDH 132   // 128 + 4

// ------------------------------------------------------------------------------------------------------
// node phase: f -> mlp_in -> x, P, Q           (interaction_net.py:61 ; factorised first edge layer)
// ------------------------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void stage_node_feat(float* bufA, const StepArgs& a, const int (&row)[R], int lane, int in0_rows) {
  // every global load of the R rows is issued before the first shared-memory store waits on one (a load -> store loop
  // serialises 6 L2 round trips per row: 22 % of node_fwd's cycles in the ncu stall profile)
  const int NA = a.NA, NC = a.NC;
  const float* pf = a.tp.pastfeat + (size_t)a.t * NA * 64;
  const float* mf = a.tp.mapfeat + (size_t)a.t * NA * 64;
  float4 head[R];
  float tail[R][2];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int ag = row[r];
    head[r] = __ldg(reinterpret_cast<const float4*>((lane < 16 ? pf : mf) + (size_t)ag * 64 + (lane & 15) * 4));
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int k = 128 + h * 32 + lane;
      float v = 0.f;
      if (k < 128 + NC) v = __ldg(a.sem + (size_t)ag * NC + k - 128);
      else if (k < 128 + NC + ZDIM) v = __ldg(a.z + (size_t)ag * ZDIM + k - 128 - NC);
      else if (k < 128 + NC + ZDIM + 2) v = __ldg(a.lw + (size_t)ag * 2 + k - 128 - NC - ZDIM);
      tail[r][h] = v;
    }
  }
#pragma unroll
  for (int r = 0; r < R; r++) {
    *reinterpret_cast<float4*>(bufA + r * LDA + lane * 4) = head[r];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int k = 128 + h * 32 + lane;
      if (k < in0_rows) bufA[r * LDA + k] = tail[r][h];
    }
  }
}

// forward of mlp_in on R staged rows. Leaves h2 (post LN/ReLU of layer 2) in bufA (ld LDA) and, if keep, the
// pre-LN activations a1 in pre1 and a2 in pre2 (ld LDH). Returns x in xacc.
template <int R>
__device__ __forceinline__ void mlp_in_fwd(const ModelDev& M, WPipe& wp, float* bufA, float* bufB, float* pre1, float* pre2,
                                           float (&xacc)[R][2], int lane) {
  float acc[R][4];
  init_bias<4, R>(acc, M.seg[S_IN0_B], lane);
  pipe_gemm<128, R>(wp, M.in0_rows, bufA, LDA, acc, lane);
  __syncwarp();
  ln_relu_store<R>(acc, M.seg[S_IN_LN1_G], M.seg[S_IN_LN1_B], bufB, LDH, pre1, LDH, lane);
  __syncwarp();
  init_bias<4, R>(acc, M.seg[S_IN3_B], lane);
  pipe_gemm<128, R>(wp, 128, bufB, LDH, acc, lane);
  __syncwarp();
  ln_relu_store<R>(acc, M.seg[S_IN_LN4_G], M.seg[S_IN_LN4_B], bufA, LDA, pre2, LDH, lane);
  __syncwarp();
  init_bias<2, R>(xacc, M.seg[S_IN6_B], lane);
  pipe_gemm<64, R>(wp, 128, bufA, LDA, xacc, lane);
  __syncwarp();
}

// small parameter vectors -> L1 (called by every consumer warp while its inputs are in flight)
__device__ __forceinline__ void prefetch_mlp_in_params(const ModelDev& M, int lane) {
  wp_prefetch_l1(M.seg[S_IN0_B], 128, lane); wp_prefetch_l1(M.seg[S_IN_LN1_G], 128, lane); wp_prefetch_l1(M.seg[S_IN_LN1_B], 128, lane);
  wp_prefetch_l1(M.seg[S_IN3_B], 128, lane); wp_prefetch_l1(M.seg[S_IN_LN4_G], 128, lane); wp_prefetch_l1(M.seg[S_IN_LN4_B], 128, lane);
  wp_prefetch_l1(M.seg[S_IN6_B], 64, lane);
}
__device__ __forceinline__ void prefetch_post_params(const ModelDev& M, int lane) {
  wp_prefetch_l1(M.seg[S_U0_B], 128, lane); wp_prefetch_l1(M.seg[S_U_LN1_G], 128, lane); wp_prefetch_l1(M.seg[S_U_LN1_B], 128, lane);
  wp_prefetch_l1(M.seg[S_U3_B], 64, lane); wp_prefetch_l1(M.seg[S_O0_B], 128, lane); wp_prefetch_l1(M.seg[S_O_LN1_G], 128, lane);
  wp_prefetch_l1(M.seg[S_O_LN1_B], 128, lane); wp_prefetch_l1(M.seg[S_O3_B], 128, lane); wp_prefetch_l1(M.seg[S_O_LN4_G], 128, lane);
  wp_prefetch_l1(M.seg[S_O_LN4_B], 128, lane); wp_prefetch_l1(M.seg[S_O6_N], 256, lane); wp_prefetch_l1(M.seg[S_O6_B], 2, lane);
}
__device__ __forceinline__ void prefetch_gru_params(const ModelDev& M, int lane) {
#pragma unroll
  for (int l = 0; l < 3; l++) {
    wp_prefetch_l1(M.seg[S_GBI0 + l * 6], 192, lane);
    wp_prefetch_l1(M.seg[S_GBH0 + l * 6], 192, lane);
  }
  wp_prefetch_l1(M.seg[S_GI_T0], 4 * 192, lane);
  wp_prefetch_l1(M.seg[S_GI_N0], 4 * 192, lane);
}

// producer-side schedules: the matrices each kernel's consumers multiply by, in consumption order
__device__ __forceinline__ void produce_mlp_in(const ModelDev& M, WPipe& wp) {
  wp_produce(wp, M.seg[S_IN0_T], M.in0_rows, 128);
  wp_produce(wp, M.seg[S_IN3_T], 128, 128);
  wp_produce(wp, M.seg[S_IN6_T], 128, 64);
}

// rows of this consumer warp (clamped: warps past the end of the batch still step through the weight ring)
template <int R>
__device__ __forceinline__ void node_rows(int warp, int NA, int& base, int (&row)[R], bool (&valid)[R]) {
  base = (blockIdx.x * NODE_WARPS + warp) * R;
#pragma unroll
  for (int r = 0; r < R; r++) {
    valid[r] = (base + r) < NA;
    row[r] = valid[r] ? base + r : NA - 1;
  }
}

__global__ void __launch_bounds__(NODE_THREADS, 2) node_fwd_kernel(ModelDev M, StepArgs a) {
  STRIVE_PDL_TRIGGER();
  STRIVE_PDL_WAIT();      // first instructions of the kernel: nothing (not even a hoisted read-only load) can precede them
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WPipe wp = wp_init(smem, NODE_WARPS);
  if (warp == NODE_WARPS) {
    if (lane == 0) {
      produce_mlp_in(M, wp);
      wp_produce(wp, M.seg[S_E0_T_XI], 64, 128);
      wp_produce(wp, M.seg[S_E0_T_XJ], 64, 128);
    }
    return;
  }
  float* bufA = smem + WP_SMEM_FLOATS + warp * (NODE_R * (LDA + LDH));
  float* bufB = bufA + NODE_R * LDA;
  const int NA = a.NA;
  int row[NODE_R], base;
  bool valid[NODE_R];
  node_rows<NODE_R>(warp, NA, base, row, valid);
  prefetch_mlp_in_params(M, lane);
  wp_prefetch_l1(M.seg[S_E0_B], 128, lane);
  wp_prefetch_l1(M.seg[S_E0_T_SEMI], a.NC * 128, lane);
  wp_prefetch_l1(M.seg[S_E0_T_SEMJ], a.NC * 128, lane);
  stage_node_feat<NODE_R>(bufA, a, row, lane, M.in0_rows);
  __syncwarp();
  float xacc[NODE_R][2];
  mlp_in_fwd<NODE_R>(M, wp, bufA, bufB, nullptr, nullptr, xacc, lane);
  // x -> tape + shared (bufB as [R][LDH], first 64 cols)
  float* xg = a.tp.x + (size_t)a.t * NA * 64;
#pragma unroll
  for (int r = 0; r < NODE_R; r++) {
    stvec<2>(bufB + r * LDH + lane * 2, xacc[r]);
    if (valid[r]) stvec<2>(xg + (size_t)row[r] * 64 + lane * 2, xacc[r]);
  }
  __syncwarp();
  // P = W_xi x + W_semi sem + b ; Q = W_xj x + W_semj sem
  float acc[NODE_R][4];
  init_bias<4, NODE_R>(acc, M.seg[S_E0_B], lane);
  pipe_gemm<128, NODE_R>(wp, 64, bufB, LDH, acc, lane);
  for (int c = 0; c < a.NC; c++) {
    float w[4];
    ldvec<4>(w, M.seg[S_E0_T_SEMI] + c * 128 + lane * 4);
#pragma unroll
    for (int r = 0; r < NODE_R; r++) {
      const float s = a.sem[(size_t)row[r] * a.NC + c];
#pragma unroll
      for (int v = 0; v < 4; v++) acc[r][v] = fmaf(s, w[v], acc[r][v]);
    }
  }
  float* Pg = a.tp.P + (size_t)a.t * NA * 128;
#pragma unroll
  for (int r = 0; r < NODE_R; r++)
    if (valid[r]) stvec<4>(Pg + (size_t)row[r] * 128 + lane * 4, acc[r]);
  init_zero<4, NODE_R>(acc);
  pipe_gemm<128, NODE_R>(wp, 64, bufB, LDH, acc, lane);
  for (int c = 0; c < a.NC; c++) {
    float w[4];
    ldvec<4>(w, M.seg[S_E0_T_SEMJ] + c * 128 + lane * 4);
#pragma unroll
    for (int r = 0; r < NODE_R; r++) {
      const float s = a.sem[(size_t)row[r] * a.NC + c];
#pragma unroll
      for (int v = 0; v < 4; v++) acc[r][v] = fmaf(s, w[v], acc[r][v]);
    }
  }
  float* Qg = a.tp.Q + (size_t)a.t * NA * 128;
#pragma unroll
  for (int r = 0; r < NODE_R; r++)
    if (valid[r]) stvec<4>(Qg + (size_t)row[r] * 128 + lane * 4, acc[r]);
}

// ------------------------------------------------------------------------------------------------------
// edge phase: one CTA per target agent i; rows = incoming edges (j -> i), max-aggregate with arg-max
// (interaction_net.py:139-184 message(), aggr='max' at :92, zeros for edge-less nodes :187-188)
// ------------------------------------------------------------------------------------------------------
struct EdgeCtx {
  int i, p0, n, ne, li;
  float pos_i[4];
};

// first edge layer for R rows of one chunk: h1pre = P_i + Q_j + W_rel rel  (lane cols)
template <int R>
__device__ __forceinline__ void edge_layer0(const StepArgs& a, const EdgeCtx& c, int chunk, const float (&Pi)[4],
                                            const float (&Wrel)[4][4], float (&acc)[R][4], int (&lj)[R], bool (&valid)[R],
                                            float (&rel)[R][4], int lane) {
  const int NA = a.NA;
  const float* Qg = a.tp.Q + (size_t)a.t * NA * 128;
  const float* posg = a.tp.pos + (size_t)a.t * NA * 4;
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int e = chunk * R + r;
    valid[r] = e < c.ne;
    const int ee = valid[r] ? e : 0;
    lj[r] = ee + (ee >= c.li ? 1 : 0);
    const int j = c.p0 + lj[r];
    const float4 pj4 = __ldg(reinterpret_cast<const float4*>(posg + (size_t)j * 4));
    const float pj[4] = {pj4.x, pj4.y, pj4.z, pj4.w};
    t2f_fwd(c.pos_i, pj, rel[r]);
#pragma unroll
    for (int d = 0; d < 4; d++)
      if (isnan(rel[r][d])) rel[r][d] = 0.f;   // interaction_net.py:162
    float q[4];
    ldvec<4>(q, Qg + (size_t)j * 128 + lane * 4);
#pragma unroll
    for (int v = 0; v < 4; v++) {
      float h = Pi[v] + q[v];
#pragma unroll
      for (int d = 0; d < 4; d++) h = fmaf(rel[r][d], Wrel[d][v], h);
      acc[r][v] = h;
    }
  }
}

__device__ __forceinline__ bool edge_ctx_init(const StepArgs& a, EdgeCtx& c) {
  c.i = blockIdx.x;
  const int s = a.scene_of[c.i];
  c.p0 = a.ptr[s];
  c.n = a.ptr[s + 1] - c.p0;
  c.ne = c.n - 1;
  c.li = c.i - c.p0;
  const float4 p = __ldg(reinterpret_cast<const float4*>(a.tp.pos + ((size_t)a.t * a.NA + c.i) * 4));
  c.pos_i[0] = p.x; c.pos_i[1] = p.y; c.pos_i[2] = p.z; c.pos_i[3] = p.w;
  return true;
}

__global__ void __launch_bounds__(EDGE_WARPS * 32) edge_fwd_kernel(ModelDev M, StepArgs a) {
  STRIVE_PDL_TRIGGER();
  STRIVE_PDL_WAIT();
  extern __shared__ __align__(16) float smem[];
  __shared__ float red_val[EDGE_WARPS][64];
  __shared__ int red_idx[EDGE_WARPS][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* bufA = smem + warp * (2 * EDGE_R * LDH);
  float* bufB = bufA + EDGE_R * LDH;
  EdgeCtx c;
  edge_ctx_init(a, c);
  const int NA = a.NA;
  float Pi[4], Wrel[4][4];
  ldvec<4>(Pi, a.tp.P + ((size_t)a.t * NA + c.i) * 128 + lane * 4);
#pragma unroll
  for (int d = 0; d < 4; d++) ldvec<4>(Wrel[d], M.seg[S_E0_T_REL] + d * 128 + lane * 4);
  float best[2] = {-INFINITY, -INFINITY};
  int bidx[2] = {255, 255};
  const int nchunks = (c.ne + EDGE_R - 1) / EDGE_R;
  for (int chunk = warp; chunk < nchunks; chunk += EDGE_WARPS) {
    float acc[EDGE_R][4], rel[EDGE_R][4];
    int lj[EDGE_R];
    bool valid[EDGE_R];
    edge_layer0<EDGE_R>(a, c, chunk, Pi, Wrel, acc, lj, valid, rel, lane);
    __syncwarp();
    ln_relu_store<EDGE_R>(acc, M.seg[S_E_LN1_G], M.seg[S_E_LN1_B], bufA, LDH, nullptr, 0, lane);
    __syncwarp();
    init_bias<4, EDGE_R>(acc, M.seg[S_E3_B], lane);
    warp_gemm<128, EDGE_R>(M.seg[S_E3_T], 128, bufA, LDH, acc, lane);
    __syncwarp();
    ln_relu_store<EDGE_R>(acc, M.seg[S_E_LN4_G], M.seg[S_E_LN4_B], bufB, LDH, nullptr, 0, lane);
    __syncwarp();
    float m[EDGE_R][2];
    init_bias<2, EDGE_R>(m, M.seg[S_E6_B], lane);
    warp_gemm<64, EDGE_R>(M.seg[S_E6_T], 128, bufB, LDH, m, lane);
#pragma unroll
    for (int r = 0; r < EDGE_R; r++) {
      if (valid[r]) {
#pragma unroll
        for (int v = 0; v < 2; v++) {
          if (m[r][v] > best[v] || (m[r][v] == best[v] && lj[r] < bidx[v])) {
            best[v] = m[r][v];
            bidx[v] = lj[r];
          }
        }
      }
    }
    __syncwarp();
  }
  red_val[warp][lane * 2] = best[0];
  red_val[warp][lane * 2 + 1] = best[1];
  red_idx[warp][lane * 2] = bidx[0];
  red_idx[warp][lane * 2 + 1] = bidx[1];
  __syncthreads();
  if (threadIdx.x < 64) {
    const int ch = threadIdx.x;
    float bv = red_val[0][ch];
    int bi = red_idx[0][ch];
#pragma unroll
    for (int w = 1; w < EDGE_WARPS; w++) {
      const float v = red_val[w][ch];
      const int ix = red_idx[w][ch];
      if (ix != 255 && (bi == 255 || v > bv || (v == bv && ix < bi))) {
        bv = v;
        bi = ix;
      }
    }
    if (bi == 255) bv = 0.f;
    a.tp.aggr[((size_t)a.t * NA + c.i) * 64 + ch] = bv;
    a.tp.arg[((size_t)a.t * NA + c.i) * 64 + ch] = (uint8_t)bi;
  }
}

// ------------------------------------------------------------------------------------------------------
// post phase: update_mlp -> mlp_out -> bicycle -> local-frame delta   (interaction_net.py:186-218, :69;
// traffic_model.py:640-680; models/common.py:47-67)
// ------------------------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void stage_update_in(float* bufA, const StepArgs& a, const int (&row)[R], int lane, int u0_rows) {
  const int NA = a.NA, NC = a.NC;
  const float* xg = a.tp.x + (size_t)a.t * NA * 64;
  const float* ag = a.tp.aggr + (size_t)a.t * NA * 64;
  float4 head[R];
  float tail[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int agn = row[r];
    head[r] = __ldg(reinterpret_cast<const float4*>((lane < 16 ? xg : ag) + (size_t)agn * 64 + (lane & 15) * 4));
    tail[r] = (lane < NC) ? __ldg(a.sem + (size_t)agn * NC + lane) : 0.f;
  }
#pragma unroll
  for (int r = 0; r < R; r++) {
    *reinterpret_cast<float4*>(bufA + r * LDA + lane * 4) = head[r];
    if (128 + lane < u0_rows) bufA[r * LDA + 128 + lane] = tail[r];
  }
}

// update_mlp + mlp_out forward on R staged rows; returns o (2 per row, replicated in every lane).
// If keep: preU (a_u), pre1 (a_1), pre2 (a_2) pre-LN activations (ld LDH) are kept for the backward pass.
template <int R>
__device__ __forceinline__ void post_mlps_fwd(const ModelDev& M, WPipe& wp, float* bufA, float* bufB, float* preU, float* pre1,
                                              float* pre2, float (&o)[R][2], int lane) {
  float acc[R][4];
  init_bias<4, R>(acc, M.seg[S_U0_B], lane);
  pipe_gemm<128, R>(wp, M.u0_rows, bufA, LDA, acc, lane);
  __syncwarp();
  ln_relu_store<R>(acc, M.seg[S_U_LN1_G], M.seg[S_U_LN1_B], bufB, LDH, preU, LDH, lane);
  __syncwarp();
  float xu[R][2];
  init_bias<2, R>(xu, M.seg[S_U3_B], lane);
  pipe_gemm<64, R>(wp, 128, bufB, LDH, xu, lane);
  __syncwarp();
  store_rows<2, R>(xu, bufA, LDA, lane);
  __syncwarp();
  init_bias<4, R>(acc, M.seg[S_O0_B], lane);
  pipe_gemm<128, R>(wp, 64, bufA, LDA, acc, lane);
  __syncwarp();
  ln_relu_store<R>(acc, M.seg[S_O_LN1_G], M.seg[S_O_LN1_B], bufB, LDH, pre1, LDH, lane);
  __syncwarp();
  init_bias<4, R>(acc, M.seg[S_O3_B], lane);
  pipe_gemm<128, R>(wp, 128, bufB, LDH, acc, lane);
  __syncwarp();
  ln_relu_store<R>(acc, M.seg[S_O_LN4_G], M.seg[S_O_LN4_B], bufA, LDA, pre2, LDH, lane);
  __syncwarp();
  float w0[4], w1[4];
  ldvec<4>(w0, M.seg[S_O6_N] + lane * 4);
  ldvec<4>(w1, M.seg[S_O6_N] + 128 + lane * 4);
  const float b0 = __ldg(M.seg[S_O6_B]), b1 = __ldg(M.seg[S_O6_B] + 1);
#pragma unroll
  for (int r = 0; r < R; r++) {
    const float4 h = *reinterpret_cast<const float4*>(bufA + r * LDA + lane * 4);
    float p0 = h.x * w0[0] + h.y * w0[1] + h.z * w0[2] + h.w * w0[3];
    float p1 = h.x * w1[0] + h.y * w1[1] + h.z * w1[2] + h.w * w1[3];
    o[r][0] = warp_sum(p0) + b0;
    o[r][1] = warp_sum(p1) + b1;
  }
}

__device__ __forceinline__ void produce_post_mlps(const ModelDev& M, WPipe& wp) {
  wp_produce(wp, M.seg[S_U0_T], M.u0_rows, 128);
  wp_produce(wp, M.seg[S_U3_T], 128, 64);
  wp_produce(wp, M.seg[S_O0_T], 64, 128);
  wp_produce(wp, M.seg[S_O3_T], 128, 128);
}

struct BikeFwd {
  float un[6];        // unnormalised prev state
  float h, newh, sn, cs, news, newhdot, pre_s, pre_hd, len;
};

// sim_traj one step (traffic_model.py:714-733 + common.py:47-67); cur = normalised new state (6)
__device__ __forceinline__ void bicycle_fwd(const float prev[6], float o0, float o1, float lw_l_norm, float cur[6], BikeFwd& b) {
  const float acc = o0 * A_STD + A_MEAN;          // traffic_model.py:645
  const float ddh = o1 * DDH_STD + DDH_MEAN;      // :646
#pragma unroll
  for (int k = 0; k < 6; k++) b.un[k] = prev[k] * kStateStd[k] + kStateMean[k];
  b.len = lw_l_norm * ATT_STD_L + ATT_MEAN_L;     // :601
  b.h = atan2f(b.un[3], b.un[2]);
  b.pre_hd = b.un[5] + ddh * BIKE_DT;
  b.newhdot = fminf(fmaxf(b.pre_hd, -BIKE_MAXHDOT), BIKE_MAXHDOT);
  b.newh = b.h + BIKE_DT * fabsf(b.un[4]) / b.len * b.newhdot;
  b.pre_s = b.un[4] + acc * BIKE_DT;
  b.news = fminf(fmaxf(b.pre_s, 0.0f), BIKE_MAXS);
  b.sn = sinf(b.newh);
  b.cs = cosf(b.newh);
  const float out[6] = {b.un[0] + b.news * b.cs * BIKE_DT, b.un[1] + b.news * b.sn * BIKE_DT, b.cs, b.sn, b.news, b.newhdot};
#pragma unroll
  for (int k = 0; k < 6; k++) cur[k] = (out[k] - kStateMean[k]) / kStateStd[k];
}

__global__ void __launch_bounds__(NODE_THREADS, 2) post_fwd_kernel(ModelDev M, StepArgs a) {
  STRIVE_PDL_TRIGGER();
  STRIVE_PDL_WAIT();      // first instructions of the kernel: nothing (not even a hoisted read-only load) can precede them
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WPipe wp = wp_init(smem, NODE_WARPS);
  if (warp == NODE_WARPS) {
    if (lane == 0) produce_post_mlps(M, wp);
    return;
  }
  float* bufA = smem + WP_SMEM_FLOATS + warp * (NODE_R * (LDA + LDH));
  float* bufB = bufA + NODE_R * LDA;
  const int NA = a.NA, t = a.t, FT = a.FT;
  int row[NODE_R], base;
  bool valid[NODE_R];
  node_rows<NODE_R>(warp, NA, base, row, valid);
  prefetch_post_params(M, lane);
  stage_update_in<NODE_R>(bufA, a, row, lane, M.u0_rows);
  __syncwarp();
  float o[NODE_R][2];
  post_mlps_fwd<NODE_R>(M, wp, bufA, bufB, nullptr, nullptr, nullptr, o, lane);
  float my_o0 = 0.f, my_o1 = 0.f;
#pragma unroll
  for (int r = 0; r < NODE_R; r++)
    if (lane == r) { my_o0 = o[r][0]; my_o1 = o[r][1]; }
  if (lane < NODE_R && base + lane < NA) {
    const int ag = base + lane;
    const float* pv = a.tp.prev + ((size_t)t * NA + ag) * 6;
    float prev[6], cur[6];
#pragma unroll
    for (int k = 0; k < 6; k++) prev[k] = pv[k];
    BikeFwd b;
    bicycle_fwd(prev, my_o0, my_o1, a.lw[(size_t)ag * 2], cur, b);
    float* tr = a.traj + ((size_t)ag * FT + t) * 4;
    *reinterpret_cast<float4*>(tr) = make_float4(cur[0], cur[1], cur[2], cur[3]);   // :665
    float glob[4] = {cur[0], cur[1], cur[2], cur[3]};
    const int s = a.scene_of[ag];
    if (a.ext != nullptr && ag == a.ptr[s]) {                                        // :667-675
      const float* e = a.ext + ((size_t)s * FT + t) * 4;
#pragma unroll
      for (int k = 0; k < 4; k++) glob[k] = e[k];
    }
    float loc[4];
    t2f_fwd(prev, glob, loc);                                                        // :654 / :674
    *reinterpret_cast<float4*>(a.tp.loc + ((size_t)t * NA + ag) * 4) = make_float4(loc[0], loc[1], loc[2], loc[3]);
    if (t + 1 < FT) {
      float* pn = a.tp.prev + ((size_t)(t + 1) * NA + ag) * 6;                       // :680
#pragma unroll
      for (int k = 0; k < 6; k++) pn[k] = cur[k];
      *reinterpret_cast<float4*>(a.tp.pos + ((size_t)(t + 1) * NA + ag) * 4) = make_float4(glob[0], glob[1], glob[2], glob[3]);  // :698
      // crop pose: unnormalise (encode_map -> normalize_scene_graph(unnorm=True)), separate mul and add as torch does
      float4 pu;
      pu.x = __fadd_rn(__fmul_rn(glob[0], kStateStd[0]), kStateMean[0]);
      pu.y = __fadd_rn(__fmul_rn(glob[1], kStateStd[1]), kStateMean[1]);
      pu.z = __fadd_rn(__fmul_rn(glob[2], kStateStd[2]), kStateMean[2]);
      pu.w = __fadd_rn(__fmul_rn(glob[3], kStateStd[3]), kStateMean[3]);
      *reinterpret_cast<float4*>(a.tp.pose + (size_t)ag * 4) = pu;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// GRU memory: one step of nn.GRU(4,64,3) (traffic_model.py:152-156, 686-688)
// ------------------------------------------------------------------------------------------------------
#define GRU_R NODE_R
// per-warp shared: xin [R][68], h [3][R][68], gate stash for backward [3][R][4][64]
#define LDG 68

// hidden state [3][64] and local-frame delta [4] of R rows -> shared (all loads in flight before the first store)
template <int R>
__device__ __forceinline__ void stage_gru_in(float* hbuf, float* locs, const float* __restrict__ memg, const float* __restrict__ locg,
                                             const int (&row)[R], int lane) {
  float4 h0[R], h1[R];
  float lc[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const float4* src = reinterpret_cast<const float4*>(memg + (size_t)row[r] * 192);
    h0[r] = __ldg(src + lane);
    h1[r] = (lane < 16) ? __ldg(src + 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    lc[r] = (lane < 4) ? __ldg(locg + (size_t)row[r] * 4 + lane) : 0.f;
  }
#pragma unroll
  for (int r = 0; r < R; r++) {
    // float4 index q = k/4 of mem[row][k]: layer q>>4, offset (q&15)*4
    *reinterpret_cast<float4*>(hbuf + ((lane >> 4) * R + r) * LDG + (lane & 15) * 4) = h0[r];
    if (lane < 16) *reinterpret_cast<float4*>(hbuf + (2 * R + r) * LDG + lane * 4) = h1[r];
    if (lane < 4) locs[r * 4 + lane] = lc[r];
  }
}

template <int R>
__device__ __forceinline__ void gru_layer_fwd(const ModelDev& M, WPipe& wp, int l, const float* xin, const float* loc /*[R][4] if l==0*/,
                                              const float* hprev, float (&hnew)[R][2], float* stash /*[R][4][64] or null*/,
                                              int lane) {
  const int so = l * 6;
  float gi[R][6], gh[R][6];
  // biases
  {
    const float* bi = M.seg[S_GBI0 + so];
    const float* bh = M.seg[S_GBH0 + so];
#pragma unroll
    for (int g = 0; g < 3; g++) {
      const float2 a = __ldg(reinterpret_cast<const float2*>(bi + g * 64 + lane * 2));
      const float2 b = __ldg(reinterpret_cast<const float2*>(bh + g * 64 + lane * 2));
#pragma unroll
      for (int r = 0; r < R; r++) {
        gi[r][g * 2] = a.x; gi[r][g * 2 + 1] = a.y;
        gh[r][g * 2] = b.x; gh[r][g * 2 + 1] = b.y;
      }
    }
  }
  if (l == 0) {
    const float* W = M.seg[S_GI_T0];   // [4][192]
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
      for (int g = 0; g < 3; g++) {
        const float2 w = __ldg(reinterpret_cast<const float2*>(W + k * 192 + g * 64 + lane * 2));
#pragma unroll
        for (int r = 0; r < R; r++) {
          const float xv = loc[r * 4 + k];
          gi[r][g * 2] = fmaf(xv, w.x, gi[r][g * 2]);
          gi[r][g * 2 + 1] = fmaf(xv, w.y, gi[r][g * 2 + 1]);
        }
      }
    }
  } else {
    pipe_gemm_gru<R>(wp, 64, xin, LDG, gi, lane);
  }
  pipe_gemm_gru<R>(wp, 64, hprev, LDG, gh, lane);
#pragma unroll
  for (int r = 0; r < R; r++) {
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const float rg = sigmoidf_(gi[r][e] + gh[r][e]);
      const float zg = sigmoidf_(gi[r][2 + e] + gh[r][2 + e]);
      const float ng = tanhf(gi[r][4 + e] + rg * gh[r][4 + e]);
      const float hp = hprev[r * LDG + lane * 2 + e];
      hnew[r][e] = (1.0f - zg) * ng + zg * hp;
      if (stash != nullptr) {
        float* st = stash + (size_t)r * 4 * 64 + lane * 2 + e;
        st[0] = rg; st[64] = zg; st[128] = ng; st[192] = gh[r][4 + e];
      }
    }
  }
}

// forward weight order of the GRU: layer 0 has no input-side GEMM (its 4-wide input is applied from L1)
__device__ __forceinline__ void produce_gru_fwd(const ModelDev& M, WPipe& wp) {
  wp_produce(wp, M.seg[S_GH_T0], 64, 192);
  wp_produce(wp, M.seg[S_GI_T1], 64, 192);
  wp_produce(wp, M.seg[S_GH_T1], 64, 192);
  wp_produce(wp, M.seg[S_GI_T2], 64, 192);
  wp_produce(wp, M.seg[S_GH_T2], 64, 192);
}

__global__ void __launch_bounds__(NODE_THREADS, 2) gru_fwd_kernel(ModelDev M, StepArgs a) {
  STRIVE_PDL_TRIGGER();
  STRIVE_PDL_WAIT();      // first instructions of the kernel: nothing (not even a hoisted read-only load) can precede them
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WPipe wp = wp_init(smem, NODE_WARPS);
  if (warp == NODE_WARPS) {
    if (lane == 0) produce_gru_fwd(M, wp);
    return;
  }
  float* xin = smem + WP_SMEM_FLOATS + warp * (GRU_R * LDG * 4 + GRU_R * 4);
  float* hbuf = xin + GRU_R * LDG;          // [3][R][LDG]
  float* locs = hbuf + 3 * GRU_R * LDG;     // [R][4]
  const int NA = a.NA, t = a.t;
  int row[GRU_R], base;
  bool valid[GRU_R];
  node_rows<GRU_R>(warp, NA, base, row, valid);
  prefetch_gru_params(M, lane);
  const float* memg = a.tp.mem + (size_t)t * NA * 192;
  stage_gru_in<GRU_R>(hbuf, locs, memg, a.tp.loc + (size_t)t * NA * 4, row, lane);
  __syncwarp();
  float* memn = a.tp.mem + (size_t)(t + 1) * NA * 192;
  float* pfn = a.tp.pastfeat + (size_t)(t + 1) * NA * 64;
  for (int l = 0; l < 3; l++) {
    float hnew[GRU_R][2];
    gru_layer_fwd<GRU_R>(M, wp, l, xin, locs, hbuf + l * GRU_R * LDG, hnew, nullptr, lane);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < GRU_R; r++) {
      stvec<2>(xin + r * LDG + lane * 2, hnew[r]);
      if (valid[r]) {
        stvec<2>(memn + (size_t)row[r] * 192 + l * 64 + lane * 2, hnew[r]);
        if (l == 2) stvec<2>(pfn + (size_t)row[r] * 64 + lane * 2, hnew[r]);
      }
    }
    __syncwarp();
  }
}

// ======================================================================================================
// backward
// ======================================================================================================

// GRU backward. In: g_mem (grad wrt mem_{t+1}), g_pf (grad wrt past_feat_{t+1} = top output). Out: g_mem <- grad wrt
// mem_t, d_loc. (Not launched for t = FT-1: no GRU step there.)
__global__ void __launch_bounds__(NODE_THREADS, 2) gru_bwd_kernel(ModelDev M, StepArgs a) {
  STRIVE_PDL_TRIGGER();
  STRIVE_PDL_WAIT();      // first instructions of the kernel: nothing (not even a hoisted read-only load) can precede them
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WPipe wp = wp_init(smem, NODE_WARPS);
  if (warp == NODE_WARPS) {
    if (lane == 0) {
      produce_gru_fwd(M, wp);
      wp_produce(wp, M.seg[S_GI_N2], 192, 64);
      wp_produce(wp, M.seg[S_GH_N2], 192, 64);
      wp_produce(wp, M.seg[S_GI_N1], 192, 64);
      wp_produce(wp, M.seg[S_GH_N1], 192, 64);
      wp_produce(wp, M.seg[S_GH_N0], 192, 64);
    }
    return;
  }
  // per-warp: xin[3][R][LDG] (inputs of layers 1,2 = new h of layers 0,1; slot 0 unused), h[3][R][LDG], loc[R][4],
  //           stash[3][R][4][64], dg [R][196]
  constexpr int PER_WARP = 6 * GRU_R * LDG + GRU_R * 4 + 3 * GRU_R * 256 + GRU_R * 196;
  float* xin = smem + WP_SMEM_FLOATS + warp * PER_WARP;
  float* hbuf = xin + 3 * GRU_R * LDG;
  float* locs = hbuf + 3 * GRU_R * LDG;
  float* stash = locs + GRU_R * 4;
  float* dgb = stash + 3 * GRU_R * 256;   // [R][196] gate grads staged for the native GEMMs
  const int NA = a.NA, t = a.t;
  int row[GRU_R], base;
  bool valid[GRU_R];
  node_rows<GRU_R>(warp, NA, base, row, valid);
  prefetch_gru_params(M, lane);
  const float* memg = a.tp.mem + (size_t)t * NA * 192;
  stage_gru_in<GRU_R>(hbuf, locs, memg, a.tp.loc + (size_t)t * NA * 4, row, lane);
  __syncwarp();
  // recompute forward, stash gates
  for (int l = 0; l < 3; l++) {
    float hnew[GRU_R][2];
    gru_layer_fwd<GRU_R>(M, wp, l, xin + l * GRU_R * LDG, locs, hbuf + l * GRU_R * LDG, hnew, stash + l * GRU_R * 256, lane);
    __syncwarp();
    if (l < 2) {
#pragma unroll
      for (int r = 0; r < GRU_R; r++) stvec<2>(xin + ((l + 1) * GRU_R + r) * LDG + lane * 2, hnew[r]);
    }
    __syncwarp();
  }
  // backward, top layer first. dh[r][e]: grad wrt new hidden of layer l for this lane's units.
  float dx_next[GRU_R][2];   // grad wrt the input of layer l+1 (= new hidden of layer l)
#pragma unroll
  for (int r = 0; r < GRU_R; r++) dx_next[r][0] = dx_next[r][1] = 0.f;
  float dloc_acc[GRU_R][4];
#pragma unroll
  for (int r = 0; r < GRU_R; r++)
#pragma unroll
    for (int k = 0; k < 4; k++) dloc_acc[r][k] = 0.f;
  for (int l = 2; l >= 0; l--) {
    const int so = l * 6;
    float dhp[GRU_R][2];   // grad wrt previous hidden (direct z-path part)
#pragma unroll
    for (int r = 0; r < GRU_R; r++) {
      float2 gm = *reinterpret_cast<const float2*>(a.tp.g_mem + (size_t)row[r] * 192 + l * 64 + lane * 2);
      float dh[2] = {gm.x + dx_next[r][0], gm.y + dx_next[r][1]};
      if (l == 2) {
        float2 gp = *reinterpret_cast<const float2*>(a.tp.g_pf + (size_t)row[r] * 64 + lane * 2);
        dh[0] += gp.x; dh[1] += gp.y;
      }
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const float* st = stash + ((size_t)l * GRU_R + r) * 256 + lane * 2 + e;
        const float rg = st[0], zg = st[64], ng = st[128], ghn = st[192];
        const float hp = hbuf[(l * GRU_R + r) * LDG + lane * 2 + e];
        const float dn = dh[e] * (1.0f - zg);
        const float dz = dh[e] * (hp - ng);
        dhp[r][e] = dh[e] * zg;
        const float dpn = dn * (1.0f - ng * ng);
        const float dr = dpn * ghn;
        const float dpz = dz * zg * (1.0f - zg);
        const float dpr = dr * rg * (1.0f - rg);
        // gate pre-activation grads: input side (dgi) = [dpr, dpz, dpn], hidden side (dgh) = [dpr, dpz, dpn*r]
        float* dgi = dgb + r * 196;   // reuse the same row buffer for gi then gh sequentially below
        dgi[lane * 2 + e] = dpr;
        dgi[64 + lane * 2 + e] = dpz;
        dgi[128 + lane * 2 + e] = dpn;
        // stash hidden-side n-gate grad in the stash slot of ghn (no longer needed)
        const_cast<float*>(st)[192] = dpn * rg;
      }
    }
    __syncwarp();
    // input grad: dx = dgi . GI_N[l] ([192][Kin])
    if (l > 0) {
      float dx[GRU_R][2];
      init_zero<2, GRU_R>(dx);
      pipe_gemm<64, GRU_R>(wp, 192, dgb, 196, dx, lane);
#pragma unroll
      for (int r = 0; r < GRU_R; r++) { dx_next[r][0] = dx[r][0]; dx_next[r][1] = dx[r][1]; }
    } else {
      const float* W = M.seg[S_GI_N0];   // [192][4]
#pragma unroll
      for (int r = 0; r < GRU_R; r++) {
        float p[4] = {0.f, 0.f, 0.f, 0.f};
        for (int n = lane; n < 192; n += 32) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(W + n * 4));
          const float g = dgb[r * 196 + n];
          p[0] = fmaf(g, w.x, p[0]); p[1] = fmaf(g, w.y, p[1]); p[2] = fmaf(g, w.z, p[2]); p[3] = fmaf(g, w.w, p[3]);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) dloc_acc[r][k] = warp_sum(p[k]);
      }
    }
    __syncwarp();
    // hidden grad: dh_prev += dgh . GH_N[l]; dgh differs from dgi only in the n-gate block
#pragma unroll
    for (int r = 0; r < GRU_R; r++) {
#pragma unroll
      for (int e = 0; e < 2; e++) dgb[r * 196 + 128 + lane * 2 + e] = stash[((size_t)l * GRU_R + r) * 256 + 192 + lane * 2 + e];
    }
    __syncwarp();
    pipe_gemm<64, GRU_R>(wp, 192, dgb, 196, dhp, lane);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < GRU_R; r++)
      if (valid[r]) stvec<2>(a.tp.g_mem + (size_t)row[r] * 192 + l * 64 + lane * 2, dhp[r]);
  }
  if (lane < 4) {
#pragma unroll
    for (int r = 0; r < GRU_R; r++)
      if (valid[r]) a.tp.d_loc[(size_t)row[r] * 4 + lane] = lane == 0 ? dloc_acc[r][0] : lane == 1 ? dloc_acc[r][1] : lane == 2 ? dloc_acc[r][2] : dloc_acc[r][3];
  }
}

// post backward: (d_traj[t], g_prev, g_pos, d_loc) -> bicycle/transform adjoint -> mlp_out/update_mlp adjoint
// outputs: g_prev <- grad wrt prev_state_t, d_xupd, d_aggr; zeroes g_pos and dQ rows for the edge phase.
__global__ void __launch_bounds__(NODE_THREADS, 2) post_bwd_kernel(ModelDev M, StepArgs a, int has_gru) {
  STRIVE_PDL_TRIGGER();
  STRIVE_PDL_WAIT();      // first instructions of the kernel: nothing (not even a hoisted read-only load) can precede them
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WPipe wp = wp_init(smem, NODE_WARPS);
  if (warp == NODE_WARPS) {
    if (lane == 0) {
      produce_post_mlps(M, wp);
      wp_produce(wp, M.seg[S_O3_N], 128, 128);
      wp_produce(wp, M.seg[S_O0_N], 128, 64);
      wp_produce(wp, M.seg[S_U3_N], 64, 128);
      wp_produce(wp, M.seg[S_U0_N_X], 128, 64);
      wp_produce(wp, M.seg[S_U0_N_AGGR], 128, 64);
    }
    return;
  }
  constexpr int PER_WARP = NODE_R * (LDA + LDH) + 3 * NODE_R * LDH;
  float* bufA = smem + WP_SMEM_FLOATS + warp * PER_WARP;
  float* bufB = bufA + NODE_R * LDA;
  float* preU = bufB + NODE_R * LDH;
  float* pre1 = preU + NODE_R * LDH;
  float* pre2 = pre1 + NODE_R * LDH;
  const int NA = a.NA, t = a.t, FT = a.FT;
  int row[NODE_R], base;
  bool valid[NODE_R];
  node_rows<NODE_R>(warp, NA, base, row, valid);
  prefetch_post_params(M, lane);
  stage_update_in<NODE_R>(bufA, a, row, lane, M.u0_rows);
  __syncwarp();
  float o[NODE_R][2];
  post_mlps_fwd<NODE_R>(M, wp, bufA, bufB, preU, pre1, pre2, o, lane);
  // scalar adjoint: lane r handles row r, then d_o is broadcast to the warp
  float my_o0 = 0.f, my_o1 = 0.f;
#pragma unroll
  for (int r = 0; r < NODE_R; r++)
    if (lane == r) { my_o0 = o[r][0]; my_o1 = o[r][1]; }
  float d_o0 = 0.f, d_o1 = 0.f;
  if (lane < NODE_R && base + lane < NA) {
    const int ag = base + lane;
    const float* pv = a.tp.prev + ((size_t)t * NA + ag) * 6;
    float prev[6], cur[6];
#pragma unroll
    for (int k = 0; k < 6; k++) prev[k] = pv[k];
    BikeFwd b;
    bicycle_fwd(prev, my_o0, my_o1, a.lw[(size_t)ag * 2], cur, b);
    const int s = a.scene_of[ag];
    const bool is_ext = (a.ext != nullptr && ag == a.ptr[s]);
    float dcur[6];
#pragma unroll
    for (int k = 0; k < 6; k++) dcur[k] = a.tp.g_prev[(size_t)ag * 6 + k];
    const float4 dt = *reinterpret_cast<const float4*>(a.d_traj + ((size_t)ag * FT + t) * 4);
    dcur[0] += dt.x; dcur[1] += dt.y; dcur[2] += dt.z; dcur[3] += dt.w;
    float dprev[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (!is_ext) {
#pragma unroll
      for (int k = 0; k < 4; k++) dcur[k] += a.tp.g_pos[(size_t)ag * 4 + k];
    }
    if (has_gru) {
      float gl[4];
#pragma unroll
      for (int k = 0; k < 4; k++) gl[k] = a.tp.d_loc[(size_t)ag * 4 + k];
      float glob[4] = {cur[0], cur[1], cur[2], cur[3]};
      if (is_ext) {
        const float* e = a.ext + ((size_t)s * FT + t) * 4;
#pragma unroll
        for (int k = 0; k < 4; k++) glob[k] = e[k];
      }
      float dpose[4] = {0.f, 0.f, 0.f, 0.f};
      t2f_bwd(prev, glob, gl, dprev, dpose);
      if (!is_ext) {
#pragma unroll
        for (int k = 0; k < 4; k++) dcur[k] += dpose[k];
      }
    }
    // normalise adjoint
    float db[6];
#pragma unroll
    for (int k = 0; k < 6; k++) db[k] = dcur[k] / kStateStd[k];
    float d_newh = -b.sn * db[2] + b.cs * db[3];
    float d_news = db[4];
    float d_newhdot = db[5];
    float dun[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    dun[0] += db[0];
    d_news += db[0] * b.cs * BIKE_DT;
    d_newh += -db[0] * b.news * b.sn * BIKE_DT;
    dun[1] += db[1];
    d_news += db[1] * b.sn * BIKE_DT;
    d_newh += db[1] * b.news * b.cs * BIKE_DT;
    const float pass_s = (b.pre_s >= 0.0f && b.pre_s <= BIKE_MAXS) ? 1.0f : 0.0f;
    dun[4] += pass_s * d_news;
    const float d_a = pass_s * d_news * BIKE_DT;
    float d_h = d_newh;
    const float sg = (b.un[4] > 0.f) ? 1.0f : ((b.un[4] < 0.f) ? -1.0f : 0.0f);
    dun[4] += d_newh * BIKE_DT * sg / b.len * b.newhdot;
    d_newhdot += d_newh * BIKE_DT * fabsf(b.un[4]) / b.len;
    const float pass_h = (b.pre_hd >= -BIKE_MAXHDOT && b.pre_hd <= BIKE_MAXHDOT) ? 1.0f : 0.0f;
    dun[5] += pass_h * d_newhdot;
    const float d_ddh = pass_h * d_newhdot * BIKE_DT;
    const float nrm = b.un[2] * b.un[2] + b.un[3] * b.un[3];
    dun[2] += d_h * (-b.un[3] / nrm);
    dun[3] += d_h * (b.un[2] / nrm);
#pragma unroll
    for (int k = 0; k < 6; k++) dprev[k] += dun[k] * kStateStd[k];
#pragma unroll
    for (int k = 0; k < 6; k++) a.tp.g_prev[(size_t)ag * 6 + k] = dprev[k];
    // hand g_pos over to the edge phase as a zeroed accumulator
#pragma unroll
    for (int k = 0; k < 4; k++) a.tp.g_pos[(size_t)ag * 4 + k] = 0.f;
    d_o0 = d_a * A_STD;
    d_o1 = d_ddh * DDH_STD;
  }
  // zero dQ rows (edge_bwd accumulates into them atomically)
#pragma unroll
  for (int r = 0; r < NODE_R; r++)
    if (valid[r]) *reinterpret_cast<float4*>(a.tp.dQ + (size_t)row[r] * 128 + lane * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  // mlp_out backward
  float dh[NODE_R][4];
  {
    float w0[4], w1[4];
    ldvec<4>(w0, M.seg[S_O6_N] + lane * 4);
    ldvec<4>(w1, M.seg[S_O6_N] + 128 + lane * 4);
#pragma unroll
    for (int r = 0; r < NODE_R; r++) {
      const float g0 = __shfl_sync(0xffffffffu, d_o0, r);
      const float g1 = __shfl_sync(0xffffffffu, d_o1, r);
#pragma unroll
      for (int v = 0; v < 4; v++) dh[r][v] = g0 * w0[v] + g1 * w1[v];
    }
  }
  ln_relu_bwd<NODE_R>(dh, pre2, LDH, M.seg[S_O_LN4_G], M.seg[S_O_LN4_B], lane);
  __syncwarp();
  store_rows<4, NODE_R>(dh, bufB, LDH, lane);
  __syncwarp();
  init_zero<4, NODE_R>(dh);
  pipe_gemm<128, NODE_R>(wp, 128, bufB, LDH, dh, lane);
  ln_relu_bwd<NODE_R>(dh, pre1, LDH, M.seg[S_O_LN1_G], M.seg[S_O_LN1_B], lane);
  __syncwarp();
  store_rows<4, NODE_R>(dh, bufB, LDH, lane);
  __syncwarp();
  float dxu[NODE_R][2];
  init_zero<2, NODE_R>(dxu);
  pipe_gemm<64, NODE_R>(wp, 128, bufB, LDH, dxu, lane);
  __syncwarp();
  // update_mlp backward
  store_rows<2, NODE_R>(dxu, bufA, LDA, lane);
  __syncwarp();
  init_zero<4, NODE_R>(dh);
  pipe_gemm<128, NODE_R>(wp, 64, bufA, LDA, dh, lane);
  ln_relu_bwd<NODE_R>(dh, preU, LDH, M.seg[S_U_LN1_G], M.seg[S_U_LN1_B], lane);
  __syncwarp();
  store_rows<4, NODE_R>(dh, bufB, LDH, lane);
  __syncwarp();
  float dxx[NODE_R][2], dag[NODE_R][2];
  init_zero<2, NODE_R>(dxx);
  init_zero<2, NODE_R>(dag);
  pipe_gemm<64, NODE_R>(wp, 128, bufB, LDH, dxx, lane);
  pipe_gemm<64, NODE_R>(wp, 128, bufB, LDH, dag, lane);
#pragma unroll
  for (int r = 0; r < NODE_R; r++) {
    if (valid[r]) {
      stvec<2>(a.tp.d_xupd + (size_t)row[r] * 64 + lane * 2, dxx[r]);
      stvec<2>(a.tp.d_aggr + (size_t)row[r] * 64 + lane * 2, dag[r]);
    }
  }
}

// edge backward: recompute the edge MLP per chunk, route d_aggr to the arg-max edges, back through the MLP.
// dP_i written (exclusive), dQ_j / g_pos_j accumulated atomically, g_pos_i accumulated atomically.
__global__ void __launch_bounds__(EDGE_WARPS * 32) edge_bwd_kernel(ModelDev M, StepArgs a) {
  STRIVE_PDL_TRIGGER();
  STRIVE_PDL_WAIT();
  extern __shared__ __align__(16) float smem[];
  __shared__ float red_dp[EDGE_WARPS][128];
  __shared__ float red_pos[EDGE_WARPS][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* bufA = smem + warp * (4 * EDGE_R * LDH);
  float* bufB = bufA + EDGE_R * LDH;
  float* pre1 = bufB + EDGE_R * LDH;
  float* pre2 = pre1 + EDGE_R * LDH;
  EdgeCtx c;
  edge_ctx_init(a, c);
  const int NA = a.NA;
  float Pi[4], Wrel[4][4];
  ldvec<4>(Pi, a.tp.P + ((size_t)a.t * NA + c.i) * 128 + lane * 4);
#pragma unroll
  for (int d = 0; d < 4; d++) ldvec<4>(Wrel[d], M.seg[S_E0_T_REL] + d * 128 + lane * 4);
  const float2 dagg = *reinterpret_cast<const float2*>(a.tp.d_aggr + (size_t)c.i * 64 + lane * 2);
  const uchar2 am = *reinterpret_cast<const uchar2*>(a.tp.arg + ((size_t)a.t * NA + c.i) * 64 + lane * 2);
  float dP[4] = {0.f, 0.f, 0.f, 0.f};
  float dposi[4] = {0.f, 0.f, 0.f, 0.f};
  const float* posg = a.tp.pos + (size_t)a.t * NA * 4;
  const int nchunks = (c.ne + EDGE_R - 1) / EDGE_R;
  for (int chunk = warp; chunk < nchunks; chunk += EDGE_WARPS) {
    float acc[EDGE_R][4], rel[EDGE_R][4];
    int lj[EDGE_R];
    bool valid[EDGE_R];
    edge_layer0<EDGE_R>(a, c, chunk, Pi, Wrel, acc, lj, valid, rel, lane);
    __syncwarp();
    ln_relu_store<EDGE_R>(acc, M.seg[S_E_LN1_G], M.seg[S_E_LN1_B], bufA, LDH, pre1, LDH, lane);
    __syncwarp();
    init_bias<4, EDGE_R>(acc, M.seg[S_E3_B], lane);
    warp_gemm<128, EDGE_R>(M.seg[S_E3_T], 128, bufA, LDH, acc, lane);
    __syncwarp();
    ln_relu_store<EDGE_R>(acc, M.seg[S_E_LN4_G], M.seg[S_E_LN4_B], bufB, LDH, pre2, LDH, lane);
    __syncwarp();
    // d_m rows (64 wide) into bufA (ld LDH)
#pragma unroll
    for (int r = 0; r < EDGE_R; r++) {
      float dm[2];
      dm[0] = (valid[r] && (int)am.x == lj[r]) ? dagg.x : 0.f;
      dm[1] = (valid[r] && (int)am.y == lj[r]) ? dagg.y : 0.f;
      stvec<2>(bufA + r * LDH + lane * 2, dm);
    }
    __syncwarp();
    float dh[EDGE_R][4];
    init_zero<4, EDGE_R>(dh);
    warp_gemm<128, EDGE_R>(M.seg[S_E6_N], 64, bufA, LDH, dh, lane);
    ln_relu_bwd<EDGE_R>(dh, pre2, LDH, M.seg[S_E_LN4_G], M.seg[S_E_LN4_B], lane);
    __syncwarp();
    store_rows<4, EDGE_R>(dh, bufA, LDH, lane);
    __syncwarp();
    init_zero<4, EDGE_R>(dh);
    warp_gemm<128, EDGE_R>(M.seg[S_E3_N], 128, bufA, LDH, dh, lane);
    ln_relu_bwd<EDGE_R>(dh, pre1, LDH, M.seg[S_E_LN1_G], M.seg[S_E_LN1_B], lane);
    __syncwarp();
    // dh = grad wrt h1pre = P_i + Q_j + W_rel rel
#pragma unroll
    for (int r = 0; r < EDGE_R; r++) {
      if (!valid[r]) continue;   // warp-uniform
      const int j = c.p0 + lj[r];
      float drel[4];
#pragma unroll
      for (int d = 0; d < 4; d++) {
        float p = dh[r][0] * Wrel[d][0] + dh[r][1] * Wrel[d][1] + dh[r][2] * Wrel[d][2] + dh[r][3] * Wrel[d][3];
        drel[d] = warp_sum(p);
      }
#pragma unroll
      for (int v = 0; v < 4; v++) {
        dP[v] += dh[r][v];
        atomicAdd(a.tp.dQ + (size_t)j * 128 + lane * 4 + v, dh[r][v]);
      }
      const float4 pj4 = __ldg(reinterpret_cast<const float4*>(posg + (size_t)j * 4));
      const float pj[4] = {pj4.x, pj4.y, pj4.z, pj4.w};
      float relchk[4];
      t2f_fwd(c.pos_i, pj, relchk);
#pragma unroll
      for (int d = 0; d < 4; d++)
        if (isnan(relchk[d])) drel[d] = 0.f;
      float dpj[4] = {0.f, 0.f, 0.f, 0.f};
      t2f_bwd(c.pos_i, pj, drel, dposi, dpj);
      if (lane < 4) atomicAdd(a.tp.g_pos + (size_t)j * 4 + lane, lane == 0 ? dpj[0] : lane == 1 ? dpj[1] : lane == 2 ? dpj[2] : dpj[3]);
    }
    __syncwarp();
  }
#pragma unroll
  for (int v = 0; v < 4; v++) red_dp[warp][lane * 4 + v] = dP[v];
  if (lane < 4) red_pos[warp][lane] = lane == 0 ? dposi[0] : lane == 1 ? dposi[1] : lane == 2 ? dposi[2] : dposi[3];
  __syncthreads();
  {
    const int ch = threadIdx.x;   // 128 threads, 128 columns
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < EDGE_WARPS; w++) s += red_dp[w][ch];
    a.tp.dP[(size_t)c.i * 128 + ch] = s;
    if (ch < 4) {
      float p = 0.f;
#pragma unroll
      for (int w = 0; w < EDGE_WARPS; w++) p += red_pos[w][ch];
      atomicAdd(a.tp.g_pos + (size_t)c.i * 4 + ch, p);
    }
  }
}

#include "edge_mma.cuh"
#include "edge_tc.cuh"

// node backward: d_x = d_xupd + dP.W_xi + dQ.W_xj ; back through mlp_in; d_z += ; g_pf <- grad wrt past_feat_t
__global__ void __launch_bounds__(NODE_THREADS, 2) node_bwd_kernel(ModelDev M, StepArgs a) {
  STRIVE_PDL_TRIGGER();
  STRIVE_PDL_WAIT();      // first instructions of the kernel: nothing (not even a hoisted read-only load) can precede them
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WPipe wp = wp_init(smem, NODE_WARPS);
  if (warp == NODE_WARPS) {
    if (lane == 0) {
      produce_mlp_in(M, wp);
      wp_produce(wp, M.seg[S_E0_N_XI], 128, 64);
      wp_produce(wp, M.seg[S_E0_N_XJ], 128, 64);
      wp_produce(wp, M.seg[S_IN6_N], 64, 128);
      wp_produce(wp, M.seg[S_IN3_N], 128, 128);
      wp_produce(wp, M.seg[S_IN0_N_PF], 128, 64);
      wp_produce(wp, M.seg[S_IN0_N_Z], 128, 32);
    }
    return;
  }
  constexpr int PER_WARP = NODE_R * (LDA + LDH) + 2 * NODE_R * LDH;
  float* bufA = smem + WP_SMEM_FLOATS + warp * PER_WARP;
  float* bufB = bufA + NODE_R * LDA;
  float* pre1 = bufB + NODE_R * LDH;
  float* pre2 = pre1 + NODE_R * LDH;
  const int NA = a.NA;
  int row[NODE_R], base;
  bool valid[NODE_R];
  node_rows<NODE_R>(warp, NA, base, row, valid);
  prefetch_mlp_in_params(M, lane);
  stage_node_feat<NODE_R>(bufA, a, row, lane, M.in0_rows);
  __syncwarp();
  float xacc[NODE_R][2];
  mlp_in_fwd<NODE_R>(M, wp, bufA, bufB, pre1, pre2, xacc, lane);
  // d_x
  float dx[NODE_R][2];
#pragma unroll
  for (int r = 0; r < NODE_R; r++) {
    const float2 u = *reinterpret_cast<const float2*>(a.tp.d_xupd + (size_t)row[r] * 64 + lane * 2);
    dx[r][0] = u.x; dx[r][1] = u.y;
    *reinterpret_cast<float4*>(bufB + r * LDH + lane * 4) = *reinterpret_cast<const float4*>(a.tp.dP + (size_t)row[r] * 128 + lane * 4);
  }
  __syncwarp();
  pipe_gemm<64, NODE_R>(wp, 128, bufB, LDH, dx, lane);
  __syncwarp();
#pragma unroll
  for (int r = 0; r < NODE_R; r++)
    *reinterpret_cast<float4*>(bufB + r * LDH + lane * 4) = *reinterpret_cast<const float4*>(a.tp.dQ + (size_t)row[r] * 128 + lane * 4);
  __syncwarp();
  pipe_gemm<64, NODE_R>(wp, 128, bufB, LDH, dx, lane);
  __syncwarp();
  store_rows<2, NODE_R>(dx, bufA, LDA, lane);
  __syncwarp();
  float dh[NODE_R][4];
  init_zero<4, NODE_R>(dh);
  pipe_gemm<128, NODE_R>(wp, 64, bufA, LDA, dh, lane);
  ln_relu_bwd<NODE_R>(dh, pre2, LDH, M.seg[S_IN_LN4_G], M.seg[S_IN_LN4_B], lane);
  __syncwarp();
  store_rows<4, NODE_R>(dh, bufB, LDH, lane);
  __syncwarp();
  init_zero<4, NODE_R>(dh);
  pipe_gemm<128, NODE_R>(wp, 128, bufB, LDH, dh, lane);
  ln_relu_bwd<NODE_R>(dh, pre1, LDH, M.seg[S_IN_LN1_G], M.seg[S_IN_LN1_B], lane);
  __syncwarp();
  store_rows<4, NODE_R>(dh, bufB, LDH, lane);
  __syncwarp();
  float dpf[NODE_R][2], dz[NODE_R][1];
  init_zero<2, NODE_R>(dpf);
  init_zero<1, NODE_R>(dz);
  pipe_gemm<64, NODE_R>(wp, 128, bufB, LDH, dpf, lane);
  pipe_gemm<32, NODE_R>(wp, 128, bufB, LDH, dz, lane);
#pragma unroll
  for (int r = 0; r < NODE_R; r++) {
    if (valid[r]) {
      stvec<2>(a.tp.g_pf + (size_t)row[r] * 64 + lane * 2, dpf[r]);
      a.d_z[(size_t)row[r] * ZDIM + lane] += dz[r][0];
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// small utility kernels
// ------------------------------------------------------------------------------------------------------
__global__ void init_tape_kernel(StepArgs a, const float* past_last, const float* map_feat0, const float* past_feat0,
                                 const int32_t* map_idx) {
  STRIVE_PDL_TRIGGER();
  STRIVE_PDL_WAIT();
  const int NA = a.NA;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NA * 64) return;
  const int ag = i >> 6, k = i & 63;
  const float pf = past_feat0[i];
  a.tp.pastfeat[i] = pf;
  a.tp.mapfeat[i] = map_feat0[i];
  a.tp.mem[(size_t)ag * 192 + k] = pf;            // traffic_model.py:625 hidden init = past_feat x 3 layers
  a.tp.mem[(size_t)ag * 192 + 64 + k] = pf;
  a.tp.mem[(size_t)ag * 192 + 128 + k] = pf;
  if (k < 6) a.tp.prev[(size_t)ag * 6 + k] = past_last[(size_t)ag * 6 + k];
  if (k < 4) a.tp.pos[(size_t)ag * 4 + k] = past_last[(size_t)ag * 6 + k];   // :604
  if (k == 0) a.tp.map_of[ag] = map_idx[a.scene_of[ag]];
  if (k < ZDIM) a.tp.z[(size_t)ag * ZDIM + k] = a.z[(size_t)ag * ZDIM + k];
}

// ------------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------------
static ModelDev model_dev(const StriveModel* m) {
  ModelDev d;
  for (int i = 0; i < S_COUNT; i++) d.seg[i] = m->seg[i];
  d.nc = m->nc;
  d.in0_rows = m->in0_rows;
  d.u0_rows = m->u0_rows;
  return d;
}

static const size_t SM_PIPE = WP_SMEM_FLOATS * 4;
static const size_t SM_NODE = SM_PIPE + NODE_WARPS * NODE_R * (LDA + LDH) * 4;
static const size_t SM_EDGE_F = EDGE_WARPS * 2 * EDGE_R * LDH * 4;
static const size_t SM_EDGE_B = EDGE_WARPS * 4 * EDGE_R * LDH * 4;
static const size_t SM_GRU_F = SM_PIPE + NODE_WARPS * (GRU_R * LDG * 4 + GRU_R * 4) * 4;
static const size_t SM_GRU_B = SM_PIPE + NODE_WARPS * (6 * GRU_R * LDG + GRU_R * 4 + 3 * GRU_R * 256 + GRU_R * 196) * 4;
static const size_t SM_POST_B = SM_PIPE + NODE_WARPS * (NODE_R * (LDA + LDH) + 3 * NODE_R * LDH) * 4;
static const size_t SM_NODE_B = SM_PIPE + NODE_WARPS * (NODE_R * (LDA + LDH) + 2 * NODE_R * LDH) * 4;

// 3 = tcgen05 forward + backward (edge_tc.cuh; default; scenes up to ET_MAX_N agents, larger ones use the mma.sync kernels),
// 2 = tcgen05 forward + mma.sync backward, 1 = mma.sync TF32 kernels both ways (edge_mma.cuh), 0 = fp32 SIMT kernels (A/B verification)
static int g_edge_impl = 3;
extern "C" int strive_edge_set_impl(int impl) {
  g_edge_impl = impl;
  return 0;
}

static int em_grid(int NA) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int want = (NA + EM_WARPS - 1) / EM_WARPS;
  return want < sms ? want : sms;
}

extern "C" int64_t strive_model_edge_frag_bytes(void) { return EM_FRAG_BYTES + ET_PACK_BYTES; }

// Packs the edge-MLP matrices of the model's weight blob into mma.sync fragment order (edge_mma.cuh) inside `buf`
// (device, 16-byte aligned, strive_model_edge_frag_bytes() bytes, owned by the caller for the lifetime of the model).
extern "C" int strive_model_set_edge_frags(StriveModel* m, void* buf, int64_t bytes, void* stream_) {
  STRIVE_CHECK(m != nullptr && buf != nullptr, STRIVE_EINVAL, "set_edge_frags: null argument");
  STRIVE_CHECK(bytes == EM_FRAG_BYTES + ET_PACK_BYTES && ((uintptr_t)buf & 15) == 0, STRIVE_ESIZE, "edge fragment buffer: %lld bytes (need %d, 16-byte aligned)",
               (long long)bytes, (int)(EM_FRAG_BYTES + ET_PACK_BYTES));
  STRIVE_CHECK(m->seg_size[S_E3_T] == 128 * 128 && m->seg_size[S_E6_T] == 128 * 64 && m->seg_size[S_E6_N] == 64 * 128 && m->seg_size[S_E3_N] == 128 * 128,
               STRIVE_ESIZE, "edge MLP is not 128-128-64");
  cudaStream_t stream = (cudaStream_t)stream_;
  uint8_t* b = (uint8_t*)buf;
  auto pack = [&](const float* W, int K, int N, int off, bool bf16) {
    const int n = (K / 8) * (N / 8) * 32;
    edge_frag_pack_kernel<<<(n + 255) / 256, 256, 0, stream>>>(W, K, N, bf16 ? nullptr : (float*)(b + off), bf16 ? (uint16_t*)(b + off) : nullptr);
  };
  pack(m->seg[S_E3_T], 128, 128, EM_F3_OFF, false);
  pack(m->seg[S_E6_T], 128, 64, EM_F6_OFF, false);
  pack(m->seg[S_E3_T], 128, 128, EM_B3T_OFF, true);
  pack(m->seg[S_E6_N], 64, 128, EM_B6N_OFF, true);
  pack(m->seg[S_E3_N], 128, 128, EM_B3N_OFF, true);
  STRIVE_LAUNCH_CHECK();
  // tcgen05 operand packs (edge_tc.cuh): native [out][in] matrices -> bf16 hi / lo, canonical K-major
  uint8_t* et = b + EM_FRAG_BYTES;
  edge_tc_pack_kernel<<<(128 * 128 + 255) / 256, 256, 0, stream>>>(m->seg[S_E3_N], 128, 128, et + ET_W3H_OFF, et + ET_W3L_OFF);
  edge_tc_pack_kernel<<<(64 * 128 + 255) / 256, 256, 0, stream>>>(m->seg[S_E6_N], 64, 128, et + ET_W6H_OFF, et + ET_W6L_OFF);
  STRIVE_LAUNCH_CHECK();
  m->edge_frags = b;
  return 0;
}

static int set_smem_attrs() {
  static unsigned done = 0;
  if (!strive_first_use_on_device(&done)) return 0;
  STRIVE_CUDA(cudaFuncSetAttribute(edge_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_EDGE_B));
  STRIVE_CUDA(cudaFuncSetAttribute(edge_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EM_FWD_SMEM));
  STRIVE_CUDA(cudaFuncSetAttribute(edge_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EM_BWD_SMEM));
  STRIVE_CUDA(cudaFuncSetAttribute(edge_fwd_tc_kernel<ET_CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ET_SMEM));
  STRIVE_CUDA(cudaFuncSetAttribute(edge_bwd_tc_kernel<ETB_CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ETB_SMEM));
  STRIVE_CUDA(cudaFuncSetAttribute(node_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_NODE));
  STRIVE_CUDA(cudaFuncSetAttribute(post_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_NODE));
  STRIVE_CUDA(cudaFuncSetAttribute(gru_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_GRU_F));
  STRIVE_CUDA(cudaFuncSetAttribute(gru_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_GRU_B));
  STRIVE_CUDA(cudaFuncSetAttribute(post_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_POST_B));
  STRIVE_CUDA(cudaFuncSetAttribute(node_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_NODE_B));
  return 0;
}

static int check_scene(const StriveModel* m, const StriveScene* sc, int ft) {
  STRIVE_CHECK(m != nullptr && sc != nullptr, STRIVE_EINVAL, "null model/scene");
  STRIVE_CHECK(sc->num_agents > 0 && sc->num_scenes > 0 && ft > 0, STRIVE_EINVAL, "empty scene batch (NA=%d S=%d FT=%d)",
               sc->num_agents, sc->num_scenes, ft);
  STRIVE_CHECK(sc->max_scene_agents <= 255, STRIVE_EUNSUPPORTED, "scene with %d agents > 255 unsupported", sc->max_scene_agents);
  STRIVE_CHECK(sc->num_classes == m->nc, STRIVE_EINVAL, "scene NC=%d != model NC=%d", sc->num_classes, m->nc);
  return 0;
}

// side stream + fork / join events of device `dev`, one set per user (0: forward GRU branch, 1: second backward sweep)
static int side_stream(int which, cudaStream_t* side, cudaEvent_t* ev_fork, cudaEvent_t* ev_join) {
  static cudaStream_t s_[2][16] = {};
  static cudaEvent_t e0_[2][16] = {}, e1_[2][16] = {};
  cudaStream_t* s = s_[which];
  cudaEvent_t *e0 = e0_[which], *e1 = e1_[which];
  int dev = 0;
  STRIVE_CUDA(cudaGetDevice(&dev));
  STRIVE_CHECK(dev >= 0 && dev < 16, STRIVE_EUNSUPPORTED, "device index %d", dev);
  if (s[dev] == nullptr) {
    STRIVE_CUDA(cudaStreamCreateWithFlags(&s[dev], cudaStreamNonBlocking));
    STRIVE_CUDA(cudaEventCreateWithFlags(&e0[dev], cudaEventDisableTiming));
    STRIVE_CUDA(cudaEventCreateWithFlags(&e1[dev], cudaEventDisableTiming));
  }
  *side = s[dev]; *ev_fork = e0[dev]; *ev_join = e1[dev];
  return 0;
}

static int fwd_side_stream(cudaStream_t* side, cudaEvent_t* ev_fork, cudaEvent_t* ev_join) { return side_stream(0, side, ev_fork, ev_join); }
static int bwd_side_stream(cudaStream_t* side, cudaEvent_t* ev_fork, cudaEvent_t* ev_join) { return side_stream(1, side, ev_fork, ev_join); }

extern "C" int strive_decode_fwd(const StriveModel* m, const StriveScene* sc, const StriveMap* map, const float* z,
                                 const float* map_feat0, const float* past_feat0, const float* ext_future, int32_t ft,
                                 float* traj_out, void* tape, int64_t tape_bytes, void* stream_) {
  int rc = check_scene(m, sc, ft);
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int NA = sc->num_agents;
  StepArgs a;
  const int64_t need = tape_carve(&a.tp, (char*)tape, NA, ft);
  STRIVE_CHECK(tape_bytes >= need, STRIVE_ESIZE, "tape too small: %lld < %lld", (long long)tape_bytes, (long long)need);
  rc = set_smem_attrs();
  if (rc) return rc;
  a.NA = NA; a.FT = ft; a.NC = sc->num_classes; a.t = 0;
  a.ptr = sc->ptr; a.scene_of = sc->scene_of; a.lw = sc->lw; a.sem = sc->sem; a.z = z; a.ext = ext_future;
  a.traj = traj_out; a.d_traj = nullptr; a.d_z = nullptr;
  ModelDev M = model_dev(m);
  KPROF("init_tape", stream, STRIVE_CUDA_LAUNCH(init_tape_kernel, (NA * 64 + 255) / 256, 256, 0, stream, a, sc->past_last, map_feat0, past_feat0, sc->map_idx));
  STRIVE_LAUNCH_CHECK();
  const int node_blocks = (NA + NODE_WARPS * NODE_R - 1) / (NODE_WARPS * NODE_R);
  const bool edge_tc = g_edge_impl >= 2 && m->edge_frags != nullptr && sc->max_scene_agents <= ET_MAX_N;
  cudaStream_t side;
  cudaEvent_t ev_fork, ev_join;
  rc = fwd_side_stream(&side, &ev_fork, &ev_join);
  if (rc) return rc;
  if (edge_tc) {
    KPROF("edge_tiles", stream, edge_tc_tiles_kernel<<<1, 1024, 0, stream>>>(sc->ptr, sc->num_scenes, a.tp.et_tiles, a.tp.et_ntiles));
    STRIVE_LAUNCH_CHECK();
  }
  for (int t = 0; t < ft; t++) {
    a.t = t;
    KPROF("node_fwd", stream, STRIVE_CUDA_LAUNCH(node_fwd_kernel, node_blocks, NODE_THREADS, SM_NODE, stream, M, a));
    STRIVE_LAUNCH_CHECK();
    if (edge_tc) {
      const int grid = NA < em_grid(NA * EM_WARPS) ? NA : em_grid(NA * EM_WARPS);      // min(#SMs, upper bound of the tile count)
      KPROF("edge_fwd", stream, STRIVE_CUDA_LAUNCH(edge_fwd_tc_kernel<ET_CPT>, grid, ET_THREADS, ET_SMEM, stream, M, a, m->edge_frags + EM_FRAG_BYTES,
                                                   (const int32_t*)a.tp.et_tiles, (const int32_t*)a.tp.et_ntiles));
    } else if (g_edge_impl != 0 && m->edge_frags != nullptr) {
      KPROF("edge_fwd", stream, STRIVE_CUDA_LAUNCH(edge_fwd_mma_kernel, em_grid(NA), EM_THREADS, EM_FWD_SMEM, stream, M, a, m->edge_frags));
    } else {
      KPROF("edge_fwd", stream, STRIVE_CUDA_LAUNCH(edge_fwd_kernel, NA, EDGE_WARPS * 32, SM_EDGE_F, stream, M, a));
    }
    STRIVE_LAUNCH_CHECK();
    KPROF("post_fwd", stream, STRIVE_CUDA_LAUNCH(post_fwd_kernel, node_blocks, NODE_THREADS, SM_NODE, stream, M, a));
    STRIVE_LAUNCH_CHECK();
    if (t + 1 < ft) {
      // the GRU step (loc -> mem, past_feat of t+1) and the map re-encode (pose -> map_feat of t+1) both hang off post_fwd and
      // meet again in node_fwd(t+1): the GRU kernel -- one partial wave -- runs on the side stream beside the encoder
      STRIVE_CUDA(cudaEventRecord(ev_fork, stream));
      STRIVE_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
      KPROF("gru_fwd", side, STRIVE_CUDA_LAUNCH(gru_fwd_kernel, node_blocks, NODE_THREADS, SM_GRU_F, side, M, a));
      const cudaError_t gru_err = cudaGetLastError();
      STRIVE_CUDA(cudaEventRecord(ev_join, side));                 // joined whatever happens: a capture must not end with a dangling branch
      rc = strive_mapenc_fwd(m, map, a.tp.pose, a.tp.map_of, NA, a.tp.mapfeat + (size_t)(t + 1) * NA * 64, a.tp.mapenc_ws,
                             a.tp.mapenc_ws_bytes, stream_);
      STRIVE_CUDA(cudaStreamWaitEvent(stream, ev_join, 0));
      STRIVE_CHECK(gru_err == cudaSuccess, 100 + (int)gru_err, "gru_fwd launch: %s", cudaGetErrorString(gru_err));
      if (rc) return rc;
    }
  }
  return 0;
}

// one adjoint sweep t = FT-1 .. 0 over the tape on `stream`; `a` carries the seed (d_traj), the output (d_z) and the carry set
static int bwd_sweep(const StriveModel* m, const StriveScene* sc, int32_t ft, const ModelDev& M, StepArgs a, bool edge_tc, cudaStream_t stream) {
  const int NA = sc->num_agents;
  STRIVE_CUDA(cudaMemsetAsync(a.d_z, 0, (size_t)NA * ZDIM * 4, stream));
  STRIVE_CUDA(cudaMemsetAsync(a.tp.g_prev, 0, (size_t)NA * 6 * 4, stream));
  STRIVE_CUDA(cudaMemsetAsync(a.tp.g_pos, 0, (size_t)NA * 4 * 4, stream));
  STRIVE_CUDA(cudaMemsetAsync(a.tp.g_pf, 0, (size_t)NA * 64 * 4, stream));
  STRIVE_CUDA(cudaMemsetAsync(a.tp.g_mem, 0, (size_t)NA * 192 * 4, stream));
  STRIVE_CUDA(cudaMemsetAsync(a.tp.d_loc, 0, (size_t)NA * 4 * 4, stream));
  const int node_blocks = (NA + NODE_WARPS * NODE_R - 1) / (NODE_WARPS * NODE_R);
  for (int t = ft - 1; t >= 0; t--) {
    a.t = t;
    const int has_gru = (t + 1 < ft) ? 1 : 0;
    if (has_gru) {
      KPROF("gru_bwd", stream, STRIVE_CUDA_LAUNCH(gru_bwd_kernel, node_blocks, NODE_THREADS, SM_GRU_B, stream, M, a));
      STRIVE_LAUNCH_CHECK();
    }
    KPROF("post_bwd", stream, STRIVE_CUDA_LAUNCH(post_bwd_kernel, node_blocks, NODE_THREADS, SM_POST_B, stream, M, a, has_gru));
    STRIVE_LAUNCH_CHECK();
    if (edge_tc) {
      const int grid = NA < em_grid(NA * EM_WARPS) ? NA : em_grid(NA * EM_WARPS);
      KPROF("edge_bwd", stream, STRIVE_CUDA_LAUNCH(edge_bwd_tc_kernel<ETB_CPT>, grid, 128 * (128 / ETB_CPT), ETB_SMEM, stream, M, a, m->edge_frags + EM_FRAG_BYTES,
                                                   (const int32_t*)a.tp.et_tiles, (const int32_t*)a.tp.et_ntiles));
    } else if (g_edge_impl != 0 && m->edge_frags != nullptr) {
      KPROF("edge_bwd", stream, STRIVE_CUDA_LAUNCH(edge_bwd_mma_kernel, em_grid(NA), EM_THREADS, EM_BWD_SMEM, stream, M, a, m->edge_frags));
    } else {
      KPROF("edge_bwd", stream, STRIVE_CUDA_LAUNCH(edge_bwd_kernel, NA, EDGE_WARPS * 32, SM_EDGE_B, stream, M, a));
    }
    STRIVE_LAUNCH_CHECK();
    KPROF("node_bwd", stream, STRIVE_CUDA_LAUNCH(node_bwd_kernel, node_blocks, NODE_THREADS, SM_NODE_B, stream, M, a));
    STRIVE_LAUNCH_CHECK();
  }
  return 0;
}

// common part of the backward entry points: argument block on the forward tape, edge tile table
static int bwd_setup(const StriveModel* m, const StriveScene* sc, int32_t ft, const float* ext_future, void* tape, int64_t tape_bytes,
                     StepArgs& a, bool& edge_tc, cudaStream_t stream) {
  int rc = check_scene(m, sc, ft);
  if (rc) return rc;
  const int NA = sc->num_agents;
  const int64_t need = tape_carve(&a.tp, (char*)tape, NA, ft);
  STRIVE_CHECK(tape_bytes >= need, STRIVE_ESIZE, "tape too small: %lld < %lld", (long long)tape_bytes, (long long)need);
  rc = set_smem_attrs();
  if (rc) return rc;
  a.NA = NA; a.FT = ft; a.NC = sc->num_classes; a.t = 0;
  a.ptr = sc->ptr; a.scene_of = sc->scene_of; a.lw = sc->lw; a.sem = sc->sem; a.ext = ext_future;
  a.traj = nullptr; a.d_traj = nullptr; a.d_z = nullptr;
  a.z = a.tp.z;   // the latent the forward pass ran on
  edge_tc = g_edge_impl >= 3 && m->edge_frags != nullptr && sc->max_scene_agents <= ET_MAX_N;
  if (edge_tc) {     // the tile table depends on ptr only; rebuilt here so a backward call never relies on which forward kernel ran
    KPROF("edge_tiles", stream, edge_tc_tiles_kernel<<<1, 1024, 0, stream>>>(sc->ptr, sc->num_scenes, a.tp.et_tiles, a.tp.et_ntiles));
    STRIVE_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int strive_decode_bwd(const StriveModel* m, const StriveScene* sc, int32_t ft, const float* ext_future,
                                 const float* d_traj, float* d_z, void* tape, int64_t tape_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  StepArgs a;
  bool edge_tc = false;
  int rc = bwd_setup(m, sc, ft, ext_future, tape, tape_bytes, a, edge_tc, stream);
  if (rc) return rc;
  a.d_traj = d_traj; a.d_z = d_z;
  return bwd_sweep(m, sc, ft, model_dev(m), a, edge_tc, stream);
}

// Two adjoint sweeps of the same forward pass (the adv / sol loops route two different seeds to two groups of latents, SURVEY
// finding 3).  The sweeps share nothing but the read-only forward tape -- the second one runs on its own carry set -- so they are
// issued on two streams (fork / join by events: under stream capture they become two parallel branches of the graph).  Every
// per-agent backward kernel is a single partial wave on a latency floor, so the pair costs little more than one sweep.
extern "C" int strive_decode_bwd_pair(const StriveModel* m, const StriveScene* sc, int32_t ft, const float* ext_future,
                                      const float* d_traj_a, float* d_z_a, const float* d_traj_b, float* d_z_b,
                                      void* tape, int64_t tape_bytes, void* stream_) {
  STRIVE_CHECK(d_traj_a && d_z_a && d_traj_b && d_z_b && d_z_a != d_z_b, STRIVE_EINVAL, "strive_decode_bwd_pair: two seeds and two distinct outputs are required");
  cudaStream_t stream = (cudaStream_t)stream_;
  StepArgs a;
  bool edge_tc = false;
  int rc = bwd_setup(m, sc, ft, ext_future, tape, tape_bytes, a, edge_tc, stream);
  if (rc) return rc;
  cudaStream_t side;
  cudaEvent_t ev_fork, ev_join;
  rc = bwd_side_stream(&side, &ev_fork, &ev_join);
  if (rc) return rc;
  const ModelDev M = model_dev(m);
  StepArgs b = a;
  tape_use_alt(b.tp);
  a.d_traj = d_traj_a; a.d_z = d_z_a;
  b.d_traj = d_traj_b; b.d_z = d_z_b;
  STRIVE_CUDA(cudaEventRecord(ev_fork, stream));                 // after the tile table, which both sweeps read
  STRIVE_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
  rc = bwd_sweep(m, sc, ft, M, a, edge_tc, stream);
  const int rc_b = bwd_sweep(m, sc, ft, M, b, edge_tc, side);
  STRIVE_CUDA(cudaEventRecord(ev_join, side));                   // always joined: a capture must not end with a dangling branch
  STRIVE_CUDA(cudaStreamWaitEvent(stream, ev_join, 0));
  return rc ? rc : rc_b;
}

extern "C" int strive_decode_tape_read(const void* tape, int32_t num_agents, int32_t ft, const char* name, int32_t t,
                                       float* out, void* stream_) {
  Tape tp;
  tape_carve(&tp, (char*)tape, num_agents, ft);
  STRIVE_CHECK(t >= 0 && t < ft, STRIVE_EINVAL, "tape_read: step %d out of range", t);
  const size_t n = (size_t)num_agents;
  if (name[0] == 'a' && name[1] == 'r' && name[2] == 'g' && name[3] == 0) {   // arg-max routing: (NA,64) uint8, raw bytes
    STRIVE_CUDA(cudaMemcpyAsync(out, tp.arg + (size_t)t * n * 64, n * 64, cudaMemcpyDeviceToDevice, (cudaStream_t)stream_));
    return 0;
  }
  const float* src = nullptr;
  size_t w = 0;
  struct { const char* nm; const float* base; size_t width; } tab[] = {
      {"x", tp.x, 64}, {"P", tp.P, 128}, {"Q", tp.Q, 128}, {"aggr", tp.aggr, 64}, {"past_feat", tp.pastfeat, 64},
      {"map_feat", tp.mapfeat, 64}, {"prev", tp.prev, 6}, {"pos", tp.pos, 4}, {"loc", tp.loc, 4}, {"mem", tp.mem, 192}};
  for (auto& e : tab) {
    bool eq = true;
    for (int i = 0;; i++) {
      if (e.nm[i] != name[i]) { eq = false; break; }
      if (e.nm[i] == 0) break;
    }
    if (eq) { src = e.base + (size_t)t * n * e.width; w = e.width; }
  }
  STRIVE_CHECK(src != nullptr, STRIVE_EINVAL, "tape_read: unknown tensor '%s'", name);
  STRIVE_CUDA(cudaMemcpyAsync(out, src, n * w * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream_));
  return 0;
}
