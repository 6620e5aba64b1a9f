// Shared device helpers for the strive_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/strive_b200.h"

// ------------------------------------------------------------------------------------------------------
// error plumbing (api.cu owns the buffer)
// ------------------------------------------------------------------------------------------------------
void strive_set_error(const char* fmt, ...);

#define STRIVE_CHECK(cond, code, ...)                          \
  do {                                                         \
    if (!(cond)) {                                             \
      strive_set_error(__VA_ARGS__);                           \
      return (code);                                           \
    }                                                          \
  } while (0)

#define STRIVE_CUDA(call)                                                                      \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      strive_set_error("%s:%d CUDA error %d (%s) in %s", __FILE__, __LINE__, (int)_e,          \
                       cudaGetErrorString(_e), #call);                                         \
      return 100 + (int)_e;                                                                    \
    }                                                                                          \
  } while (0)

#define STRIVE_LAUNCH_CHECK() STRIVE_CUDA(cudaGetLastError())

// optional per-launch CUDA-event timing (strive_profile_enable); zero overhead when disabled
extern int g_strive_profile_on;
void strive_prof_begin(const char* name, cudaStream_t s);
void strive_prof_end(cudaStream_t s);
#define KPROF(name, stream, ...)                              \
  do {                                                        \
    if (g_strive_profile_on) strive_prof_begin(name, stream); \
    __VA_ARGS__;                                              \
    if (g_strive_profile_on) strive_prof_end(stream);         \
  } while (0)

// ------------------------------------------------------------------------------------------------------
// programmatic dependent launch: a kernel launched through strive_launch may start (and run its prologue: weights ->
// shared memory, barrier init, TMEM alloc) while its predecessor in the stream drains; STRIVE_PDL_WAIT() blocks until the
// predecessor grid has completed and its writes are visible, so everything after it is ordered exactly as a normal launch.
// Pre-wait code only reads model weights and touches its own shared memory / TMEM.  Both instructions are no-ops in a
// kernel that was launched without the attribute (<<<>>>, profiling mode).
// ------------------------------------------------------------------------------------------------------
extern int g_strive_pdl;
#define STRIVE_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define STRIVE_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
// wait placed after a prologue: the pointers to the predecessor's output pass THROUGH the asm, so no load from them (not
// even a read-only ld.global.nc, which the compiler may otherwise move across a "memory" clobber) can be scheduled above it.
// (Rollout kernels whose first loads were builtin __ldg() reads produced wrong results with the wait behind a prologue.)
#define STRIVE_PDL_WAIT_PTRS(p, q) asm volatile("griddepcontrol.wait;" : "+l"(p), "+l"(q)::"memory")

template <typename... KArgs, typename... Args>
inline cudaError_t strive_launch(int pdl_class, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = ((g_strive_pdl & pdl_class) != 0 && g_strive_profile_on == 0) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

#define STRIVE_CUDA_LAUNCH(kern, grid, block, smem, stream, ...) (void)strive_launch(STRIVE_PDL_CLASS, kern, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)

// Function attributes (opt-in dynamic shared memory) belong to a device, not to the process: every launcher keeps a bitmask of the
// devices it has configured (one process per GPU is the intended use; a process that drives several devices must still work).
static inline bool strive_first_use_on_device(unsigned* done_mask, int* dev_out = nullptr) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) dev = 0;
  if (dev_out) *dev_out = dev;
  if (*done_mask & (1u << dev)) return false;
  *done_mask |= 1u << dev;
  return true;
}

enum StriveErr { STRIVE_OK = 0, STRIVE_EINVAL = 1, STRIVE_ESIZE = 2, STRIVE_EUNSUPPORTED = 3 };

// ------------------------------------------------------------------------------------------------------
// constants of the reference model
// ------------------------------------------------------------------------------------------------------
#define ZDIM 32
#define FEAT 64
#define HID 128
#define LN_EPS 1e-5f

// datasets/utils.py:121-140
__device__ __constant__ const float kStateMean[6] = {0.0f, 0.0f, 0.0f, 0.0f, 1.802009f, -0.000037f};
__device__ __constant__ const float kStateStd[6] = {15.0f, 15.0f, 1.0f, 1.0f, 3.507907f, 0.055684f};
#define ATT_MEAN_L 4.844294f
#define ATT_STD_L 1.084860f
#define A_MEAN 0.409074f
#define A_STD 1.045530f
#define DDH_MEAN 0.000046f
#define DDH_STD 0.075032f
#define BIKE_DT 0.5f
#define BIKE_MAXHDOT 6.283185307179586f
#define BIKE_MAXS 50.0f

// ------------------------------------------------------------------------------------------------------
// packed model: segment table (order shared with strive_b200/weights.py through strive_model_layout)
// ------------------------------------------------------------------------------------------------------
enum Seg {
  // map encoder: conv weights k-major [Cin*ks*ks][Cout], k=(c*ks+ky)*ks+kx; FC as a 2x2 "conv" [512][64]
  S_CW0, S_CB0, S_GG0, S_GB0, S_CW1, S_CB1, S_GG1, S_GB1, S_CW2, S_CB2, S_GG2, S_GB2,
  S_CW3, S_CB3, S_GG3, S_GB3, S_CW4, S_CB4, S_GG4, S_GB4, S_CW5, S_CB5, S_GG5, S_GB5,
  S_FCW, S_FCB,
  // decoder_net.mlp_in   (_T = [in][out] transposed, _N = native [out][in] (or a column slice of it))
  S_IN0_T, S_IN0_B, S_IN_LN1_G, S_IN_LN1_B, S_IN3_T, S_IN3_B, S_IN_LN4_G, S_IN_LN4_B, S_IN6_T, S_IN6_B,
  S_IN0_N_PF, S_IN0_N_Z, S_IN3_N, S_IN6_N,
  // decoder_net.msg.0.edge_mlp, first layer split by input block [x_i | x_j | sem_i | sem_j | rel]
  S_E0_T_XI, S_E0_T_XJ, S_E0_T_SEMI, S_E0_T_SEMJ, S_E0_T_REL, S_E0_B, S_E0_N_XI, S_E0_N_XJ,
  S_E_LN1_G, S_E_LN1_B, S_E3_T, S_E3_N, S_E3_B, S_E_LN4_G, S_E_LN4_B, S_E6_T, S_E6_N, S_E6_B,
  // decoder_net.msg.0.update_mlp, input [x | aggr | sem | pad]
  S_U0_T, S_U0_B, S_U_LN1_G, S_U_LN1_B, S_U3_T, S_U3_B, S_U0_N_X, S_U0_N_AGGR, S_U3_N,
  // decoder_net.mlp_out
  S_O0_T, S_O0_B, S_O_LN1_G, S_O_LN1_B, S_O3_T, S_O3_B, S_O_LN4_G, S_O_LN4_B, S_O6_N, S_O6_B, S_O0_N, S_O3_N,
  // decoder_memory (3-layer GRU, gate order r,z,n): GI_T [Kin][192], GH_T [64][192], native [192][Kin]/[192][64]
  S_GI_T0, S_GH_T0, S_GBI0, S_GBH0, S_GI_N0, S_GH_N0,
  S_GI_T1, S_GH_T1, S_GBI1, S_GBH1, S_GI_N1, S_GH_N1,
  S_GI_T2, S_GH_T2, S_GBI2, S_GBH2, S_GI_N2, S_GH_N2,
  S_COUNT
};

struct StriveModel {
  const float* seg[S_COUNT];
  int64_t seg_size[S_COUNT];
  int nc;         // number of semantic classes
  int in0_rows;   // rows of IN0_T (= 64+64+NC+32+2 rounded up to 4)
  int u0_rows;    // rows of U0_T  (= 64+64+NC rounded up to 4)
  const uint8_t* tc_blob;   // bf16 hi/lo conv weights in UMMA canonical layout (strive_model_set_tc_weights) or null
  int64_t tc_off[8];        // byte offsets of conv1..conv6, fc, conv3-for-CTA-pairs inside tc_blob
  float h_cbias[4][64];     // host copies of the conv1..conv4 biases (kernel arguments of the tensor-core convolutions)
  const uint8_t* edge_frags; // mma.sync weight fragment packs of the edge MLP (strive_model_set_edge_frags) or null
};

__host__ __device__ inline int round_up4(int x) { return (x + 3) & ~3; }

// ------------------------------------------------------------------------------------------------------
// warp helpers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int V>
__device__ __forceinline__ void ldvec(float (&w)[V], const float* __restrict__ p) {
  if constexpr (V == 4) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
  } else if constexpr (V == 2) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    w[0] = t.x; w[1] = t.y;
  } else {
    w[0] = __ldg(p);
  }
}

template <int V>
__device__ __forceinline__ void stvec(float* p, const float (&w)[V]) {
  if constexpr (V == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(w[0], w[1], w[2], w[3]);
  } else if constexpr (V == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(w[0], w[1]);
  } else {
    p[0] = w[0];
  }
}

// acc[r][v] += sum_k xs[r*ldx + k] * Wm[k*OUT + lane*V + v],  V = OUT/32.   red % 4 == 0, ldx % 4 == 0.
// Wm is a global [red][OUT] matrix (coalesced across lanes, L1/L2 resident); xs is per-warp shared memory
// (broadcast reads).  R rows share every weight load.
#ifndef WG_UNROLL
#define WG_UNROLL 4
#endif
constexpr int kWgUnroll = WG_UNROLL;   // k-steps of 4 in flight per warp in the GEMV loops (weight loads are L2-latency bound)
template <int OUT, int R>
__device__ __forceinline__ void warp_gemm(const float* __restrict__ Wm, int red, const float* xs, int ldx,
                                          float (&acc)[R][OUT / 32], int lane) {
  constexpr int V = OUT / 32;
  const float* wp = Wm + lane * V;
#pragma unroll kWgUnroll
  for (int k = 0; k < red; k += 4) {
    float w[4][V];
#pragma unroll
    for (int kk = 0; kk < 4; kk++) ldvec<V>(w[kk], wp + (size_t)(k + kk) * OUT);
#pragma unroll
    for (int r = 0; r < R; r++) {
      const float4 xv = *reinterpret_cast<const float4*>(xs + r * ldx + k);
#pragma unroll
      for (int v = 0; v < V; v++) {
        acc[r][v] = fmaf(xv.x, w[0][v], acc[r][v]);
        acc[r][v] = fmaf(xv.y, w[1][v], acc[r][v]);
        acc[r][v] = fmaf(xv.z, w[2][v], acc[r][v]);
        acc[r][v] = fmaf(xv.w, w[3][v], acc[r][v]);
      }
    }
  }
}

// GRU gate GEMM: Wm [red][192], lane owns columns g*64 + lane*2 + {0,1} for g = 0,1,2 (r,z,n), so that all three
// gates of a hidden unit live in one lane.  acc[r][g*2 + e].
template <int R>
__device__ __forceinline__ void warp_gemm_gru(const float* __restrict__ Wm, int red, const float* xs, int ldx,
                                              float (&acc)[R][6], int lane) {
  const float* wp = Wm + lane * 2;
#pragma unroll kWgUnroll
  for (int k = 0; k < red; k += 4) {
    float w[4][6];
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
#pragma unroll
      for (int g = 0; g < 3; g++) {
        float2 t = __ldg(reinterpret_cast<const float2*>(wp + (size_t)(k + kk) * 192 + g * 64));
        w[kk][g * 2] = t.x;
        w[kk][g * 2 + 1] = t.y;
      }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
      const float4 xv = *reinterpret_cast<const float4*>(xs + r * ldx + k);
#pragma unroll
      for (int v = 0; v < 6; v++) {
        acc[r][v] = fmaf(xv.x, w[0][v], acc[r][v]);
        acc[r][v] = fmaf(xv.y, w[1][v], acc[r][v]);
        acc[r][v] = fmaf(xv.z, w[2][v], acc[r][v]);
        acc[r][v] = fmaf(xv.w, w[3][v], acc[r][v]);
      }
    }
  }
}

template <int V, int R>
__device__ __forceinline__ void init_bias(float (&acc)[R][V], const float* __restrict__ bias, int lane) {
  float b[V];
  ldvec<V>(b, bias + lane * V);
#pragma unroll
  for (int r = 0; r < R; r++)
#pragma unroll
    for (int v = 0; v < V; v++) acc[r][v] = b[v];
}

template <int V, int R>
__device__ __forceinline__ void init_zero(float (&acc)[R][V]) {
#pragma unroll
  for (int r = 0; r < R; r++)
#pragma unroll
    for (int v = 0; v < V; v++) acc[r][v] = 0.f;
}

// rows of a 128-wide activation held as acc[r][4] (lane owns cols lane*4..+3):
// y = relu(LayerNorm(a) * g + b), written to ys (shared, row stride ldy). Optionally keeps the pre-LN `a` in apre.
template <int R>
__device__ __forceinline__ void ln_relu_store(const float (&acc)[R][4], const float* __restrict__ gam,
                                              const float* __restrict__ bet, float* ys, int ldy, float* apre,
                                              int ldp, int lane) {
  float g[4], b[4];
  ldvec<4>(g, gam + lane * 4);
  ldvec<4>(b, bet + lane * 4);
#pragma unroll
  for (int r = 0; r < R; r++) {
    float s = acc[r][0] + acc[r][1] + acc[r][2] + acc[r][3];
    const float mean = warp_sum(s) * (1.0f / 128.0f);
    float d[4];
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < 4; v++) {
      d[v] = acc[r][v] - mean;
      q = fmaf(d[v], d[v], q);
    }
    const float var = warp_sum(q) * (1.0f / 128.0f);
    const float rstd = 1.0f / sqrtf(var + LN_EPS);
    float y[4];
#pragma unroll
    for (int v = 0; v < 4; v++) y[v] = fmaxf(fmaf(d[v] * rstd, g[v], b[v]), 0.f);
    stvec<4>(ys + r * ldy + lane * 4, y);
    if (apre != nullptr) stvec<4>(apre + r * ldp + lane * 4, acc[r]);
  }
}

// backward of h = relu(LN(a)*g+b) for 128-wide rows.  dh (regs, lane cols) -> da (regs). `apre` holds pre-LN a.
template <int R>
__device__ __forceinline__ void ln_relu_bwd(float (&dh)[R][4], const float* apre, int ldp,
                                            const float* __restrict__ gam, const float* __restrict__ bet, int lane) {
  float g[4], b[4];
  ldvec<4>(g, gam + lane * 4);
  ldvec<4>(b, bet + lane * 4);
#pragma unroll
  for (int r = 0; r < R; r++) {
    const float4 av = *reinterpret_cast<const float4*>(apre + r * ldp + lane * 4);
    float a[4] = {av.x, av.y, av.z, av.w};
    const float mean = warp_sum(a[0] + a[1] + a[2] + a[3]) * (1.0f / 128.0f);
    float d[4];
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < 4; v++) {
      d[v] = a[v] - mean;
      q = fmaf(d[v], d[v], q);
    }
    const float var = warp_sum(q) * (1.0f / 128.0f);
    const float rstd = 1.0f / sqrtf(var + LN_EPS);
    float xh[4], dg[4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < 4; v++) {
      xh[v] = d[v] * rstd;
      const float y = fmaf(xh[v], g[v], b[v]);
      const float dy = (y > 0.f) ? dh[r][v] : 0.f;
      dg[v] = dy * g[v];
      s1 += dg[v];
      s2 = fmaf(dg[v], xh[v], s2);
    }
    const float m1 = warp_sum(s1) * (1.0f / 128.0f);
    const float m2 = warp_sum(s2) * (1.0f / 128.0f);
#pragma unroll
    for (int v = 0; v < 4; v++) dh[r][v] = rstd * (dg[v] - m1 - xh[v] * m2);
  }
}

template <int V, int R>
__device__ __forceinline__ void store_rows(const float (&acc)[R][V], float* ys, int ldy, int lane) {
#pragma unroll
  for (int r = 0; r < R; r++) stvec<V>(ys + r * ldy + lane * V, acc[r]);
}

// ------------------------------------------------------------------------------------------------------
// geometry: utils/transforms.py:78-139 transform2frame, non-inverse; frame f=(fx,fy,c,s), pose p=(px,py,pc,ps)
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void t2f_fwd(const float f[4], const float p[4], float out[4]) {
  const float dx = p[0] - f[0], dy = p[1] - f[1];
  out[0] = f[2] * dx + f[3] * dy;
  out[1] = -f[3] * dx + f[2] * dy;
  out[2] = p[2] * f[2] + p[3] * f[3];
  out[3] = p[3] * f[2] - p[2] * f[3];
}

// accumulates into df, dp
__device__ __forceinline__ void t2f_bwd(const float f[4], const float p[4], const float g[4], float df[4], float dp[4]) {
  const float dx = p[0] - f[0], dy = p[1] - f[1];
  const float ddx = f[2] * g[0] - f[3] * g[1];
  const float ddy = f[3] * g[0] + f[2] * g[1];
  dp[0] += ddx;
  dp[1] += ddy;
  df[0] -= ddx;
  df[1] -= ddy;
  df[2] += dx * g[0] + dy * g[1] + p[2] * g[2] + p[3] * g[3];
  df[3] += dy * g[0] - dx * g[1] + p[3] * g[2] - p[2] * g[3];
  dp[2] += f[2] * g[2] - f[3] * g[3];
  dp[3] += f[3] * g[2] + f[2] * g[3];
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
