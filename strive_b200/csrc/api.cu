// C-ABI glue: error string, model handle (packed weights), fused Adam.
#include <stdarg.h>
#include <string.h>
#include <new>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void strive_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* strive_last_error(void) { return g_err; }
extern "C" int strive_abi_version(void) { return STRIVE_ABI_VERSION; }

// ------------------------------------------------------------------------------------------------------
// segment sizes (floats); the order is enum Seg in common.cuh
// ------------------------------------------------------------------------------------------------------
static void seg_sizes(int nc, int64_t* sz) {
  const int chans[7] = {4, 16, 32, 64, 64, 128, 128};
  const int ks[6] = {7, 5, 5, 3, 3, 3};
  for (int l = 0; l < 6; l++) {
    sz[S_CW0 + 4 * l] = (int64_t)chans[l] * ks[l] * ks[l] * chans[l + 1];
    sz[S_CB0 + 4 * l] = chans[l + 1];
    sz[S_GG0 + 4 * l] = chans[l + 1];
    sz[S_GB0 + 4 * l] = chans[l + 1];
  }
  sz[S_FCW] = 512 * 64; sz[S_FCB] = 64;
  const int in0 = round_up4(64 + 64 + nc + ZDIM + 2);
  const int u0 = round_up4(64 + 64 + nc);
  sz[S_IN0_T] = (int64_t)in0 * 128; sz[S_IN0_B] = 128; sz[S_IN_LN1_G] = 128; sz[S_IN_LN1_B] = 128;
  sz[S_IN3_T] = 128 * 128; sz[S_IN3_B] = 128; sz[S_IN_LN4_G] = 128; sz[S_IN_LN4_B] = 128;
  sz[S_IN6_T] = 128 * 64; sz[S_IN6_B] = 64;
  sz[S_IN0_N_PF] = 128 * 64; sz[S_IN0_N_Z] = 128 * 32; sz[S_IN3_N] = 128 * 128; sz[S_IN6_N] = 64 * 128;
  sz[S_E0_T_XI] = 64 * 128; sz[S_E0_T_XJ] = 64 * 128; sz[S_E0_T_SEMI] = nc * 128; sz[S_E0_T_SEMJ] = nc * 128;
  sz[S_E0_T_REL] = 4 * 128; sz[S_E0_B] = 128; sz[S_E0_N_XI] = 128 * 64; sz[S_E0_N_XJ] = 128 * 64;
  sz[S_E_LN1_G] = 128; sz[S_E_LN1_B] = 128; sz[S_E3_T] = 128 * 128; sz[S_E3_N] = 128 * 128; sz[S_E3_B] = 128;
  sz[S_E_LN4_G] = 128; sz[S_E_LN4_B] = 128; sz[S_E6_T] = 128 * 64; sz[S_E6_N] = 64 * 128; sz[S_E6_B] = 64;
  sz[S_U0_T] = (int64_t)u0 * 128; sz[S_U0_B] = 128; sz[S_U_LN1_G] = 128; sz[S_U_LN1_B] = 128;
  sz[S_U3_T] = 128 * 64; sz[S_U3_B] = 64; sz[S_U0_N_X] = 128 * 64; sz[S_U0_N_AGGR] = 128 * 64; sz[S_U3_N] = 64 * 128;
  sz[S_O0_T] = 64 * 128; sz[S_O0_B] = 128; sz[S_O_LN1_G] = 128; sz[S_O_LN1_B] = 128;
  sz[S_O3_T] = 128 * 128; sz[S_O3_B] = 128; sz[S_O_LN4_G] = 128; sz[S_O_LN4_B] = 128;
  sz[S_O6_N] = 2 * 128; sz[S_O6_B] = 2; sz[S_O0_N] = 128 * 64; sz[S_O3_N] = 128 * 128;
  for (int l = 0; l < 3; l++) {
    const int kin = (l == 0) ? 4 : 64;
    sz[S_GI_T0 + 6 * l] = (int64_t)kin * 192; sz[S_GH_T0 + 6 * l] = 64 * 192;
    sz[S_GBI0 + 6 * l] = 192; sz[S_GBH0 + 6 * l] = 192;
    sz[S_GI_N0 + 6 * l] = (int64_t)192 * kin; sz[S_GH_N0 + 6 * l] = 192 * 64;
  }
}

int g_strive_pdl = 3;      // bit 0: rollout kernels, bit 1: map-encoder kernels
extern "C" int strive_set_pdl(int on) {
  g_strive_pdl = on;
  return 0;
}

extern "C" int strive_model_layout(int num_classes, int64_t* seg_sizes_host, int max_segs, int* n_segs_out) {
  STRIVE_CHECK(num_classes >= 1 && num_classes <= 16, STRIVE_EINVAL, "num_classes=%d out of range", num_classes);
  STRIVE_CHECK(seg_sizes_host && n_segs_out && max_segs >= (int)S_COUNT, STRIVE_EINVAL, "layout buffer too small (%d < %d)", max_segs, (int)S_COUNT);
  seg_sizes(num_classes, seg_sizes_host);
  *n_segs_out = S_COUNT;
  return 0;
}

extern "C" int strive_model_create(const float* blob, int64_t blob_floats, const int64_t* seg_sizes_host, int n_segs,
                                   int num_classes, StriveModel** out) {
  STRIVE_CHECK(blob && seg_sizes_host && out, STRIVE_EINVAL, "strive_model_create: null argument");
  STRIVE_CHECK(n_segs == (int)S_COUNT, STRIVE_ESIZE, "segment count %d != %d", n_segs, (int)S_COUNT);
  STRIVE_CHECK(num_classes >= 1 && num_classes <= 16, STRIVE_EINVAL, "num_classes=%d out of range", num_classes);
  int64_t want[S_COUNT];
  seg_sizes(num_classes, want);
  StriveModel* m = new (std::nothrow) StriveModel();
  STRIVE_CHECK(m != nullptr, STRIVE_EINVAL, "out of host memory");
  int64_t off = 0;
  for (int i = 0; i < S_COUNT; i++) {
    if (seg_sizes_host[i] != want[i]) {
      strive_set_error("segment %d has %lld floats, expected %lld", i, (long long)seg_sizes_host[i], (long long)want[i]);
      delete m;
      return STRIVE_ESIZE;
    }
    m->seg[i] = blob + off;
    m->seg_size[i] = want[i];
    off += (want[i] + 3) & ~(int64_t)3;   // every segment starts 16-byte aligned
  }
  if (off != blob_floats) {
    strive_set_error("weight blob has %lld floats, expected %lld", (long long)blob_floats, (long long)off);
    delete m;
    return STRIVE_ESIZE;
  }
  m->nc = num_classes;
  m->tc_blob = nullptr;
  m->edge_frags = nullptr;
  m->in0_rows = round_up4(64 + 64 + num_classes + ZDIM + 2);
  m->u0_rows = round_up4(64 + 64 + num_classes);
  *out = m;
  return 0;
}

extern "C" void strive_model_destroy(StriveModel* m) { delete m; }

// tensor-core weight blob: [conv1 10752 B int8 digit planes + 16 fp32 scales][conv2 51200 B][conv3 2 x 102400 B][conv4 147456 B][conv5][conv6][fc]
// [conv3 for CTA pairs: 2 ranks x 2 K chunks x 25 taps x 2048 B]  (layouts in mapenc_tc.cu)
#define TC_SEGS 8
static const int64_t kTcBytes[TC_SEGS] = {7 * 2 * 48 * 16 + 64, 1 * (1 * 25 * 2 * 1024), 2 * (2 * 25 * 2 * 1024), 2 * (4 * 9 * 2 * 1024),
                                          9 * 2 * 128 * 64 * 2, 18 * 2 * 128 * 64 * 2, 8 * 2 * 64 * 64 * 2, 2 * 2 * 25 * 2048};
extern "C" int64_t strive_model_tc_bytes(void) {
  int64_t t = 0;
  for (int i = 0; i < TC_SEGS; i++) t += kTcBytes[i];
  return t;
}
extern "C" int strive_model_set_tc_weights(StriveModel* m, const void* blob, int64_t bytes) {
  STRIVE_CHECK(m != nullptr, STRIVE_EINVAL, "null model");
  STRIVE_CHECK(blob != nullptr && bytes == strive_model_tc_bytes(), STRIVE_ESIZE, "tc weight blob has %lld bytes, expected %lld", (long long)bytes,
               (long long)strive_model_tc_bytes());
  STRIVE_CHECK(((uintptr_t)blob & 15) == 0, STRIVE_EINVAL, "tc weight blob must be 16-byte aligned");
  // conv1..conv4 biases travel as kernel arguments (constant bank): keep host copies (one-time synchronous copy)
  static const int kBiasSeg[4] = {S_CB0, S_CB1, S_CB2, S_CB3};
  static const int kBiasN[4] = {16, 32, 64, 64};
  for (int l = 0; l < 4; l++) {
    STRIVE_CHECK(m->seg_size[kBiasSeg[l]] == kBiasN[l], STRIVE_ESIZE, "conv%d bias has %lld entries", l + 1, (long long)m->seg_size[kBiasSeg[l]]);
    for (int i = 0; i < 64; i++) m->h_cbias[l][i] = 0.f;
    STRIVE_CUDA(cudaMemcpy(m->h_cbias[l], m->seg[kBiasSeg[l]], sizeof(float) * kBiasN[l], cudaMemcpyDeviceToHost));
  }
  m->tc_blob = (const uint8_t*)blob;
  int64_t off = 0;
  for (int i = 0; i < TC_SEGS; i++) { m->tc_off[i] = off; off += kTcBytes[i]; }
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam, amsgrad=False, weight_decay=0, maximize=False) as used by the latent loops
// (refine_traffic_optim.py:166, init_optim.py:21, adv_gen_optim.py:72, sol_optim.py:47)
// ------------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ z, const float* __restrict__ ga, const float* __restrict__ gb,
                            float* __restrict__ m, float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                            float bc1, float bc2_sqrt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float g = ga[i];
  if (gb != nullptr) g += gb[i];
  const float mi = m[i] + (g - m[i]) * (1.0f - b1);          // exp_avg.lerp_(grad, 1-beta1)
  const float vi = v[i] * b2 + (1.0f - b2) * g * g;          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  z[i] = z[i] - (lr / bc1) * (mi / denom);
}

extern "C" int strive_adam_step(float* z, const float* g_a, const float* g_b, float* exp_avg, float* exp_avg_sq, int64_t n,
                                int32_t step_count, float lr, float beta1, float beta2, float eps, void* stream) {
  STRIVE_CHECK(z && g_a && exp_avg && exp_avg_sq && n > 0 && step_count >= 1, STRIVE_EINVAL, "strive_adam_step: bad arguments");
  const double bc1 = 1.0 - pow((double)beta1, (double)step_count);
  const double bc2 = 1.0 - pow((double)beta2, (double)step_count);
  KPROF("adam", (cudaStream_t)stream, adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(z, g_a, g_b, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                             (float)bc1, (float)sqrt(bc2)));
  STRIVE_LAUNCH_CHECK();
  return 0;
}

// Device-resident variant for the fused loops: the step counter lives in device memory (so one captured CUDA graph of an
// iteration can be replayed: no per-step host argument changes), and the gradient of a row is taken from one of two adjoint
// sweeps -- adv / sol loops run two sweeps over one rollout tape (adv_gen_optim.py:119-130, sol_optim.py:73-77): row_sel[r] != 0
// selects g_a, else g_b; g_direct (latent terms written by the loss kernel, zero outside its row mask) is added on top.
__global__ void adam_dev_kernel(float* __restrict__ z, const float* __restrict__ ga, const float* __restrict__ gb,
                                const float* __restrict__ gd, const uint8_t* __restrict__ row_sel, int width, float* __restrict__ m,
                                float* __restrict__ v, int64_t n, const int32_t* __restrict__ step_dev, float lr, float b1, float b2,
                                float eps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int step = *step_dev + 1;
  const float bc1 = (float)(1.0 - pow((double)b1, (double)step));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, (double)step));
  float g = (row_sel == nullptr || row_sel[i / width] != 0) ? ga[i] : gb[i];
  if (gd != nullptr) g += gd[i];
  const float mi = m[i] + (g - m[i]) * (1.0f - b1);
  const float vi = v[i] * b2 + (1.0f - b2) * g * g;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  z[i] = z[i] - (lr / bc1) * (mi / denom);
}
__global__ void step_inc_kernel(int32_t* s) { *s += 1; }

extern "C" int strive_adam_step_dev(float* z, const float* g_a, const float* g_b, const float* g_direct, const uint8_t* row_sel,
                                    int32_t row_width, float* exp_avg, float* exp_avg_sq, int64_t n, int32_t* step_dev, float lr,
                                    float beta1, float beta2, float eps, void* stream) {
  STRIVE_CHECK(z && g_a && exp_avg && exp_avg_sq && step_dev && n > 0 && row_width > 0, STRIVE_EINVAL, "strive_adam_step_dev: bad arguments");
  STRIVE_CHECK(row_sel == nullptr || g_b != nullptr, STRIVE_EINVAL, "strive_adam_step_dev: row_sel needs g_b");
  KPROF("adam", (cudaStream_t)stream, adam_dev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      z, g_a, g_b, g_direct, row_sel, row_width, exp_avg, exp_avg_sq, n, step_dev, lr, beta1, beta2, eps));
  STRIVE_LAUNCH_CHECK();
  step_inc_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
  STRIVE_LAUNCH_CHECK();
  return 0;
}

// layout self-check for foreign-function bindings (tests compare against ctypes.sizeof / offsets)
#include <stddef.h>
extern "C" int strive_struct_layout(int64_t* out, int max_n) {
  const int64_t v[] = {(int64_t)sizeof(StriveScene), (int64_t)offsetof(StriveScene, ptr), (int64_t)offsetof(StriveScene, sem),
                       (int64_t)sizeof(StriveMap), (int64_t)offsetof(StriveMap, lin_l),
                       (int64_t)sizeof(StriveLossCfg), (int64_t)offsetof(StriveLossCfg, group_agent_ptr),
                       (int64_t)offsetof(StriveLossCfg, w_coll_veh), (int64_t)offsetof(StriveLossCfg, attack_mask),
                       (int64_t)offsetof(StriveLossCfg, lw_un)};
  const int n = (int)(sizeof(v) / sizeof(v[0]));
  if (max_n < n) return -1;
  for (int i = 0; i < n; i++) out[i] = v[i];
  return n;
}

// ------------------------------------------------------------------------------------------------------
// per-kernel device timing with CUDA events on the launching stream (bench.py roofline numbers)
// ------------------------------------------------------------------------------------------------------
#include <vector>
#include <string>
#include <map>
int g_strive_profile_on = 0;
struct ProfRec { const char* name; cudaEvent_t e0, e1; };
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_pool;
static cudaEvent_t prof_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
void strive_prof_begin(const char* name, cudaStream_t s) {
  ProfRec r;
  r.name = name; r.e0 = prof_event(); r.e1 = prof_event();
  cudaEventRecord(r.e0, s);
  g_prof.push_back(r);
}
void strive_prof_end(cudaStream_t s) { cudaEventRecord(g_prof.back().e1, s); }

extern "C" int strive_profile_enable(int on) {
  g_strive_profile_on = on ? 1 : 0;
  return 0;
}
// Synchronises the device, aggregates "name count total_ms" lines into buf, clears the records. Returns bytes written.
extern "C" int64_t strive_profile_report(char* buf, int64_t cap) {
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<long long, double>> agg;
  for (auto& r : g_prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    auto& a = agg[r.name];
    a.first += 1;
    a.second += ms;
    g_pool.push_back(r.e0);
    g_pool.push_back(r.e1);
  }
  g_prof.clear();
  std::string out;
  for (auto& kv : agg) {
    char line[256];
    snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  if ((int64_t)out.size() + 1 > cap) return -1;
  memcpy(buf, out.c_str(), out.size() + 1);
  return (int64_t)out.size();
}
