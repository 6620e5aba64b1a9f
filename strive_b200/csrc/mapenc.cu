// Map encoder: rotated nearest-neighbour crop of the uint8 raster fused into conv1, then 5 more
// [Conv2d(stride 2, pad 0) -> GroupNorm(1,C) -> ReLU] layers and the final Linear.
//
// Restates reference src/models/traffic_model.py:416-451 (encode_map), :69-87 (map_conv/map_feature),
// src/datasets/map_env.py:168-203 (get_map_crop), src/datasets/nuscenes_utils.py:205-264 (gen_car_coords,
// get_map_obs).  The 4x256x256 crop (256 KB/agent as uint8, 5 MB/agent as the reference's int64 index + float
// tensors) is never materialised: conv1 gathers its input tile straight from the raster.
//
// GroupNorm(1,C) needs whole-sample statistics, so each layer writes its RAW conv output plus per-crop
// (sum, sumsq) in fp64, and the NEXT layer applies normalise+affine+ReLU while staging its input tile.
//
// v0 = fp32 SIMT direct convolution (bit-for-bit crop, fp32 accumulate).  See DESIGN.md for the tensor-core plan.
#include "common.cuh"

#define CROP 256

// 1 = tensor-core conv1..fc (default), 0 = fp32 SIMT reference kernels (kept for A/B verification, strive_mapenc_set_impl)
static int g_mapenc_impl = 1;

// ------------------------------------------------------------------------------------------------------
// crop sampling: exact restatement of get_map_obs (nuscenes_utils.py:248-263)
// ------------------------------------------------------------------------------------------------------
struct CropFrame {
  float px, py, hc, hs;
  double dx0, dx1;
  const uint8_t* base;   // raster + m*C*H*W
};

__device__ __forceinline__ void crop_pixel(const CropFrame& f, float l, float w, int H, int W, long long& xp, long long& yp) {
  // gen_car_coords (:232-233): (l*hcos - w*hsin) + x ; (l*hsin + w*hcos) + y   -- separate fp32 roundings, no FMA
  float gx = __fadd_rn(__fsub_rn(__fmul_rn(l, f.hc), __fmul_rn(w, f.hs)), f.px);
  float gy = __fadd_rn(__fadd_rn(__fmul_rn(l, f.hs), __fmul_rn(w, f.hc)), f.py);
  if (isnan(gx)) gx = 0.f;   // :251
  if (isnan(gy)) gy = 0.f;
  // :254-255  float32 / float64 -> float64, torch.round = half-to-even; x uses dx[:,0], y uses dx[:,1]
  const double qx = rint((double)gx / f.dx0);
  const double qy = rint((double)gy / f.dx1);
  xp = (long long)qx;
  yp = (long long)qy;
  if (yp < 0 || yp >= H || xp < 0 || xp >= W) { xp = 0; yp = 0; }   // :260-262
}

int tc_crop_pack_unpacked(const StriveMap* map, const float* pose, const int32_t* map_of, int n, uint8_t* out, cudaStream_t stream);

__global__ void map_crop_kernel(StriveMap map, const float* __restrict__ pose, const int32_t* __restrict__ map_of, int n,
                                uint8_t* __restrict__ out) {
  const int crop = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over 256*256
  if (crop >= n || idx >= CROP * CROP) return;
  const int iy = idx / CROP, ix = idx % CROP;
  const int m = map_of[crop];
  CropFrame f;
  f.px = pose[crop * 4 + 0]; f.py = pose[crop * 4 + 1]; f.hc = pose[crop * 4 + 2]; f.hs = pose[crop * 4 + 3];
  f.dx0 = map.dx[m * 2 + 0]; f.dx1 = map.dx[m * 2 + 1];
  f.base = map.raster + (size_t)m * map.C * map.H * map.W;
  long long xp, yp;
  crop_pixel(f, map.lin_l[iy], map.lin_w[ix], map.H, map.W, xp, yp);
  for (int c = 0; c < map.C; c++)
    out[(((size_t)crop * map.C + c) * CROP + iy) * CROP + ix] = f.base[((size_t)c * map.H + yp) * map.W + xp];
}

extern "C" int strive_map_crop(const StriveMap* map, const float* pose_un, const int32_t* map_of, int32_t n,
                               uint8_t* out_crop, void* stream) {
  STRIVE_CHECK(map && pose_un && map_of && out_crop && n > 0, STRIVE_EINVAL, "strive_map_crop: bad arguments");
  // impl 1: the production crop_pack kernel of the tensor-core encoder (unpacked); impl 0: straight float64-division restatement
  if (g_mapenc_impl == 1 && map->packed != nullptr) return tc_crop_pack_unpacked(map, pose_un, map_of, n, out_crop, (cudaStream_t)stream);
  dim3 grid((CROP * CROP + 255) / 256, n);
  KPROF("map_crop", (cudaStream_t)stream, map_crop_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*map, pose_un, map_of, n, out_crop));
  STRIVE_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// conv1: 4 -> 16, k7, s2, 256 -> 125, input gathered from the raster (binary), output raw + stats
// tile = 32x32 outputs, 256 threads, each thread 2x2 outputs x 16 channels
// ------------------------------------------------------------------------------------------------------
#define C1_T 32
#define C1_PH (C1_T * 2 + 5)     // 69 input rows/cols per tile
#define C1_PITCH 80              // bytes; 4 rows * 80 B = 80 words = 16 mod 32 -> the two half-warps hit disjoint banks
#define C1_OUT 125
#define C1_TILES 4

__global__ void __launch_bounds__(256) conv1_gather_kernel(StriveMap map, const float* __restrict__ pose,
                                                           const int32_t* __restrict__ map_of, const float* __restrict__ Wk,
                                                           const float* __restrict__ bias, float* __restrict__ out,
                                                           double* __restrict__ out_stats, int n) {
  __shared__ __align__(16) uint8_t patch[4][C1_PH][C1_PITCH];
  __shared__ __align__(16) float wsm[196 * 16];
  __shared__ float red[2][8];
  const int crop = blockIdx.y;
  const int tile = blockIdx.x;
  const int oy0 = (tile / C1_TILES) * C1_T, ox0 = (tile % C1_TILES) * C1_T;
  const int tid = threadIdx.x;
  for (int i = tid; i < 196 * 16; i += 256) wsm[i] = __ldg(Wk + i);
  {
    const int m = map_of[crop];
    CropFrame f;
    f.px = pose[crop * 4 + 0]; f.py = pose[crop * 4 + 1]; f.hc = pose[crop * 4 + 2]; f.hs = pose[crop * 4 + 3];
    f.dx0 = map.dx[m * 2 + 0]; f.dx1 = map.dx[m * 2 + 1];
    f.base = map.raster + (size_t)m * 4 * map.H * map.W;
    const size_t plane = (size_t)map.H * map.W;
    for (int i = tid; i < C1_PH * C1_PH; i += 256) {
      const int r = i / C1_PH, cidx = i % C1_PH;
      const int iy = oy0 * 2 + r, ix = ox0 * 2 + cidx;
      uint8_t v0 = 0, v1 = 0, v2 = 0, v3 = 0;
      if (iy < CROP && ix < CROP) {
        long long xp, yp;
        crop_pixel(f, __ldg(map.lin_l + iy), __ldg(map.lin_w + ix), map.H, map.W, xp, yp);
        const uint8_t* p = f.base + (size_t)yp * map.W + xp;
        v0 = __ldg(p); v1 = __ldg(p + plane); v2 = __ldg(p + 2 * plane); v3 = __ldg(p + 3 * plane);
      }
      patch[0][r][cidx] = v0; patch[1][r][cidx] = v1; patch[2][r][cidx] = v2; patch[3][r][cidx] = v3;
    }
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][16];
#pragma unroll
  for (int p = 0; p < 4; p++)
#pragma unroll
    for (int c = 0; c < 16; c++) acc[p][c] = 0.f;
  const int ry = ty * 4, rx = tx * 4;   // input origin of this thread's 2x2 outputs
  for (int c = 0; c < 4; c++) {
#pragma unroll
    for (int ky = 0; ky < 7; ky++) {
#pragma unroll
      for (int kx = 0; kx < 7; kx++) {
        const float4* wp = reinterpret_cast<const float4*>(wsm + ((c * 7 + ky) * 7 + kx) * 16);
        const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
        const float w[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
        float v[4];
        v[0] = patch[c][ry + ky][rx + kx] ? 1.f : 0.f;
        v[1] = patch[c][ry + ky][rx + 2 + kx] ? 1.f : 0.f;
        v[2] = patch[c][ry + 2 + ky][rx + kx] ? 1.f : 0.f;
        v[3] = patch[c][ry + 2 + ky][rx + 2 + kx] ? 1.f : 0.f;
#pragma unroll
        for (int p = 0; p < 4; p++)
#pragma unroll
          for (int co = 0; co < 16; co++) acc[p][co] = fmaf(v[p], w[co], acc[p][co]);
      }
    }
  }
  float s1 = 0.f, s2 = 0.f;
  float* o = out + (size_t)crop * 16 * C1_OUT * C1_OUT;
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int oy = oy0 + ty * 2 + (p >> 1), ox = ox0 + tx * 2 + (p & 1);
    if (oy < C1_OUT && ox < C1_OUT) {
#pragma unroll
      for (int co = 0; co < 16; co++) {
        const float v = acc[p][co] + __ldg(bias + co);
        o[((size_t)co * C1_OUT + oy) * C1_OUT + ox] = v;
        s1 += v;
        s2 = fmaf(v, v, s2);
      }
    }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((tid & 31) == 0) { red[0][tid >> 5] = s1; red[1][tid >> 5] = s2; }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; w++) { a += red[0][w]; b += red[1][w]; }
    atomicAdd(out_stats + (size_t)crop * 2, a);
    atomicAdd(out_stats + (size_t)crop * 2 + 1, b);
  }
}

// ------------------------------------------------------------------------------------------------------
// generic stride-2 conv with GroupNorm+ReLU applied to the INPUT while staging (layers 2..6 and the FC,
// which is a 2x2 "conv" on the 128x2x2 activation).  Input / output NCHW fp32.
//   block = G crops x (TH x TW) output tile x COC output channels, thread = PXT pixels x COC channels.
//   shared: input patch, columns de-interleaved by parity so stride-2 reads are unit-stride across lanes.
// ------------------------------------------------------------------------------------------------------
template <int CIN, int COUT, int KS, int HIN, int HOUT, int TH, int TW, int G, int CC, int COC, int PXT, bool FINAL>
struct ConvCfg {
  static constexpr int S = 2;
  static constexpr int PH = (TH - 1) * S + KS;
  static constexpr int PW = (TW - 1) * S + KS;
  static constexpr int PWH = (PW + 1) / 2;
  static constexpr int TILES_X = (HOUT + TW - 1) / TW;
  static constexpr int TILES = TILES_X * TILES_X;
  static constexpr int NTHREADS = G * TH * TW / PXT;
  static constexpr int PATCH_FLOATS = G * CC * PH * 2 * PWH;
  static constexpr int W_FLOATS = CC * KS * KS * COC;
  static constexpr size_t SMEM = (size_t)(PATCH_FLOATS + W_FLOATS + 4 * G) * 4;
};

template <int CIN, int COUT, int KS, int HIN, int HOUT, int TH, int TW, int G, int CC, int COC, int PXT, bool FINAL, bool IN_NHWC>
__global__ void __launch_bounds__(ConvCfg<CIN, COUT, KS, HIN, HOUT, TH, TW, G, CC, COC, PXT, FINAL>::NTHREADS)
conv_gn_kernel(const float* __restrict__ in, const double* __restrict__ in_stats, const float* __restrict__ gam,
               const float* __restrict__ bet, const float* __restrict__ Wk, const float* __restrict__ bias,
               float* __restrict__ out, double* __restrict__ out_stats, int n) {
  using Cfg = ConvCfg<CIN, COUT, KS, HIN, HOUT, TH, TW, G, CC, COC, PXT, FINAL>;
  constexpr int S = 2, PH = Cfg::PH, PW = Cfg::PW, PWH = Cfg::PWH, NT = Cfg::NTHREADS;
  extern __shared__ __align__(16) float smem[];
  float* patch = smem;                         // [G][CC][PH][2][PWH]
  float* wsm = patch + Cfg::PATCH_FLOATS;      // [CC*KS*KS][COC]
  float* gstat = wsm + Cfg::W_FLOATS;          // [G][2] mean, rstd
  float* ssum = gstat + 2 * G;                 // [G][2]
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int ty0 = (tile / Cfg::TILES_X) * TH, tx0 = (tile % Cfg::TILES_X) * TW;
  const int crop0 = blockIdx.y * G;
  const int co0 = blockIdx.z * COC;
  if (tid < G) {
    const int crop = crop0 + tid;
    float mean = 0.f, rstd = 0.f;
    if (crop < n) {
      const double cnt = (double)CIN * HIN * HIN;
      const double mu = in_stats[(size_t)crop * 2] / cnt;
      double var = in_stats[(size_t)crop * 2 + 1] / cnt - mu * mu;
      if (var < 0.0) var = 0.0;
      mean = (float)mu;
      rstd = (float)(1.0 / sqrt(var + 1e-5));
    }
    gstat[tid * 2] = mean;
    gstat[tid * 2 + 1] = rstd;
    ssum[tid * 2] = 0.f;
    ssum[tid * 2 + 1] = 0.f;
  }
  constexpr int PIX_PER_G = TH * TW / PXT;
  const int g = tid / PIX_PER_G;
  const int rem = tid % PIX_PER_G;
  const int py = rem / TW, px = rem % TW;
  float acc[PXT][COC];
#pragma unroll
  for (int p = 0; p < PXT; p++)
#pragma unroll
    for (int c = 0; c < COC; c++) acc[p][c] = 0.f;

  for (int c0 = 0; c0 < CIN; c0 += CC) {
    __syncthreads();
    // stage the input patch with GroupNorm + ReLU applied
    for (int i = tid; i < G * CC * PH * PW; i += NT) {
      const int col = i % PW;
      int r = i / PW;
      const int row = r % PH;
      r /= PH;
      const int c = r % CC;
      const int gg = r / CC;
      const int crop = crop0 + gg;
      const int iy = ty0 * S + row, ix = tx0 * S + col;
      float v = 0.f;
      if (crop < n && iy < HIN && ix < HIN) {
        const float x = IN_NHWC ? __ldg(in + (((size_t)crop * HIN + iy) * HIN + ix) * CIN + c0 + c)
                                : __ldg(in + (((size_t)crop * CIN + c0 + c) * HIN + iy) * HIN + ix);
        const float xn = (x - gstat[gg * 2]) * gstat[gg * 2 + 1];
        v = fmaxf(fmaf(xn, __ldg(gam + c0 + c), __ldg(bet + c0 + c)), 0.f);
      }
      patch[(((gg * CC + c) * PH + row) * 2 + (col & 1)) * PWH + (col >> 1)] = v;
    }
    for (int i = tid; i < CC * KS * KS * COC; i += NT) {
      const int co = i % COC, k = i / COC;
      wsm[i] = __ldg(Wk + ((size_t)c0 * KS * KS + k) * COUT + co0 + co);
    }
    __syncthreads();
    for (int c = 0; c < CC; c++) {
#pragma unroll
      for (int ky = 0; ky < KS; ky++) {
#pragma unroll
        for (int kx = 0; kx < KS; kx++) {
          const float* wrow = wsm + ((c * KS + ky) * KS + kx) * COC;
          float v[PXT];
#pragma unroll
          for (int p = 0; p < PXT; p++) {
            const int row = S * (py + p * (TH / PXT)) + ky;
            v[p] = patch[(((g * CC + c) * PH + row) * 2 + (kx & 1)) * PWH + px + (kx >> 1)];
          }
#pragma unroll
          for (int q = 0; q < COC / 4; q++) {
            const float4 w = *reinterpret_cast<const float4*>(wrow + q * 4);
#pragma unroll
            for (int p = 0; p < PXT; p++) {
              acc[p][q * 4 + 0] = fmaf(v[p], w.x, acc[p][q * 4 + 0]);
              acc[p][q * 4 + 1] = fmaf(v[p], w.y, acc[p][q * 4 + 1]);
              acc[p][q * 4 + 2] = fmaf(v[p], w.z, acc[p][q * 4 + 2]);
              acc[p][q * 4 + 3] = fmaf(v[p], w.w, acc[p][q * 4 + 3]);
            }
          }
        }
      }
    }
  }
  const int crop = crop0 + g;
  float s1 = 0.f, s2 = 0.f;
  if (crop < n) {
#pragma unroll
    for (int p = 0; p < PXT; p++) {
      const int oy = ty0 + py + p * (TH / PXT), ox = tx0 + px;
      if (oy < HOUT && ox < HOUT) {
#pragma unroll
        for (int c = 0; c < COC; c++) {
          const float v = acc[p][c] + __ldg(bias + co0 + c);
          if (FINAL) out[(size_t)crop * COUT + co0 + c] = v;
          else out[(((size_t)crop * COUT + co0 + c) * HOUT + oy) * HOUT + ox] = v;
          s1 += v;
          s2 = fmaf(v, v, s2);
        }
      }
    }
  }
  if (!FINAL) {
    atomicAdd(&ssum[g * 2], s1);
    atomicAdd(&ssum[g * 2 + 1], s2);
    __syncthreads();
    if (tid < G && crop0 + tid < n) {
      atomicAdd(out_stats + (size_t)(crop0 + tid) * 2, (double)ssum[tid * 2]);
      atomicAdd(out_stats + (size_t)(crop0 + tid) * 2 + 1, (double)ssum[tid * 2 + 1]);
    }
  }
}

template <int CIN, int COUT, int KS, int HIN, int HOUT, int TH, int TW, int G, int CC, int COC, int PXT, bool FINAL, bool IN_NHWC = false>
static int launch_conv(const char* name, const float* in, const double* in_stats, const float* gam, const float* bet, const float* Wk,
                       const float* bias, float* out, double* out_stats, int n, cudaStream_t stream) {
  using Cfg = ConvCfg<CIN, COUT, KS, HIN, HOUT, TH, TW, G, CC, COC, PXT, FINAL>;
  static_assert(CIN % CC == 0 && COUT % COC == 0 && COC % 4 == 0 && (TH % PXT) == 0, "bad conv tiling");
  auto kern = conv_gn_kernel<CIN, COUT, KS, HIN, HOUT, TH, TW, G, CC, COC, PXT, FINAL, IN_NHWC>;
  static unsigned attr_done = 0;
  if (strive_first_use_on_device(&attr_done)) {
    STRIVE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
  }
  dim3 grid(Cfg::TILES, (n + G - 1) / G, COUT / COC);
  KPROF(name, stream, kern<<<grid, Cfg::NTHREADS, Cfg::SMEM, stream>>>(in, in_stats, gam, bet, Wk, bias, out, out_stats, n));
  STRIVE_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// workspace + driver
// ------------------------------------------------------------------------------------------------------
#define MAPENC_CHUNK 2048
extern "C" int strive_mapenc_set_impl(int impl) {
  g_mapenc_impl = impl ? 1 : 0;
  return 0;
}
int tc_launch_conv1(const StriveMap* map, const float* pose, const int32_t* map_of, const uint8_t* wpack, const float* bias, float* out,
                    double* out_stats, uint8_t* packed_crop, int n, cudaStream_t stream);
int tc_launch_conv2(const float* in, const double* in_stats, const float* gam, const float* bet, const uint8_t* wpack, const float* bias,
                    float* out, double* out_stats, int n, cudaStream_t stream);
int tc_launch_conv3(const float* in, const double* in_stats, const float* gam, const float* bet, const uint8_t* wpack, const uint8_t* wpack_pair, const float* bias,
                    float* out, double* out_stats, int n, cudaStream_t stream);
int tc_launch_conv4(const float* in, const double* in_stats, const float* gam, const float* bet, const uint8_t* wpack, const float* bias,
                    float* out, double* out_stats, int n, cudaStream_t stream);
int tc_launch_conv5(const float* in, const double* in_stats, const float* gam, const float* bet, const uint8_t* wpack, const float* bias,
                    float* out, double* out_stats, int n, cudaStream_t stream);
int tc_launch_conv6(const float* in, const double* in_stats, const float* gam, const float* bet, const uint8_t* wpack, const float* bias,
                    float* out, double* out_stats, int n, cudaStream_t stream);
int tc_launch_fc(const float* in, const double* in_stats, const float* gam, const float* bet, const uint8_t* wpack, const float* bias,
                 float* out, int n, cudaStream_t stream);
static const size_t kActFloats[6] = {16 * 125 * 125, 32 * 61 * 61, 64 * 29 * 29, 64 * 14 * 14, 128 * 6 * 6, 128 * 2 * 2};

extern "C" int64_t strive_mapenc_workspace_bytes(int32_t n) {
  const size_t c = (size_t)(n < MAPENC_CHUNK ? n : MAPENC_CHUNK);
  size_t fl = 0;
  for (int i = 0; i < 6; i++) fl += ((kActFloats[i] * c + 63) & ~(size_t)63);
  return (int64_t)(fl * 4 + 6 * c * 2 * 8 + c * 65536 + 512);
}

// second stream of the half-chunk pipeline (one per device, created on first use; the drivers are single-threaded per device)
static int g_mapenc_split = 0;     // measured on B200 at 2048 crops: +0.1 .. 0.2 % (the SMs sit at the 1 kW power cap: overlap lowers the clock) -> off by default
extern "C" int strive_mapenc_set_split(int on) {
  g_mapenc_split = on;
  return 0;
}
static int mapenc_side_stream(cudaStream_t* side, cudaEvent_t* ev_fork, cudaEvent_t* ev_join) {
  static cudaStream_t s[16] = {};
  static cudaEvent_t e0[16] = {}, e1[16] = {};
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

extern "C" int strive_mapenc_fwd(const StriveModel* m, const StriveMap* map, const float* pose_un, const int32_t* map_of,
                                 int32_t n, float* out_feat, void* workspace, int64_t workspace_bytes, void* stream_) {
  STRIVE_CHECK(m && map && pose_un && map_of && out_feat && workspace, STRIVE_EINVAL, "strive_mapenc_fwd: null argument");
  STRIVE_CHECK(n > 0, STRIVE_EINVAL, "strive_mapenc_fwd: n=%d", n);
  STRIVE_CHECK(map->C == 4, STRIVE_EUNSUPPORTED, "map encoder expects 4 raster layers (conv_channel_in=4), got %d", map->C);
  STRIVE_CHECK(workspace_bytes >= strive_mapenc_workspace_bytes(n), STRIVE_ESIZE, "mapenc workspace too small");
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t c = (size_t)(n < MAPENC_CHUNK ? n : MAPENC_CHUNK);
  float* act[6];
  {
    float* p = (float*)workspace;
    for (int i = 0; i < 6; i++) { act[i] = p; p += ((kActFloats[i] * c + 63) & ~(size_t)63); }
    // stats follow (8-byte aligned because every block above is a multiple of 64 floats)
    workspace = (void*)p;
  }
  double* stats = (double*)workspace;   // [6][c][2]
  uint8_t* packed_crop = (uint8_t*)(stats + 6 * c * 2);   // [c][256][256] bit-packed crops (tensor-core path)
  packed_crop = (uint8_t*)(((uintptr_t)packed_crop + 255) & ~(uintptr_t)255);
  const float* const* sg = m->seg;
  for (int start = 0; start < n; start += MAPENC_CHUNK) {
    const int cn = (n - start) < MAPENC_CHUNK ? (n - start) : MAPENC_CHUNK;
    STRIVE_CUDA(cudaMemsetAsync(stats, 0, 6 * c * 2 * sizeof(double), stream));
    double* st[6];
    for (int i = 0; i < 6; i++) st[i] = stats + (size_t)i * c * 2;
    const float* pose = pose_un + (size_t)start * 4;
    const int32_t* mo = map_of + start;
    int rc;
    if (g_mapenc_impl == 1 && m->tc_blob != nullptr) {
      // tensor-core path (mapenc_tc.cu): conv1..conv4 on tcgen05, activations channel-blocked fp32.
      // The crops of a chunk are independent, so a chunk CAN run as two half-chunks on two streams (strive_mapenc_set_split(1); fork /
      // join with events; inside a stream capture the side stream becomes a second branch of the graph): the crop gather of one half
      // (L1 / ALU bound, no TMEM, little shared memory) is then co-resident with the TMEM-read / HBM-write bound conv1 of the other
      // and tail waves are filled.  Measured: 56.11 -> 56.05 ms per iteration at BASELINE configs[1] -- the step is power-capped
      // (sw_power_cap, 1.77 .. 1.95 GHz), more overlap buys a lower clock -- so the default is one stream.
      auto half = [&](int h0, int hn, cudaStream_t s) -> int {
        float* a[6];
        double* t[6];
        for (int i = 0; i < 6; i++) { a[i] = act[i] + kActFloats[i] * (size_t)h0; t[i] = st[i] + (size_t)h0 * 2; }
        int r = tc_launch_conv1(map, pose + (size_t)h0 * 4, mo + h0, m->tc_blob + m->tc_off[0], m->h_cbias[0], a[0], t[0], packed_crop + (size_t)h0 * 65536, hn, s);
        if (r) return r;
        r = tc_launch_conv2(a[0], t[0], sg[S_GG0], sg[S_GB0], m->tc_blob + m->tc_off[1], m->h_cbias[1], a[1], t[1], hn, s);
        if (r) return r;
        r = tc_launch_conv3(a[1], t[1], sg[S_GG1], sg[S_GB1], m->tc_blob + m->tc_off[2], m->tc_blob + m->tc_off[7], m->h_cbias[2], a[2], t[2], hn, s);
        if (r) return r;
        r = tc_launch_conv4(a[2], t[2], sg[S_GG2], sg[S_GB2], m->tc_blob + m->tc_off[3], m->h_cbias[3], a[3], t[3], hn, s);
        if (r) return r;
        r = tc_launch_conv5(a[3], t[3], sg[S_GG3], sg[S_GB3], m->tc_blob + m->tc_off[4], sg[S_CB4], a[4], t[4], hn, s);
        if (r) return r;
        r = tc_launch_conv6(a[4], t[4], sg[S_GG4], sg[S_GB4], m->tc_blob + m->tc_off[5], sg[S_CB5], a[5], t[5], hn, s);
        if (r) return r;
        return tc_launch_fc(a[5], t[5], sg[S_GG5], sg[S_GB5], m->tc_blob + m->tc_off[6], sg[S_FCB], out_feat + (size_t)(start + h0) * 64, hn, s);
      };
      const bool split = g_mapenc_split != 0 && g_strive_profile_on == 0 && cn >= 512;
      if (!split) {
        rc = half(0, cn, stream);
        if (rc) return rc;
        continue;
      }
      cudaStream_t side;
      cudaEvent_t ev_fork, ev_join;
      rc = mapenc_side_stream(&side, &ev_fork, &ev_join);
      if (rc) return rc;
      const int h0n = (cn / 2) & ~1;
      STRIVE_CUDA(cudaEventRecord(ev_fork, stream));                 // after the statistics memset
      STRIVE_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
      rc = half(0, h0n, stream);
      if (rc) return rc;
      rc = half(h0n, cn - h0n, side);
      if (rc) return rc;
      STRIVE_CUDA(cudaEventRecord(ev_join, side));
      STRIVE_CUDA(cudaStreamWaitEvent(stream, ev_join, 0));
      continue;
    } else {
      dim3 g1(C1_TILES * C1_TILES, cn);
      KPROF("conv1_gather", stream, conv1_gather_kernel<<<g1, 256, 0, stream>>>(*map, pose, mo, sg[S_CW0], sg[S_CB0], act[0], st[0], cn));
      STRIVE_LAUNCH_CHECK();
      //             CIN COUT KS HIN HOUT TH  TW  G  CC COC PXT FINAL
      rc = launch_conv<16, 32, 5, 125, 61, 16, 16, 1, 8, 32, 2, false>("conv2", act[0], st[0], sg[S_GG0], sg[S_GB0], sg[S_CW1], sg[S_CB1], act[1], st[1], cn, stream);
      if (rc) return rc;
      rc = launch_conv<32, 64, 5, 61, 29, 16, 16, 1, 8, 32, 2, false>("conv3", act[1], st[1], sg[S_GG1], sg[S_GB1], sg[S_CW2], sg[S_CB2], act[2], st[2], cn, stream);
      if (rc) return rc;
      rc = launch_conv<64, 64, 3, 29, 14, 14, 14, 1, 16, 32, 2, false>("conv4", act[2], st[2], sg[S_GG2], sg[S_GB2], sg[S_CW3], sg[S_CB3], act[3], st[3], cn, stream);
      if (rc) return rc;
      rc = launch_conv<64, 128, 3, 14, 6, 6, 6, 4, 16, 32, 2, false>("conv5", act[3], st[3], sg[S_GG3], sg[S_GB3], sg[S_CW4], sg[S_CB4], act[4], st[4], cn, stream);
      if (rc) return rc;
    }
    rc = launch_conv<128, 128, 3, 6, 2, 2, 2, 32, 16, 32, 2, false>("conv6", act[4], st[4], sg[S_GG4], sg[S_GB4], sg[S_CW5], sg[S_CB5], act[5], st[5], cn, stream);
    if (rc) return rc;
    rc = launch_conv<128, 64, 2, 2, 1, 1, 1, 64, 16, 32, 1, true>("fc", act[5], st[5], sg[S_GG5], sg[S_GB5], sg[S_FCW], sg[S_FCB], out_feat + (size_t)start * 64, nullptr, cn, stream);
    if (rc) return rc;
  }
  return 0;
}
