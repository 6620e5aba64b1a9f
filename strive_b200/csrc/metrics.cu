// Success / plausibility checks that bracket the latent loop in the reference drivers:
//   on-layer fraction   src/datasets/nuscenes_utils.py:266-298 (check_on_layer)  -> compute_coll_rate_env,
//                       src/losses/traffic_model.py:366-419
//   line / layer test   src/datasets/nuscenes_utils.py:300-333 (check_line_layer) -> determine_feasibility_nusc,
//                       src/utils/scenario_gen.py:91-99
//   rectangle IoU hits  src/losses/adv_gen_nusc.py:517-623 (check_single_veh_coll / check_pairwise_veh_coll; the reference
//                       builds shapely Polygons from nutils.get_corners, nuscenes_utils.py:416-428, and thresholds
//                       intersection / union at VEH_COLL_THRESH)
// Integer / byte work (pixel indices, layer reads) is exact; the IoU is evaluated in float64 on float32 corners, as the
// reference does (numpy float32 corners handed to shapely's float64 geometry).
#include "common.cuh"

// ------------------------------------------------------------------------------------------------------
// fraction of an L x W grid over each car's footprint that reads 1 in `layer` of its map; one warp per car
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) on_layer_frac_kernel(StriveMap map, int layer, const float* __restrict__ cars,
                                                            const float* __restrict__ lw, const int32_t* __restrict__ map_of,
                                                            const float* __restrict__ lin_l, const float* __restrict__ lin_w, int L, int W,
                                                            int n, float* __restrict__ frac) {
  const int widx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (widx >= n) return;
  const float4 p = *reinterpret_cast<const float4*>(cars + (size_t)widx * 4);
  const bool nan_in = isnan(p.x + p.y + p.z + p.w);          // traffic_model.py:400: NaN frames are not checked (fraction 1)
  const float l = lw[(size_t)widx * 2], w = lw[(size_t)widx * 2 + 1];
  const int m = map_of[widx];
  const double dx0 = map.dx[m * 2], dx1 = map.dx[m * 2 + 1];
  const uint8_t* lay = map.raster + ((size_t)m * map.C + layer) * map.H * map.W;
  int cnt = 0;
  if (!nan_in) {
    for (int s = lane; s < L * W; s += 32) {
      const int ia = s / W, ib = s % W;
      // gen_car_coords ls/ws branch (nuscenes_utils.py:222-224, 229-232), float32 with torch's operation order
      const float lw_ = __fmul_rn(__ldg(lin_l + ia), l) * 0.5f;
      const float ww_ = __fmul_rn(__ldg(lin_w + ib), w) * 0.5f;
      const float wx = __fadd_rn(__fsub_rn(__fmul_rn(lw_, p.z), __fmul_rn(ww_, p.w)), p.x);
      const float wy = __fadd_rn(__fadd_rn(__fmul_rn(lw_, p.w), __fmul_rn(ww_, p.z)), p.y);
      long long xp = (long long)rint((double)wx / dx0);       // :285-286 float32 / float64 -> float64, round half even
      long long yp = (long long)rint((double)wy / dx1);
      if (yp < 0 || yp >= map.H || xp < 0 || xp >= map.W) { xp = 0; yp = 0; }   // :291-292
      cnt += (__ldg(lay + (size_t)yp * map.W + xp) != 0) ? 1 : 0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) frac[widx] = nan_in ? 1.0f : (float)cnt / (float)(L * W);       // :295 sum(float) / (L*W)
}

extern "C" int strive_on_layer_frac(const StriveMap* map, int32_t layer, const float* cars_un, const float* lw_un, const int32_t* map_of,
                                    const float* lin_l, const float* lin_w, int32_t L, int32_t W, int32_t n, float* frac_out,
                                    void* stream_) {
  STRIVE_CHECK(map != nullptr && map->raster != nullptr && map->dx != nullptr, STRIVE_EINVAL, "on_layer_frac: null map");
  STRIVE_CHECK(layer >= 0 && layer < map->C, STRIVE_EINVAL, "on_layer_frac: layer %d outside 0..%d", layer, map->C - 1);
  STRIVE_CHECK(L > 0 && W > 0 && n >= 0, STRIVE_EINVAL, "on_layer_frac: bad sizes L=%d W=%d n=%d", L, W, n);
  if (n == 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  const long long threads = (long long)n * 32;
  KPROF("on_layer_frac", stream, on_layer_frac_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(*map, layer, cars_un, lw_un, map_of, lin_l,
                                                                                                      lin_w, L, W, n, frac_out));
  STRIVE_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// does the straight line start -> end touch a 0 pixel of `layer`?  one thread per line, NL samples (batch-global, :316-320)
// ------------------------------------------------------------------------------------------------------
__global__ void line_layer_kernel(StriveMap map, int layer, const float* __restrict__ start, const float* __restrict__ end,
                                  const int32_t* __restrict__ map_of, const float* __restrict__ lin01, int NL, int n,
                                  uint8_t* __restrict__ hit, int* __restrict__ oob) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float sx = start[i * 2], sy = start[i * 2 + 1], ex = end[i * 2], ey = end[i * 2 + 1];
  const int m = map_of[i];
  const double dx0 = map.dx[m * 2], dx1 = map.dx[m * 2 + 1];
  const uint8_t* lay = map.raster + ((size_t)m * map.C + layer) * map.H * map.W;
  int bad = 0;
  for (int k = 0; k < NL; k++) {
    const float wgt = __ldg(lin01 + k);
    const float om = __fsub_rn(1.0f, wgt);
    const float x = __fadd_rn(__fmul_rn(sx, om), __fmul_rn(ex, wgt));           // :321
    const float y = __fadd_rn(__fmul_rn(sy, om), __fmul_rn(ey, wgt));
    long long xp = (long long)rint((double)x / dx0);
    long long yp = (long long)rint((double)y / dx1);
    if (yp < -map.H || yp >= map.H || xp < -map.W || xp >= map.W) { atomicAdd(oob, 1); continue; }   // the reference would raise IndexError
    if (yp < 0) yp += map.H;                                                      // torch advanced indexing wraps negatives
    if (xp < 0) xp += map.W;
    bad += (__ldg(lay + (size_t)yp * map.W + xp) == 0) ? 1 : 0;
  }
  hit[i] = bad > 0 ? 1 : 0;
}

extern "C" int strive_line_layer(const StriveMap* map, int32_t layer, const float* start_un, const float* end_un, const int32_t* map_of,
                                 const float* lin01, int32_t num_samples, int32_t n, uint8_t* hit_out, int32_t* oob_count, void* stream_) {
  STRIVE_CHECK(map != nullptr && map->raster != nullptr && map->dx != nullptr, STRIVE_EINVAL, "line_layer: null map");
  STRIVE_CHECK(layer >= 0 && layer < map->C, STRIVE_EINVAL, "line_layer: layer %d outside 0..%d", layer, map->C - 1);
  STRIVE_CHECK(num_samples >= 0 && n >= 0 && oob_count != nullptr, STRIVE_EINVAL, "line_layer: bad arguments");
  if (n == 0) return 0;
  cudaStream_t stream = (cudaStream_t)stream_;
  STRIVE_CUDA(cudaMemsetAsync(oob_count, 0, sizeof(int32_t), stream));
  KPROF("line_layer", stream, line_layer_kernel<<<(n + 127) / 128, 128, 0, stream>>>(*map, layer, start_un, end_un, map_of, lin01, num_samples, n,
                                                                               hit_out, oob_count));
  STRIVE_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// rotated-rectangle IoU hits.  hit[i][j][t] = IoU(rect(a_i(t)), rect(b_j(t))) > thresh; NaN in either state -> 0.
// ------------------------------------------------------------------------------------------------------
struct P2 { double x, y; };

// nutils.get_corners (nuscenes_utils.py:416-428) in float32 as numpy evaluates it for float32 inputs:
// corner (px,py) of [-l/2,-w/2],[l/2,-w/2],[l/2,w/2],[-l/2,w/2] times [[c, s], [-s, c]] plus (x, y)
__device__ __forceinline__ void rect_corners(const float st[4], float l, float w, P2 out[4]) {
  const float h = atan2f(st[3], st[2]);
  const float c = cosf(h), s = sinf(h);
  const float hl = l / 2.0f, hw = w / 2.0f;
  const float bx[4] = {-hl, hl, hl, -hl};
  const float by[4] = {-hw, -hw, hw, hw};
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const float x = __fadd_rn(__fadd_rn(__fmul_rn(bx[k], c), __fmul_rn(by[k], -s)), st[0]);   // np.dot row: bx*c + by*(-s)
    const float y = __fadd_rn(__fadd_rn(__fmul_rn(bx[k], s), __fmul_rn(by[k], c)), st[1]);
    out[k].x = (double)x;
    out[k].y = (double)y;
  }
}

__device__ __forceinline__ double poly_area(const P2* p, int n) {
  double a = 0.0;
  for (int k = 0; k < n; k++) {
    const P2 u = p[k], v = p[(k + 1 == n) ? 0 : k + 1];
    a += u.x * v.y - v.x * u.y;
  }
  return 0.5 * a;
}

// Sutherland-Hodgman: clip convex polygon `subj` (n <= 8) against the CCW convex quadrilateral `clip`; returns the area
__device__ double rect_intersection_area(const P2 subj[4], const P2 clip[4]) {
  P2 cur[10], nxt[10];
  int n = 4;
  for (int k = 0; k < 4; k++) cur[k] = subj[k];
  for (int e = 0; e < 4 && n > 0; e++) {
    const P2 a = clip[e], b = clip[(e + 1) & 3];
    const double ex = b.x - a.x, ey = b.y - a.y;
    int m = 0;
    for (int k = 0; k < n; k++) {
      const P2 p = cur[k], q = cur[(k + 1 == n) ? 0 : k + 1];
      const double dp = ex * (p.y - a.y) - ey * (p.x - a.x);     // > 0: left of the edge = inside
      const double dq = ex * (q.y - a.y) - ey * (q.x - a.x);
      if (dp >= 0.0) nxt[m++] = p;
      if ((dp > 0.0 && dq < 0.0) || (dp < 0.0 && dq > 0.0)) {
        const double t = dp / (dp - dq);
        P2 r;
        r.x = p.x + t * (q.x - p.x);
        r.y = p.y + t * (q.y - p.y);
        nxt[m++] = r;
      }
    }
    n = m;
    for (int k = 0; k < n; k++) cur[k] = nxt[k];
  }
  if (n < 3) return 0.0;
  return fabs(poly_area(cur, n));
}

__global__ void __launch_bounds__(128) veh_iou_kernel(const float* __restrict__ traj_a, const float* __restrict__ lw_a, int na,
                                                      const float* __restrict__ traj_b, const float* __restrict__ lw_b, int nb, int T,
                                                      double thresh, uint8_t* __restrict__ hit, float* __restrict__ iou_out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)na * nb * T) return;
  const int t = (int)(idx % T);
  const int j = (int)((idx / T) % nb);
  const int i = (int)(idx / ((long long)T * nb));
  const float4 sa = *reinterpret_cast<const float4*>(traj_a + ((size_t)i * T + t) * 4);
  const float4 sb = *reinterpret_cast<const float4*>(traj_b + ((size_t)j * T + t) * 4);
  uint8_t h = 0;
  float iou = 0.f;
  if (!isnan(sa.x + sa.y + sa.z + sa.w) && !isnan(sb.x + sb.y + sb.z + sb.w)) {
    const float a4[4] = {sa.x, sa.y, sa.z, sa.w}, b4[4] = {sb.x, sb.y, sb.z, sb.w};
    P2 ra[4], rb[4];
    rect_corners(a4, lw_a[i * 2], lw_a[i * 2 + 1], ra);
    rect_corners(b4, lw_b[j * 2], lw_b[j * 2 + 1], rb);
    // cheap reject: centres further apart than the two half-diagonals
    const double cdx = (double)sa.x - (double)sb.x, cdy = (double)sa.y - (double)sb.y;
    const double ra2 = 0.25 * ((double)lw_a[i * 2] * lw_a[i * 2] + (double)lw_a[i * 2 + 1] * lw_a[i * 2 + 1]);
    const double rb2 = 0.25 * ((double)lw_b[j * 2] * lw_b[j * 2] + (double)lw_b[j * 2 + 1] * lw_b[j * 2 + 1]);
    const double reach = sqrt(ra2) + sqrt(rb2) + 1e-3;
    if (cdx * cdx + cdy * cdy <= reach * reach) {
      const double inter = rect_intersection_area(ra, rb);
      const double aa = fabs(poly_area(ra, 4)), ab = fabs(poly_area(rb, 4));
      const double uni = aa + ab - inter;
      const double v = uni > 0.0 ? inter / uni : 0.0;
      iou = (float)v;
      h = v > thresh ? 1 : 0;
    }
  }
  hit[idx] = h;
  if (iou_out != nullptr) iou_out[idx] = iou;
}

extern "C" int strive_veh_iou_hits(const float* traj_a_un, const float* lw_a_un, int32_t na, const float* traj_b_un, const float* lw_b_un,
                                   int32_t nb, int32_t T, double iou_thresh, uint8_t* hit_out, float* iou_out, void* stream_) {
  STRIVE_CHECK(na >= 0 && nb >= 0 && T >= 0, STRIVE_EINVAL, "veh_iou_hits: negative size");
  const long long tot = (long long)na * nb * T;
  if (tot == 0) return 0;
  STRIVE_CHECK(traj_a_un && lw_a_un && traj_b_un && lw_b_un && hit_out, STRIVE_EINVAL, "veh_iou_hits: null pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  KPROF("veh_iou", stream, veh_iou_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, stream>>>(traj_a_un, lw_a_un, na, traj_b_un, lw_b_un, nb, T,
                                                                                       iou_thresh, hit_out, iou_out));
  STRIVE_LAUNCH_CHECK();
  return 0;
}
