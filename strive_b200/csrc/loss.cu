// Fused forward+backward of the latent-optimisation losses (reference src/losses/adv_gen_nusc.py).
//
//   interp_traj          :625-644   linear x3 upsample (align_corners=False) + heading renormalisation
//   VehCollLoss          :405-512   5-circle all-pairs penalty inside a collision block (block-diagonal work only)
//   EnvCollLoss          :366-403   + datasets/nuscenes_utils.py:334-390 get_coll_point (drivable raster samples)
//   MotionPriorLoss      :343-364   + losses/common.py:26-41 log_normal
//   AvoidCollLoss        :303-341,  AdvGenLoss :93-262 (+ check_behind :646-673), TgtMatchingLoss :27-51
//
// Every .mean() of the reference is over data-dependent counts of one reference batch ("group"); kernels
// accumulate raw sums/counts per group in fp64 and un-normalised gradients, and the finalize kernel applies
// weight/count, the heading-renormalisation and interpolation adjoints and the state normaliser, writing
// dL/d(traj_normalised) directly.  No host synchronisation anywhere.
#include "common.cuh"

enum Acc { A_SUMA, A_CNTA, A_SUMB, A_CNTB, A_SUME, A_CNTE, A_PRIOR, A_INIT, A_CRASH, A_MATCH, A_NOTBEHIND, A_NATK, A_N = 16 };

struct LossWs {
  float* ti;      // [NA][T3][4] interpolated, unnormalised, heading renormalised
  float* un;      // [NA][T3]    norm of the interpolated heading before renormalisation
  float* gA;      // [NA][T3][4] d(sum of veh penalties, term A)/d ti
  float* gB;      // [NA][T3][4] term B (ego pairs, AdvGenLoss coll_veh_plan)
  float* gE;      // [NA][T3][2] env term
  float* dist;    // [NA][FT]    attacker-target distance (ADV)
  float* gC;      // [NA][FT][2] d(sum_b crash_b)/d pos (unnormalised) (ADV)
  float* rew;     // [NA]        prior_reweight (1 for ego / non ADV)
  int32_t* behind;  // [NA]
  double* acc;    // [G][A_N]
};

static int64_t loss_carve(LossWs* w, char* base, int NA, int FT, int G) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += (bytes + 255) & ~(size_t)255;
    return p;
  };
  const size_t n = NA, T3 = 3 * (size_t)FT;
  w->ti = (float*)take(n * T3 * 16);
  w->un = (float*)take(n * T3 * 4);
  w->gA = (float*)take(n * T3 * 16);
  w->gB = (float*)take(n * T3 * 16);
  w->gE = (float*)take(n * T3 * 8);
  w->dist = (float*)take(n * FT * 4);
  w->gC = (float*)take(n * FT * 8);
  w->rew = (float*)take(n * 4);
  w->behind = (int32_t*)take(n * 4);
  w->acc = (double*)take((size_t)G * A_N * 8);
  return (int64_t)off;
}

extern "C" int64_t strive_loss_workspace_bytes(int32_t num_agents, int32_t ft, int32_t num_groups) {
  LossWs w;
  return loss_carve(&w, nullptr, num_agents, ft, num_groups);
}

struct LossArgs {
  StriveLossCfg cfg;
  LossWs ws;
  StriveMap map;
  int NA, S, FT, T3;
  const int32_t* ptr;
  const int32_t* scene_of;
  const int32_t* map_idx;
  const float* traj;
  const float* z;
  const float* prior_mu;
  const float* prior_var;
  const float* init_z;
  const uint8_t* z_mask;
  const float* match_tgt;
  const uint8_t* match_mask;
  const float* adv_tgt;
  float* d_traj;
  float* d_traj_match;
  float* d_z;
  float* terms;
};

__device__ __forceinline__ void unnorm4(const LossArgs& a, const float* p, float o[4]) {
  // MeanStdNormalizer.unnormalize: (x*std) + mean, separate roundings (datasets/utils.py:104)
  if (a.cfg.traj_unnormalized) {
#pragma unroll
    for (int k = 0; k < 4; k++) o[k] = p[k];
  } else {
#pragma unroll
    for (int k = 0; k < 4; k++) o[k] = __fadd_rn(__fmul_rn(p[k], kStateStd[k]), kStateMean[k]);
  }
}
__device__ __forceinline__ float out_scale(const LossArgs& a, int k) { return a.cfg.traj_unnormalized ? 1.0f : kStateStd[k]; }

// F.interpolate(mode='linear', scale_factor=3, align_corners=False) source index/weights for output o
__device__ __forceinline__ void interp_src(int o, int FT, int& i0, int& i1, float& l0, float& l1) {
  const float scale = (float)(1.0 / 3.0);
  float src = scale * ((float)o + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  i1 = i0 + ((i0 < FT - 1) ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.0f - l1;
}

__device__ __forceinline__ bool is_ego(const LossArgs& a, int ag) { return ag == a.ptr[a.scene_of[ag]]; }
// trajectory attacked by the adversarial term at (scene s, step t): the external planner future, or -- closed-loop mode -- the
// model's own prediction of the target, row ptr[s] of the rollout (adv_gen_optim.py:143)
__device__ __forceinline__ const float* adv_tgt_at(const LossArgs& a, int s, int t) {
  return a.cfg.adv_own_pred ? a.traj + ((size_t)a.ptr[s] * a.FT + t) * 4 : a.adv_tgt + ((size_t)s * a.FT + t) * 4;
}

// ------------------------------------------------------------------------------------------------------
__global__ void loss_zero_kernel(LossArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.cfg.num_groups * A_N) a.ws.acc[i] = 0.0;
  if (i < a.NA) { a.ws.rew[i] = 1.0f; a.ws.behind[i] = 0; }
}

__global__ void interp_kernel(LossArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.NA * a.T3) return;
  const int ag = idx / a.T3, o = idx % a.T3;
  int i0, i1;
  float l0, l1;
  interp_src(o, a.FT, i0, i1, l0, l1);
  float p0[4], p1[4];
  unnorm4(a, a.traj + ((size_t)ag * a.FT + i0) * 4, p0);
  unnorm4(a, a.traj + ((size_t)ag * a.FT + i1) * 4, p1);
  float v[4];
#pragma unroll
  for (int k = 0; k < 4; k++) v[k] = l0 * p0[k] + l1 * p1[k];
  const float nrm = sqrtf(v[2] * v[2] + v[3] * v[3]);
  a.ws.un[idx] = nrm;
  *reinterpret_cast<float4*>(a.ws.ti + (size_t)idx * 4) = make_float4(v[0], v[1], v[2] / nrm, v[3] / nrm);
}

// ------------------------------------------------------------------------------------------------------
// ADV stage 1: attacker-target distances + "behind" flags   (adv_gen_nusc.py:111-123, 646-673)
// ------------------------------------------------------------------------------------------------------
__global__ void adv_dist_kernel(LossArgs a) {
  const int ag = blockIdx.x * blockDim.x + threadIdx.x;
  if (ag >= a.NA) return;
  if (is_ego(a, ag)) return;
  const int s = a.scene_of[ag];
  const int g = a.cfg.group_of[ag];
  int nbehind = 0;
  const int t0 = a.cfg.crash_min_t;
  for (int t = t0; t < a.FT; t++) {
    float p[4], q[4];
    unnorm4(a, a.traj + ((size_t)ag * a.FT + t) * 4, p);
    unnorm4(a, adv_tgt_at(a, s, t), q);
    const float dx = p[0] - q[0], dy = p[1] - q[1];
    const float d = sqrtf(dx * dx + dy * dy);
    a.ws.dist[(size_t)ag * a.FT + t] = d;
    if (a.cfg.use_infront) {
      const float cs = (dx / d) * q[2] + (dy / d) * q[3];
      if (cs < a.cfg.crash_min_infront) nbehind++;
    }
  }
  const int behind_all = (a.cfg.use_infront && nbehind == a.FT - t0) ? 1 : 0;
  a.ws.behind[ag] = behind_all;
  atomicAdd(a.ws.acc + (size_t)g * A_N + A_NATK, 1.0);
  if (!behind_all) atomicAdd(a.ws.acc + (size_t)g * A_N + A_NOTBEHIND, 1.0);
}

// ADV stage 2: one CTA per scene: softmin over (attackers x time), crash loss, prior_reweight, position grads
__global__ void __launch_bounds__(256) adv_crash_kernel(LossArgs a, int32_t* adv_min_out) {
  __shared__ float red[256];
  __shared__ int redi[256];
  __shared__ float s_min, s_den, s_crash;
  const int s = blockIdx.x;
  const int p0 = a.ptr[s], n = a.ptr[s + 1] - p0;
  const int g = a.cfg.group_of[p0];
  const int t0 = a.cfg.crash_min_t, NT = a.FT - t0;
  const int ne = (n - 1) * NT;
  const int tid = threadIdx.x;
  const bool all_behind = a.cfg.use_infront && (a.ws.acc[(size_t)g * A_N + A_NOTBEHIND] == 0.0);   // :120-122
  auto masked = [&](int ag) {
    bool m = false;
    if (a.cfg.use_infront && !all_behind && a.ws.behind[ag]) m = true;
    if (a.cfg.attack_mask != nullptr && a.cfg.attack_mask[ag] == 0) m = true;
    return m;
  };
  // min over unmasked entries
  float mn = INFINITY;
  for (int e = tid; e < ne; e += 256) {
    const int ag = p0 + 1 + e / NT, t = t0 + e % NT;
    if (!masked(ag)) mn = fminf(mn, a.ws.dist[(size_t)ag * a.FT + t]);
  }
  red[tid] = mn;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (tid < o) red[tid] = fminf(red[tid], red[tid + o]); __syncthreads(); }
  if (tid == 0) s_min = red[0];
  __syncthreads();
  const float dmin = s_min;
  const bool none = !(dmin < INFINITY);   // everything masked -> softmin NaN -> zeros (:135)
  float den = 0.f;
  for (int e = tid; e < ne; e += 256) {
    const int ag = p0 + 1 + e / NT, t = t0 + e % NT;
    if (!masked(ag)) den += expf(-(a.ws.dist[(size_t)ag * a.FT + t] - dmin));
  }
  red[tid] = den;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
  if (tid == 0) s_den = red[0];
  __syncthreads();
  den = s_den;
  float cr = 0.f, best = -1.f;
  int besti = 0;
  for (int e = tid; e < ne; e += 256) {
    const int ag = p0 + 1 + e / NT, t = t0 + e % NT;
    float w = 0.f;
    const float d = a.ws.dist[(size_t)ag * a.FT + t];
    if (!none && !masked(ag)) w = expf(-(d - dmin)) / den;
    cr += w * d * d;
    if (w > best) { best = w; besti = e; }
  }
  red[tid] = cr;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
  if (tid == 0) s_crash = red[0];
  __syncthreads();
  const float crash = s_crash;
  // arg-max of the softmin (first max, as torch.max) for min_agt / min_t (:137-138)
  red[tid] = best;
  redi[tid] = besti;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
      if (red[tid + o] > red[tid] || (red[tid + o] == red[tid] && redi[tid + o] < redi[tid])) { red[tid] = red[tid + o]; redi[tid] = redi[tid + o]; }
    }
    __syncthreads();
  }
  if (tid == 0) {
    atomicAdd(a.ws.acc + (size_t)g * A_N + A_CRASH, (double)crash);
    if (adv_min_out != nullptr) {
      const int e = (ne > 0) ? redi[0] : 0;
      adv_min_out[s * 2] = e / NT + 1;
      adv_min_out[s * 2 + 1] = e % NT + t0;
    }
  }
  // gradients + prior_reweight; gC holds d(sum_b crash_b)/d pos; the 1/B mean and weight are applied in finalize
  for (int k = tid; k < (n - 1); k += 256) {
    const int ag = p0 + 1 + k;
    float wsum = 0.f;
    for (int t = 0; t < a.FT; t++) {
      float gx = 0.f, gy = 0.f;
      if (t >= t0) {
        const float d = a.ws.dist[(size_t)ag * a.FT + t];
        float w = 0.f;
        if (!none && !masked(ag)) w = expf(-(d - dmin)) / den;
        wsum += w;
        if (w > 0.f && d > 0.f) {
          const float dd = w * (2.0f * d + crash - d * d);    // d crash / d dist
          float p[4], q[4];
          unnorm4(a, a.traj + ((size_t)ag * a.FT + t) * 4, p);
          unnorm4(a, adv_tgt_at(a, s, t), q);
          gx = dd * (p[0] - q[0]) / d;
          gy = dd * (p[1] - q[1]) / d;
        }
      }
      a.ws.gC[((size_t)ag * a.FT + t) * 2] = gx;
      a.ws.gC[((size_t)ag * a.FT + t) * 2 + 1] = gy;
    }
    a.ws.rew[ag] = 1.0f - wsum;    // :151-152
  }
  if (!a.cfg.adv_own_pred) {
    if (tid == 0) {
      for (int t = 0; t < a.FT; t++) { a.ws.gC[((size_t)p0 * a.FT + t) * 2] = 0.f; a.ws.gC[((size_t)p0 * a.FT + t) * 2 + 1] = 0.f; }
    }
  } else {
    // the target's own predicted position enters every distance with the opposite sign (the "behind" test is detached, :655-671)
    __syncthreads();
    for (int t = tid; t < a.FT; t += 256) {
      float gx = 0.f, gy = 0.f;
      for (int k = 0; k < n - 1; k++) {
        gx -= a.ws.gC[((size_t)(p0 + 1 + k) * a.FT + t) * 2];
        gy -= a.ws.gC[((size_t)(p0 + 1 + k) * a.FT + t) * 2 + 1];
      }
      a.ws.gC[((size_t)p0 * a.FT + t) * 2] = gx;
      a.ws.gC[((size_t)p0 * a.FT + t) * 2 + 1] = gy;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// vehicle-vehicle collisions: thread per (agent i, interpolated step), loop over the other agents of the block
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) veh_coll_kernel(LossArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.NA * a.T3) return;
  const int i = idx / a.T3, o = idx % a.T3;
  const bool adv = (a.cfg.kind & STRIVE_LOSS_ADV) != 0;
  const int blk = a.cfg.cblock_of[i];
  const int b0 = a.cfg.cblock_ptr[blk], b1 = a.cfg.cblock_ptr[blk + 1];
  const int g = a.cfg.group_of[i];
  const float4 pi = *reinterpret_cast<const float4*>(a.ws.ti + (size_t)idx * 4);
  float cxi[5], cix[5], ciy[5];
#pragma unroll
  for (int k = 0; k < 5; k++) {
    cxi[k] = a.cfg.circ_cx[(size_t)i * 5 + k];
    cix[k] = pi.z * cxi[k] + pi.x;     // inverse transform2frame of (cx, 0): (c*cx + x, s*cx + y)  (:481)
    ciy[k] = pi.w * cxi[k] + pi.y;
  }
  const float ri = a.cfg.lw_un[(size_t)i * 2 + 1] * 0.5f;
  const bool ego_i = is_ego(a, i);
  const int sv = a.cfg.single_veh_idx;
  const bool single_i = (sv >= 0) && (i == a.ptr[a.scene_of[i]] + sv);
  const float rew_i = a.ws.rew[i];
  float gAx = 0.f, gAy = 0.f, gAc = 0.f, gAs = 0.f;
  float gBx = 0.f, gBy = 0.f, gBc = 0.f, gBs = 0.f;
  float sumA = 0.f, sumB = 0.f;
  int cntA = 0, cntB = 0;
  for (int j = b0; j < b1; j++) {
    if (j == i) continue;
    bool ego_j = false;
    if (sv >= 0) {
      const bool single_j = (j == a.ptr[a.scene_of[j]] + sv);
      if (!single_i && !single_j) continue;                    // :453-461
    }
    if (adv) ego_j = is_ego(a, j);
    const float4 pj = *reinterpret_cast<const float4*>(a.ws.ti + ((size_t)j * a.T3 + o) * 4);
    const float rj = a.cfg.lw_un[(size_t)j * 2 + 1] * 0.5f;
    const float pdist = ri + rj + a.cfg.veh_coll_buffer;       // :440
    // quick reject on centre distance (cannot change the result: every circle centre is within l/2 of the pose)
    float best = INFINITY;
    int bk = 0;
    float bjx = 0.f, bjy = 0.f;
#pragma unroll
    for (int l = 0; l < 5; l++) {
      const float cj = a.cfg.circ_cx[(size_t)j * 5 + l];
      const float jx = pj.z * cj + pj.x, jy = pj.w * cj + pj.y;
#pragma unroll
      for (int k = 0; k < 5; k++) {
        const float ddx = cix[k] - jx, ddy = ciy[k] - jy;
        const float d2 = ddx * ddx + ddy * ddy;
        if (d2 < best) { best = d2; bk = k; bjx = jx; bjy = jy; }
      }
    }
    const float d = sqrtf(best);
    if (!(d <= pdist)) continue;                                // :495
    const float pen = 1.0f - d / pdist;                         // :506
    float gx = 0.f, gy = 0.f, gc = 0.f, gs = 0.f;
    if (d > 0.f) {
      float cbx = cix[0], cby = ciy[0], cbo = cxi[0];
#pragma unroll
      for (int k = 1; k < 5; k++)
        if (bk == k) { cbx = cix[k]; cby = ciy[k]; cbo = cxi[k]; }
      const float inv = -1.0f / (d * pdist);
      gx = inv * (cbx - bjx);
      gy = inv * (cby - bjy);
      gc = gx * cbo;
      gs = gy * cbo;
    }
    // the symmetric ordered pair (j,i) contributes the same amount to agent i -> factor 2
    if (adv && (ego_i || ego_j)) {
      const float w = ego_i ? a.ws.rew[j] : rew_i;              // :192-204 ego_pen_mat
      sumB += pen * w;
      cntB++;
      gBx += 2.f * w * gx; gBy += 2.f * w * gy; gBc += 2.f * w * gc; gBs += 2.f * w * gs;
    } else {
      sumA += pen;
      cntA++;
      gAx += 2.f * gx; gAy += 2.f * gy; gAc += 2.f * gc; gAs += 2.f * gs;
    }
  }
  *reinterpret_cast<float4*>(a.ws.gA + (size_t)idx * 4) = make_float4(gAx, gAy, gAc, gAs);
  if (adv) *reinterpret_cast<float4*>(a.ws.gB + (size_t)idx * 4) = make_float4(gBx, gBy, gBc, gBs);
  if (cntA > 0) {
    atomicAdd(a.ws.acc + (size_t)g * A_N + A_SUMA, (double)sumA);
    atomicAdd(a.ws.acc + (size_t)g * A_N + A_CNTA, (double)cntA);
  }
  if (cntB > 0) {
    atomicAdd(a.ws.acc + (size_t)g * A_N + A_SUMB, (double)sumB);
    atomicAdd(a.ws.acc + (size_t)g * A_N + A_CNTB, (double)cntB);
  }
}

// ------------------------------------------------------------------------------------------------------
// environment collisions: one warp per (agent, interpolated step); lanes sweep the L x W footprint grid
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) env_coll_kernel(LossArgs a) {
  const int widx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (widx >= a.NA * a.T3) return;
  const int i = widx / a.T3;
  float gx = 0.f, gy = 0.f;
  bool active = true;
  const bool adv = (a.cfg.kind & STRIVE_LOSS_ADV) != 0;
  if (adv && is_ego(a, i)) active = false;                                            // :86, :213
  const int sv = a.cfg.single_veh_idx;
  if (sv >= 0 && i != a.ptr[a.scene_of[i]] + sv) active = false;                      // :296-300
  if (active) {
    const int g = a.cfg.group_of[i];
    const int L = a.cfg.env_L[g], W = a.cfg.env_W[g];
    const float* linl = a.cfg.env_lin_l + (size_t)g * 128;
    const float* linw = a.cfg.env_lin_w + (size_t)g * 128;
    const float4 p = *reinterpret_cast<const float4*>(a.ws.ti + (size_t)widx * 4);
    const float l = a.cfg.lw_un[(size_t)i * 2], w = a.cfg.lw_un[(size_t)i * 2 + 1];
    const int m = a.cfg.agent_map[i];
    const double dx0 = a.map.dx[m * 2], dx1 = a.map.dx[m * 2 + 1];
    const uint8_t* driv = a.map.raster + (size_t)m * a.map.C * a.map.H * a.map.W;    // layer 0
    int num = 0;
    float sx = 0.f, sy = 0.f;
    for (int sidx = lane; sidx < L * W; sidx += 32) {
      const int ia = sidx / W, ib = sidx % W;
      // gen_car_coords ls/ws branch (nuscenes_utils.py:223-224): linspace(-1,1,.) * ls / 2
      const float lw_ = __fmul_rn(__ldg(linl + ia), l) * 0.5f;
      const float ww_ = __fmul_rn(__ldg(linw + ib), w) * 0.5f;
      const float wx = __fadd_rn(__fsub_rn(__fmul_rn(lw_, p.z), __fmul_rn(ww_, p.w)), p.x);
      const float wy = __fadd_rn(__fadd_rn(__fmul_rn(lw_, p.w), __fmul_rn(ww_, p.z)), p.y);
      long long xp = (long long)rint((double)wx / dx0);
      long long yp = (long long)rint((double)wy / dx1);
      if (yp < 0 || yp >= a.map.H || xp < 0 || xp >= a.map.W) { xp = 0; yp = 0; }
      if (__ldg(driv + (size_t)yp * a.map.W + xp) == 0) { num++; sx += wx; sy += wy; }
    }
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) {
      num += __shfl_xor_sync(0xffffffffu, num, o2);
      sx += __shfl_xor_sync(0xffffffffu, sx, o2);
      sy += __shfl_xor_sync(0xffffffffu, sy, o2);
    }
    if (num > 0 && num < L * W) {                                                     // nan otherwise (:380-381)
      const float cx = sx / (float)num, cy = sy / (float)num;
      const float ddx = p.x - cx, ddy = p.y - cy;
      const float d = sqrtf(ddx * ddx + ddy * ddy);
      const float pend = sqrtf(l * l / 4.0f + w * w / 4.0f);                          // adv_gen_nusc.py:371
      const float pen = 1.0f - d / pend;
      if (d > 0.f) { gx = -ddx / (d * pend); gy = -ddy / (d * pend); }
      if (lane == 0) {
        atomicAdd(a.ws.acc + (size_t)g * A_N + A_SUME, (double)pen);
        atomicAdd(a.ws.acc + (size_t)g * A_N + A_CNTE, 1.0);
      }
    }
  }
  if (lane == 0) { a.ws.gE[(size_t)widx * 2] = gx; a.ws.gE[(size_t)widx * 2 + 1] = gy; }
}

// ------------------------------------------------------------------------------------------------------
// latent terms: one warp per agent row (32 lanes = 32 latent dims)
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) latent_terms_kernel(LossArgs a) {
  const int ag = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (ag >= a.NA) return;
  float dz = 0.f;
  const bool on = (a.z_mask == nullptr) || (a.z_mask[ag] != 0);
  if (on && (a.cfg.kind & (STRIVE_LOSS_AVOID | STRIVE_LOSS_ADV))) {
    const int g = a.cfg.group_of[ag];
    const bool adv = (a.cfg.kind & STRIVE_LOSS_ADV) != 0;
    const float zr = a.z[(size_t)ag * ZDIM + lane];
    const float mu = a.prior_mu[(size_t)ag * ZDIM + lane], var = a.prior_var[(size_t)ag * ZDIM + lane];
    const float rows = (float)a.cfg.group_zrows[g];
    float wp = a.cfg.w_motion_prior, wi = a.cfg.w_init_z;
    if (adv) {
      const float rw = a.ws.rew[ag];
      wp = rw * a.cfg.w_motion_prior + (1.0f - rw) * a.cfg.w_motion_prior_atk;   // :160-161
      wi = rw * a.cfg.w_init_z + (1.0f - rw) * a.cfg.w_init_z_atk;               // :219-220
    }
    const bool use_prior = a.cfg.w_motion_prior > 0.f;
    const bool use_init = a.cfg.w_init_z > 0.f;
    if (use_prior) {
      // -log_normal (losses/common.py:38-40)
      const float nll = logf(sqrtf(var)) + 0.9189385332046727f + (zr - mu) * (zr - mu) / (2.0f * var);
      const float tot = warp_sum(nll);
      if (lane == 0) atomicAdd(a.ws.acc + (size_t)g * A_N + A_PRIOR, (double)(adv ? tot * wp : tot));
      dz += (adv ? wp : a.cfg.w_motion_prior) / rows * (zr - mu) / var;
    }
    if (use_init) {
      const float df = a.init_z[(size_t)ag * ZDIM + lane] - zr;
      const float tot = warp_sum(df * df);
      if (lane == 0) atomicAdd(a.ws.acc + (size_t)g * A_N + A_INIT, (double)(adv ? tot * wi : tot));
      // AVOID: w*mean over rows (:335-336); ADV: weighted SUM, .mean() of a scalar (:217-229)
      dz += -2.0f * df * (adv ? wi : a.cfg.w_init_z / rows);
    }
  }
  a.d_z[(size_t)ag * ZDIM + lane] = dz;
}

// ------------------------------------------------------------------------------------------------------
// match term sums (TgtMatchingLoss :39-41): thread per (agent, t)
// ------------------------------------------------------------------------------------------------------
__global__ void match_sum_kernel(LossArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.NA * a.FT) return;
  if (a.match_mask[idx] == 0) return;
  const int ag = idx / a.FT;
  float p[4], q[4];
  unnorm4(a, a.traj + (size_t)idx * 4, p);
  unnorm4(a, a.match_tgt + (size_t)idx * 4, q);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 4; k++) s += (p[k] - q[k]) * (p[k] - q[k]);
  atomicAdd(a.ws.acc + (size_t)a.cfg.group_of[ag] * A_N + A_MATCH, (double)s);
}

// ------------------------------------------------------------------------------------------------------
// finalize: d_traj (normalised) = state_std * interp^T( renorm^T( wA/cA gA + wB/cB gB + wE/cE gE ) ) + crash/match
// ------------------------------------------------------------------------------------------------------
__global__ void finalize_kernel(LossArgs a) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.NA * a.FT) return;
  const int ag = idx / a.FT, t = idx % a.FT;
  const int g = a.cfg.group_of[ag];
  const double* acc = a.ws.acc + (size_t)g * A_N;
  const bool adv = (a.cfg.kind & STRIVE_LOSS_ADV) != 0;
  float out[4] = {0.f, 0.f, 0.f, 0.f};
  if (a.cfg.kind & (STRIVE_LOSS_AVOID | STRIVE_LOSS_ADV)) {
    const float sA = (a.cfg.w_coll_veh > 0.f && acc[A_CNTA] > 0.0) ? (float)(a.cfg.w_coll_veh / acc[A_CNTA]) : 0.f;
    const float sB = (adv && a.cfg.w_coll_veh_plan > 0.f && acc[A_CNTB] > 0.0) ? (float)(a.cfg.w_coll_veh_plan / acc[A_CNTB]) : 0.f;
    const float sE = (a.cfg.w_coll_env > 0.f && acc[A_CNTE] > 0.0) ? (float)(a.cfg.w_coll_env / acc[A_CNTE]) : 0.f;
    const int lo = max(0, 3 * t - 3), hi = min(a.T3 - 1, 3 * t + 5);
    for (int o = lo; o <= hi; o++) {
      int i0, i1;
      float l0, l1;
      interp_src(o, a.FT, i0, i1, l0, l1);
      float wgt = 0.f;
      if (i0 == t) wgt += l0;
      if (i1 == t) wgt += l1;
      if (wgt == 0.f) continue;
      const size_t k = (size_t)ag * a.T3 + o;
      const float4 ga = *reinterpret_cast<const float4*>(a.ws.gA + k * 4);
      float d[4] = {sA * ga.x, sA * ga.y, sA * ga.z, sA * ga.w};
      if (adv) {
        const float4 gb = *reinterpret_cast<const float4*>(a.ws.gB + k * 4);
        d[0] += sB * gb.x; d[1] += sB * gb.y; d[2] += sB * gb.z; d[3] += sB * gb.w;
      }
      d[0] += sE * a.ws.gE[k * 2];
      d[1] += sE * a.ws.gE[k * 2 + 1];
      // heading renormalisation adjoint: h = u/|u|
      const float4 tiv = *reinterpret_cast<const float4*>(a.ws.ti + k * 4);
      const float dot = tiv.z * d[2] + tiv.w * d[3];
      const float inv = 1.0f / a.ws.un[k];
      const float du2 = (d[2] - tiv.z * dot) * inv, du3 = (d[3] - tiv.w * dot) * inv;
      out[0] += wgt * d[0]; out[1] += wgt * d[1]; out[2] += wgt * du2; out[3] += wgt * du3;
    }
    if (adv && a.cfg.w_adv_crash > 0.f) {
      const int gs0 = a.scene_of[a.cfg.group_agent_ptr[g]];
      const int gs1 = a.scene_of[a.cfg.group_agent_ptr[g + 1] - 1] + 1;
      const float sc = a.cfg.w_adv_crash / (float)(gs1 - gs0);     // mean over the B scenes of the batch (:250)
      out[0] += sc * a.ws.gC[(size_t)idx * 2];
      out[1] += sc * a.ws.gC[(size_t)idx * 2 + 1];
    }
#pragma unroll
    for (int k = 0; k < 4; k++) a.d_traj[(size_t)idx * 4 + k] = out[k] * out_scale(a, k);
  }
  if (a.cfg.kind & STRIVE_LOSS_MATCH) {
    float dm[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.match_mask[idx] != 0) {
      float p[4], q[4];
      unnorm4(a, a.traj + (size_t)idx * 4, p);
      unnorm4(a, a.match_tgt + (size_t)idx * 4, q);
      float w = 0.f;
      if (a.cfg.w_match_ext > 0.f) w += a.cfg.w_match_ext;
      if (a.cfg.w_motion_prior_ext > 0.f) w += a.cfg.w_motion_prior_ext;      // the :46 quirk
      const float sc = 2.0f * w / (float)a.cfg.group_match_rows[g];
#pragma unroll
      for (int k = 0; k < 4; k++) dm[k] = sc * (p[k] - q[k]) * out_scale(a, k);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) a.d_traj_match[(size_t)idx * 4 + k] = dm[k];
  }
}

__global__ void terms_kernel(LossArgs a) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.cfg.num_groups) return;
  const double* acc = a.ws.acc + (size_t)g * A_N;
  float* t = a.terms + (size_t)g * STRIVE_TERMS;
  for (int k = 0; k < STRIVE_TERMS; k++) t[k] = 0.f;
  const bool adv = (a.cfg.kind & STRIVE_LOSS_ADV) != 0;
  double loss = 0.0;
  if (a.cfg.kind & (STRIVE_LOSS_AVOID | STRIVE_LOSS_ADV)) {
    const double rows = (double)a.cfg.group_zrows[g];
    if (a.cfg.w_coll_veh > 0.f) {
      const double mA = acc[A_CNTA] > 0.0 ? acc[A_SUMA] / acc[A_CNTA] : 0.0;
      t[1] = (float)mA; t[2] = (float)acc[A_CNTA];
      loss += a.cfg.w_coll_veh * mA;
    }
    if (adv && a.cfg.w_coll_veh_plan > 0.f) {
      const double mB = acc[A_CNTB] > 0.0 ? acc[A_SUMB] / acc[A_CNTB] : 0.0;
      t[7] = (float)mB; t[8] = (float)acc[A_CNTB];
      loss += a.cfg.w_coll_veh_plan * mB;
    }
    if (a.cfg.w_coll_env > 0.f) {
      const double mE = acc[A_CNTE] > 0.0 ? acc[A_SUME] / acc[A_CNTE] : 0.0;
      t[3] = (float)mE; t[4] = (float)acc[A_CNTE];
      loss += a.cfg.w_coll_env * mE;
    }
    if (a.cfg.w_motion_prior > 0.f) {
      const double mP = acc[A_PRIOR] / rows;
      t[5] = (float)mP;
      loss += adv ? mP : a.cfg.w_motion_prior * mP;     // ADV: weights already folded per row (:160-162,234)
    }
    if (a.cfg.w_init_z > 0.f) {
      if (adv) { t[6] = (float)acc[A_INIT]; loss += acc[A_INIT]; }
      else { t[6] = (float)(acc[A_INIT] / rows); loss += a.cfg.w_init_z * acc[A_INIT] / rows; }
    }
    if (adv) {
      const int gs0 = a.scene_of[a.cfg.group_agent_ptr[g]];
      const int gs1 = a.scene_of[a.cfg.group_agent_ptr[g + 1] - 1] + 1;
      const double mC = acc[A_CRASH] / (double)(gs1 - gs0);
      t[9] = (float)mC;
      loss += a.cfg.w_adv_crash * mC;
    }
    t[0] = (float)loss;
  }
  if (a.cfg.kind & STRIVE_LOSS_MATCH) {
    const double mM = acc[A_MATCH] / (double)a.cfg.group_match_rows[g];
    t[10] = (float)mM;
    double lm = 0.0;
    if (a.cfg.w_match_ext > 0.f) lm += a.cfg.w_match_ext * mM;
    if (a.cfg.w_motion_prior_ext > 0.f) lm += a.cfg.w_motion_prior_ext * mM;   // :46
    t[11] = (float)lm;
  }
}

// ------------------------------------------------------------------------------------------------------
extern "C" int strive_loss_fwd_bwd(const StriveLossCfg* cfg, const StriveScene* sc, const StriveMap* map, int32_t ft,
                                   const float* traj, const float* z, const float* prior_mu, const float* prior_var,
                                   const float* init_z, const uint8_t* z_mask, const float* match_tgt,
                                   const uint8_t* match_mask, const float* adv_tgt, float* d_traj, float* d_traj_match,
                                   float* d_z_direct, float* terms, void* workspace, int64_t workspace_bytes, void* stream_) {
  STRIVE_CHECK(cfg && sc && map && traj && terms && workspace, STRIVE_EINVAL, "strive_loss_fwd_bwd: null argument");
  const int kind = cfg->kind;
  const bool main_term = (kind & (STRIVE_LOSS_AVOID | STRIVE_LOSS_ADV)) != 0;
  STRIVE_CHECK(!((kind & STRIVE_LOSS_AVOID) && (kind & STRIVE_LOSS_ADV)), STRIVE_EINVAL, "AVOID and ADV are exclusive");
  STRIVE_CHECK(kind != 0, STRIVE_EINVAL, "loss kind is empty");
  if (main_term) {
    STRIVE_CHECK(d_traj && d_z_direct && z && prior_mu && prior_var, STRIVE_EINVAL, "loss: missing latent/gradient buffers");
    STRIVE_CHECK(cfg->w_init_z <= 0.f || init_z != nullptr, STRIVE_EINVAL, "loss: init_z weight > 0 but init_z is null");
    STRIVE_CHECK(cfg->cblock_ptr && cfg->cblock_of && cfg->circ_cx && cfg->lw_un && cfg->group_zrows, STRIVE_EINVAL, "loss: missing cfg arrays");
    STRIVE_CHECK(cfg->w_coll_env <= 0.f || (cfg->env_L && cfg->env_W && cfg->env_lin_l && cfg->env_lin_w), STRIVE_EINVAL, "loss: env grid missing");
  }
  if (kind & STRIVE_LOSS_ADV) STRIVE_CHECK(adv_tgt != nullptr || cfg->adv_own_pred, STRIVE_EINVAL, "ADV loss needs adv_tgt (or adv_own_pred)");
  if (kind & STRIVE_LOSS_MATCH)
    STRIVE_CHECK(match_tgt && match_mask && d_traj_match && cfg->group_match_rows, STRIVE_EINVAL, "MATCH loss needs target, mask, rows");
  STRIVE_CHECK(cfg->group_of && cfg->group_agent_ptr && cfg->num_groups > 0, STRIVE_EINVAL, "loss: group arrays missing");
  STRIVE_CHECK(!main_term || cfg->w_coll_env <= 0.f || cfg->agent_map != nullptr, STRIVE_EINVAL, "loss: agent_map missing");
  cudaStream_t stream = (cudaStream_t)stream_;
  LossArgs a;
  a.cfg = *cfg;
  a.map = *map;
  a.NA = sc->num_agents; a.S = sc->num_scenes; a.FT = ft; a.T3 = 3 * ft;
  const int64_t need = loss_carve(&a.ws, (char*)workspace, a.NA, ft, cfg->num_groups);
  STRIVE_CHECK(workspace_bytes >= need, STRIVE_ESIZE, "loss workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)need);
  a.ptr = sc->ptr; a.scene_of = sc->scene_of; a.map_idx = sc->map_idx;
  a.traj = traj; a.z = z; a.prior_mu = prior_mu; a.prior_var = prior_var; a.init_z = init_z; a.z_mask = z_mask;
  a.match_tgt = match_tgt; a.match_mask = match_mask; a.adv_tgt = adv_tgt;
  a.d_traj = d_traj; a.d_traj_match = d_traj_match; a.d_z = d_z_direct; a.terms = terms;
  const int NA = a.NA, T3 = a.T3;
  {
    const int nz = max(NA, cfg->num_groups * A_N);
    KPROF("loss_zero", stream, loss_zero_kernel<<<(nz + 255) / 256, 256, 0, stream>>>(a));
    STRIVE_LAUNCH_CHECK();
  }
  if (main_term) {
    KPROF("interp", stream, interp_kernel<<<(NA * T3 + 255) / 256, 256, 0, stream>>>(a));
    STRIVE_LAUNCH_CHECK();
    if (kind & STRIVE_LOSS_ADV) {
      KPROF("adv_dist", stream, adv_dist_kernel<<<(NA + 127) / 128, 128, 0, stream>>>(a));
      STRIVE_LAUNCH_CHECK();
      KPROF("adv_crash", stream, adv_crash_kernel<<<a.S, 256, 0, stream>>>(a, cfg->adv_min_out));
      STRIVE_LAUNCH_CHECK();
    }
    if (cfg->w_coll_veh > 0.f || ((kind & STRIVE_LOSS_ADV) && cfg->w_coll_veh_plan > 0.f)) {
      KPROF("veh_coll", stream, veh_coll_kernel<<<(NA * T3 + 127) / 128, 128, 0, stream>>>(a));
      STRIVE_LAUNCH_CHECK();
    } else {
      STRIVE_CUDA(cudaMemsetAsync(a.ws.gA, 0, (size_t)NA * T3 * 16, stream));
      STRIVE_CUDA(cudaMemsetAsync(a.ws.gB, 0, (size_t)NA * T3 * 16, stream));
    }
    if (cfg->w_coll_env > 0.f) {
      const long long threads = (long long)NA * T3 * 32;
      KPROF("env_coll", stream, env_coll_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(a));
      STRIVE_LAUNCH_CHECK();
    } else {
      STRIVE_CUDA(cudaMemsetAsync(a.ws.gE, 0, (size_t)NA * T3 * 8, stream));
    }
    KPROF("latent_terms", stream, latent_terms_kernel<<<(NA * 32 + 255) / 256, 256, 0, stream>>>(a));
    STRIVE_LAUNCH_CHECK();
  }
  if (kind & STRIVE_LOSS_MATCH) {
    KPROF("match_sum", stream, match_sum_kernel<<<(NA * ft + 255) / 256, 256, 0, stream>>>(a));
    STRIVE_LAUNCH_CHECK();
  }
  KPROF("finalize", stream, finalize_kernel<<<(NA * ft + 255) / 256, 256, 0, stream>>>(a));
  STRIVE_LAUNCH_CHECK();
  KPROF("terms", stream, terms_kernel<<<(cfg->num_groups + 63) / 64, 64, 0, stream>>>(a));
  STRIVE_LAUNCH_CHECK();
  return 0;
}
