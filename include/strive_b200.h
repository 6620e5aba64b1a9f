/*
 * strive_b200 -- C ABI of the B200-native STRIVE latent-optimisation hot path.
 *
 * The reference (nv-tlabs/STRIVE) has no native/FFI boundary: its hot path is the Python method surface
 *   TrafficModel.decode_embedding                 src/models/traffic_model.py:405-414 (-> 589-704)
 *   TrafficModel.encode_map                       src/models/traffic_model.py:416-451
 *   AvoidCollLoss / AdvGenLoss / TgtMatchingLoss  src/losses/adv_gen_nusc.py:14-341
 *   torch.optim.Adam on z                         src/refine_traffic_optim.py:163-218, utils/*_optim.py
 * These entry points are what a ctypes binding under those methods calls (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller (PyTorch) owns all memory,
 *     including workspaces; the library keeps no state between calls except the opaque model handle.
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*), one device per process.
 *   - return 0 on success; otherwise an error code and strive_last_error() describes it (the Python wrapper
 *     raises RuntimeError so the drivers' `except RuntimeError` batch-skip keeps working,
 *     src/refine_traffic_optim.py:381-388).
 *   - float = fp32, state/feature layouts are row-major exactly as the reference tensors.
 */
#ifndef STRIVE_B200_H
#define STRIVE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STRIVE_ABI_VERSION 1

typedef struct StriveModel StriveModel;

/* ---- error handling ---------------------------------------------------------------------------------- */
const char* strive_last_error(void);
int strive_abi_version(void);
/* sizeof/offsetof table of the ABI structs so foreign bindings can verify their layout (returns count) */
int strive_struct_layout(int64_t* out, int max_n);

/* ---- per-kernel timing (bench.py) ----------------------------------------------------------------------
 * When enabled every kernel launch is bracketed by CUDA events on its stream; strive_profile_report
 * synchronises and writes "kernel_name launches total_ms" lines. Not for use under CUDA-graph capture. */
int strive_profile_enable(int on);
int64_t strive_profile_report(char* buf, int64_t cap);

/* ---- tcgen05 primitive self-test (tests only): A (128,32), B (32,32), X0/X1 (144,8) fp32 (rounded to bf16 inside);
 * D0 = A B^T, D1[m] = [X0[m+1] | X1[m+3]] B[:, :16]^T, both (128,32) fp32. */
int strive_tc_selftest(const float* A, const float* B, const float* X0, const float* X1, float* D0, float* D1, void* stream);
/* CTA-pair variant (tcgen05 cta_group::2, cluster of two CTAs): A (256,32), B (n,32) fp32 (rounded to bf16 inside), D = A B^T (256,n), n = 64 or 128;
 * CTA r stages rows [128 r, +128) of A and rows [n/2 r, +n/2) of B; flag (device int32): bit r set = CTA r timed out (bounded waits). */
int strive_tc_selftest_pair(const float* A, const float* B, float* D, int32_t* flag, int32_t n, void* stream);
/* Pipeline diagnostics of the tensor-core convolutions: out32 = [4 kernels conv1..conv4][8] SM-cycle counters summed over
 * CTAs since the last reset ([0] producers blocked on a ring slot, [1] producer total, [2] MMA warp starved, [3] MMA warp
 * blocked on an accumulator, [4] MMA total, [5] epilogue idle, [6] epilogue total, [7] CTAs).  No reference counterpart. */
int strive_tc_trace(unsigned long long* out32, int reset);
/* Timing experiments only: bit0/1/2/3 disable the epilogue stores / producer shared stores / producer global loads / MMAs of
 * conv2..conv4 (results become garbage).  0 = normal operation. */
int strive_tc_debug(int flags);

/* ---- model weights -------------------------------------------------------------------------------------
 * Replaces: torch state_dict of decoder_net.*, decoder_memory.*, map_conv.*, map_feature.* loaded by
 * utils/torch.py:32-60 (load_state).  `blob` is a device buffer packed by strive_b200/weights.py in the
 * segment order returned by strive_model_layout(); seg_sizes_host (n_segs int64) is checked against it. */
int strive_model_layout(int num_classes, int64_t* seg_sizes_host, int max_segs, int* n_segs_out);
int strive_model_create(const float* blob, int64_t blob_floats, const int64_t* seg_sizes_host, int n_segs,
                        int num_classes, StriveModel** out);
void strive_model_destroy(StriveModel* m);
/* bf16 hi/lo split conv1..conv4 weights in the tcgen05 K-major operand layout (packed by strive_b200/weights.py
 * pack_tc_weights; derived from the same map_conv.* tensors).  Without it the map encoder refuses the tensor-core path. */
int64_t strive_model_tc_bytes(void);
int strive_model_set_tc_weights(StriveModel* m, const void* blob, int64_t bytes);
/* 1 = tensor-core map encoder (default), 0 = fp32 SIMT kernels (A/B verification only) */
int strive_mapenc_set_impl(int impl);
/* 1: a chunk of >= 512 crops runs as two half-chunks on two streams (crop gather of one half overlaps conv1 of the other);
 * 0 (default): one stream.  Measured +0.1 % on a power-capped B200, kept as an A/B switch. */
int strive_mapenc_set_split(int on);
/* conv3 of the tensor-core encoder on CTA pairs (tcgen05 cta_group::2, clusters of two CTAs): 1 = on (default), 0 = single-CTA kernel
 * (A/B; also the automatic fallback when the device cannot hold ~74 CTA pairs at once).  The two kernels agree to fp32 re-association
 * level (1e-5 on the features); each is bitwise reproducible and batch-invariant.  2 = pairs even when fewer of them fit (tools). */
int strive_mapenc_set_pair(int on);
/* Edge MLP of the decoder GNN (interaction_net.py:139-184) on the warp-level tensor path: the library packs the edge-MLP
 * matrices of the weight blob into mma.sync fragment order inside `buf` (device, 16-byte aligned,
 * strive_model_edge_frag_bytes() bytes, owned by the caller for the lifetime of the model).  Without it -- or with
 * strive_edge_set_impl(0) -- the fp32 SIMT edge kernels run (A/B verification). */
int64_t strive_model_edge_frag_bytes(void);
int strive_model_set_edge_frags(StriveModel* m, void* buf, int64_t bytes, void* stream);
int strive_edge_set_impl(int impl);
/* Programmatic dependent launch.  mask bit 0: rollout kernels (griddepcontrol.wait is their first instruction: only the launch
 * latency overlaps), bit 1: map-encoder kernels (their prologue -- weights to shared memory, barrier init, TMEM alloc --
 * overlaps the tail of the predecessor; the pointers to the predecessor's output pass through the wait).  Default 3;
 * 0 = plain stream-ordered launches.  Results are identical (tests/test_gpu_fullsize.py checks run-to-run determinism). */
int strive_set_pdl(int on);

/* ---- scene description (all device) --------------------------------------------------------------------
 * Mirrors the torch_geometric Batch the drivers build (src/datasets/nuscenes_dataset.py:609-687): edges are
 * the full directed clique inside each scene, so only `ptr` is needed. */
typedef struct StriveScene {
  int32_t num_agents;          /* NA */
  int32_t num_scenes;          /* S  */
  int32_t max_scene_agents;    /* host-known max(ptr[s+1]-ptr[s]); must be <= 255 */
  int32_t num_classes;         /* NC */
  const int32_t* ptr;          /* (S+1) */
  const int32_t* scene_of;     /* (NA) scene id per agent (= Batch.batch) */
  const int32_t* map_idx;      /* (S) map id per scene */
  const float* past_last;      /* (NA,6) normalised last past state  = scene_graph.past[:, -1, :] */
  const float* lw;             /* (NA,2) normalised length/width */
  const float* sem;            /* (NA,NC) one-hot class */
} StriveScene;

typedef struct StriveMap {
  const uint8_t* raster;       /* (M,C,H,W) uint8 = NuScenesMapEnv.nusc_raster, src/datasets/map_env.py:165 */
  const double* dx;            /* (M,2) float64 metres/pixel = nusc_dx, map_env.py:166 */
  int32_t M, C, H, W;
  const float* lin_l;          /* (256) torch.linspace(bounds[0], bounds[2], 256) float32, nuscenes_utils.py:219 */
  const float* lin_w;          /* (256) torch.linspace(bounds[1], bounds[3], 256) float32, nuscenes_utils.py:220 */
  const uint8_t* packed;       /* (M,H,packed_pitch) uint8: bit c = (raster[m,c,y,x] != 0); the raster must be binary (nuScenes get_map_mask) */
  int32_t packed_pitch;        /* bytes per row of `packed`: >= W, a multiple of 16 (rows are staged with 16-byte cp.async), base 16-byte aligned */
} StriveMap;

/* ---- map encoder ---------------------------------------------------------------------------------------
 * Replaces TrafficModel.encode_map + NuScenesMapEnv.get_map_crop + get_map_obs + map_conv + map_feature.
 * pose_un: (N,4) UNNORMALISED (x,y,hx,hy); map_of: (N) map id per pose; out: (N,64).
 * workspace: strive_mapenc_workspace_bytes(N) bytes. */
int64_t strive_mapenc_workspace_bytes(int32_t n);
int strive_mapenc_fwd(const StriveModel* m, const StriveMap* map, const float* pose_un, const int32_t* map_of,
                      int32_t n, float* out_feat, void* workspace, int64_t workspace_bytes, void* stream);
/* test hook: the gathered crop itself, (N,4,256,256) uint8 (reference get_map_obs output) */
int strive_map_crop(const StriveMap* map, const float* pose_un, const int32_t* map_of, int32_t n,
                    uint8_t* out_crop, void* stream);

/* ---- decoder rollout -----------------------------------------------------------------------------------
 * Replaces TrafficModel.decode_embedding -> autoregressive_decoder (traffic_model.py:589-704), single-sample
 * branch (z (NA,32)); the NS=1 3-D z of sol_optim.py:38-44 is the same computation on a view.
 *   z, map_feat0, past_feat0: (NA,32),(NA,64),(NA,64); ext_future: (S,FT,4) normalised or NULL
 *   traj_out: (NA,FT,4) normalised global (x,y,hx,hy) = 'future_pred'
 *   tape: strive_decode_tape_bytes(NA,FT) bytes, consumed by strive_decode_bwd.
 * strive_decode_bwd may be called several times per forward with different d_traj seeds (adv/sol loops decode
 * twice with identical values and different detach masks, utils/adv_gen_optim.py:119-130); d_z is OVERWRITTEN. */
int64_t strive_decode_tape_bytes(int32_t num_agents, int32_t ft);
int strive_decode_fwd(const StriveModel* m, const StriveScene* sc, const StriveMap* map, const float* z,
                      const float* map_feat0, const float* past_feat0, const float* ext_future, int32_t ft,
                      float* traj_out, void* tape, int64_t tape_bytes, void* stream);
int strive_decode_bwd(const StriveModel* m, const StriveScene* sc, int32_t ft, const float* ext_future,
                      const float* d_traj, float* d_z, void* tape, int64_t tape_bytes, void* stream);
/* Two sweeps of the same forward tape in one call: (d_traj_a -> d_z_a) and (d_traj_b -> d_z_b), results identical to two
 * strive_decode_bwd calls.  Replaces the two backward passes of adv_gen_optim.py:119-130,170-171 / sol_optim.py:75-86,108-109
 * (one seed per latent group).  The second sweep runs on an internal side stream with its own carry buffers inside the tape
 * (fork / join by events on `stream`; capturable), so the pair overlaps on the device. */
int strive_decode_bwd_pair(const StriveModel* m, const StriveScene* sc, int32_t ft, const float* ext_future,
                           const float* d_traj_a, float* d_z_a, const float* d_traj_b, float* d_z_b,
                           void* tape, int64_t tape_bytes, void* stream);
/* test hook: copies one named tape tensor of step t to `out` (float). names: x,P,Q,aggr,past_feat,map_feat,prev,pos,loc,mem;
 * "arg" copies the (NA,64) uint8 arg-max routing table of the max aggregation (local source index in the scene, 255 = none). */
int strive_decode_tape_read(const void* tape, int32_t num_agents, int32_t ft, const char* name, int32_t t,
                            float* out, void* stream);

/* ---- losses --------------------------------------------------------------------------------------------
 * One fused forward+backward evaluation of the reference loss modules on a decoded future.
 * Loss-normalisation groups = the batches the reference driver would have formed (every .mean() in
 * adv_gen_nusc.py is over one such batch); a group is a contiguous range of scenes (hence of agents).
 * Collision blocks = the sets of agents that may collide with each other: the whole group for
 * AvoidCollLoss(ptr=None) as built by refine_traffic_optim.py:176-181; one scene when ptr is given
 * (adv_gen_optim.py:76-84, sol_optim.py:57-63).  Blocks are contiguous agent ranges. */
#define STRIVE_LOSS_AVOID 1   /* AvoidCollLoss.forward      adv_gen_nusc.py:303-341 */
#define STRIVE_LOSS_ADV   2   /* AdvGenLoss.forward         adv_gen_nusc.py:93-262  */
#define STRIVE_LOSS_MATCH 4   /* TgtMatchingLoss.forward    adv_gen_nusc.py:27-51 (incl. the :46 quirk) */

typedef struct StriveLossCfg {
  int32_t kind;                /* bitmask of STRIVE_LOSS_*; AVOID and ADV are mutually exclusive */
  int32_t traj_unnormalized;   /* 1: traj/targets/d_traj are in UNNORMALISED units (the reference loss modules' own
                                  calling convention); 0: normalised decoder output (fused loop) */
  int32_t num_groups;
  const int32_t* group_agent_ptr; /* (G+1) agent ranges of the groups */
  const int32_t* group_of;     /* (NA) group per agent */
  const int32_t* agent_map;    /* (NA) map id per agent (= mapixes of the reference loss constructors) */
  const int32_t* group_zrows;  /* (G) number of latent rows that enter the prior/init means of the group */
  const int32_t* group_match_rows; /* (G) number of (agent,t) rows in the match mean, or NULL */
  const int32_t* cblock_ptr;   /* (NB+1) agent ranges of collision blocks */
  const int32_t* cblock_of;    /* (NA) collision block per agent */
  float w_coll_veh, w_coll_env, w_motion_prior, w_init_z;
  float w_coll_veh_plan, w_init_z_atk, w_motion_prior_atk, w_adv_crash, w_match_ext, w_motion_prior_ext;
  float veh_coll_buffer;
  int32_t single_veh_idx;      /* -1 = all agents; k = only pairs/rows involving agent ptr[s]+k (sol_optim.py:57-63) */
  int32_t crash_min_t;
  int32_t use_infront;         /* crash_loss_min_infront is not None */
  float crash_min_infront;
  const int32_t* attack_mask;  /* (NA) 1 = allowed attacker (attack_agt_idx), or NULL = all */
  int32_t* adv_min_out;        /* (S,2) OUT: (min_agt local index, min_t) of the softmin arg-max, or NULL (:137-138) */
  const int32_t* env_L;        /* (G) get_coll_point grid, nuscenes_utils.py:351-354, computed on host per group */
  const int32_t* env_W;
  const float* env_lin_l;      /* (G,128) torch.linspace(-1,1,L) zero padded */
  const float* env_lin_w;      /* (G,128) */
  const float* circ_cx;        /* (NA,5) VehCollLoss centre offsets, adv_gen_nusc.py:432-437 (torch.linspace) */
  const float* lw_un;          /* (NA,2) unnormalised length/width */
  int32_t adv_own_pred;        /* closed-loop planner mode (adv_gen_optim.py:98-103,143): the attacked trajectory is the model's OWN
                                  prediction of the target = row ptr[s] of traj (adv_tgt may be NULL) and the crash term's gradient
                                  w.r.t. it is written to those rows of d_traj; 0 = adv_tgt is an external constant */
  int32_t reserved0;
} StriveLossCfg;

#define STRIVE_TERMS 16
/* terms (G,16) float: [0] loss of the AVOID/ADV module  [1] coll_veh mean [2] coll_veh count [3] coll_env mean
 * [4] coll_env count [5] motion_prior mean [6] init term (mean for AVOID, weighted sum for ADV) [7] coll_veh_plan mean
 * [8] coll_veh_plan count [9] adv_crash mean [10] match_ext mean [11] loss of the MATCH module */
int64_t strive_loss_workspace_bytes(int32_t num_agents, int32_t ft, int32_t num_groups);
/* traj: (NA,FT,4) NORMALISED decoder output.  z/prior_mu/prior_var/init_z: (NA,32) rows in graph order; rows with
 * z_mask[a]==0 carry no latent term (z_mask NULL = all rows).  match_tgt (NA,FT,4) normalised + match_mask (NA,FT)
 * select the rows of the MATCH mean.  adv_tgt (S,FT,4) normalised = planner trajectory attacked by ADV (NULL with adv_own_pred).
 * outputs: d_traj / d_traj_match (NA,FT,4) gradients wrt the NORMALISED traj of the AVOID|ADV and MATCH modules
 * (separate seeds: the reference routes them to different latents), d_z_direct (NA,32). */
int strive_loss_fwd_bwd(const StriveLossCfg* cfg, const StriveScene* sc, const StriveMap* map, int32_t ft,
                        const float* traj, const float* z, const float* prior_mu, const float* prior_var,
                        const float* init_z, const uint8_t* z_mask, const float* match_tgt,
                        const uint8_t* match_mask, const float* adv_tgt,
                        float* d_traj, float* d_traj_match, float* d_z_direct, float* terms,
                        void* workspace, int64_t workspace_bytes, void* stream);

/* ---- fused Adam on z (torch.optim.Adam defaults: betas (0.9,0.999), eps 1e-8, no weight decay) ---------
 * g = g_a + g_b (g_b may be NULL); step_count is the 1-based step number. */
int strive_adam_step(float* z, const float* g_a, const float* g_b, float* exp_avg, float* exp_avg_sq,
                     int64_t n, int32_t step_count, float lr, float beta1, float beta2, float eps, void* stream);
/* Device-resident form used by the fused loops (replaces the torch.optim.Adam([tgt_z, other_z]).step() of
 * utils/adv_gen_optim.py:72,175, utils/sol_optim.py:47,112, utils/init_optim.py:21,57): the 0-based count of completed steps
 * lives in device memory (*step_dev, incremented by the call) so a captured CUDA graph of one iteration can be replayed;
 * the gradient of row r (row_width elements) is g_a if row_sel == NULL or row_sel[r] != 0, else g_b -- the two adjoint
 * sweeps of one rollout -- plus g_direct (NULL = none). */
int strive_adam_step_dev(float* z, const float* g_a, const float* g_b, const float* g_direct, const uint8_t* row_sel,
                         int32_t row_width, float* exp_avg, float* exp_avg_sq, int64_t n, int32_t* step_dev, float lr,
                         float beta1, float beta2, float eps, void* stream);

/* ---- success / plausibility checks around the loop (all inputs UNNORMALISED, device) -----
 * strive_on_layer_frac: nutils.check_on_layer, src/datasets/nuscenes_utils.py:266-298 (used by compute_coll_rate_env,
 *   src/losses/traffic_model.py:366-419): fraction of an L x W footprint grid of each car (cars_un (n,4), lw_un (n,2)) that
 *   reads non-zero in raster layer `layer` of map map_of[i]; L, W and the two torch.linspace(-1,1,.) tables are the
 *   batch-global values the reference derives on the host (:278-282); a car with a NaN state gets 1.0
 *   (traffic_model.py:400-408: NaN frames never count as collisions).
 * strive_line_layer: nutils.check_line_layer, nuscenes_utils.py:300-333 (determine_feasibility_nusc,
 *   src/utils/scenario_gen.py:91-99): hit[i] = 1 iff one of the num_samples points torch.linspace(0,1,.) (lin01) along
 *   start_un[i] -> end_un[i] reads 0 in `layer`.  Negative pixel indices wrap as torch indexing does; indices that would
 *   raise IndexError in the reference are counted in *oob_count (device int) and skipped.
 * strive_veh_iou_hits: the polygon test of check_single_veh_coll / check_pairwise_veh_coll, src/losses/adv_gen_nusc.py:517-623
 *   (shapely intersection / union of nutils.get_corners rectangles, nuscenes_utils.py:416-428):
 *   hit[(i*nb + j)*T + t] = IoU(rect(traj_a[i,t], lw_a[i]), rect(traj_b[j,t], lw_b[j])) > iou_thresh, 0 if either state has a
 *   NaN; iou_out (same shape, float) optional. */
int strive_on_layer_frac(const StriveMap* map, int32_t layer, const float* cars_un, const float* lw_un, const int32_t* map_of,
                         const float* lin_l, const float* lin_w, int32_t L, int32_t W, int32_t n, float* frac_out, void* stream);
int strive_line_layer(const StriveMap* map, int32_t layer, const float* start_un, const float* end_un, const int32_t* map_of,
                      const float* lin01, int32_t num_samples, int32_t n, uint8_t* hit_out, int32_t* oob_count, void* stream);
int strive_veh_iou_hits(const float* traj_a_un, const float* lw_a_un, int32_t na, const float* traj_b_un, const float* lw_b_un,
                        int32_t nb, int32_t T, double iou_thresh, uint8_t* hit_out, float* iou_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
