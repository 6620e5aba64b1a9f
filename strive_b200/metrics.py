"""Success / plausibility checks that bracket the latent loop in the reference drivers, on the GPU
through the C-ABI (csrc/metrics.cu).  Same call surface as the reference functions they replace:

  compute_coll_rate_env      src/losses/traffic_model.py:366-419      (check_on_layer, datasets/nuscenes_utils.py:266-298)
  check_single_veh_coll      src/losses/adv_gen_nusc.py:517-565        (shapely IoU of get_corners rectangles)
  check_pairwise_veh_coll    src/losses/adv_gen_nusc.py:567-623
  determine_feasibility_nusc src/utils/scenario_gen.py:30-107          (check_line_layer, nuscenes_utils.py:300-333)

The reference passes `map_env.nusc_raster[:, layer]` + `map_env.nusc_dx` to the two raster helpers; here they take the
MapEnv and the layer index (the kernels read the resident raster in place).  No CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi
from .runtime import MapEnv

VEH_COLL_THRESH = 0.02   # adv_gen_nusc.py:515 / losses/traffic_model.py:18
ENV_COLL_THRESH = 0.05   # losses/traffic_model.py:17


def _env(map_env):
    if not isinstance(map_env, MapEnv):
        raise RuntimeError('strive_b200: map_env must be a strive_b200.MapEnv')
    return map_env


def _f32(t, dev):
    return t.detach().to(dev, torch.float32).contiguous()


def check_on_layer(map_env, layer, cars, lw, mapixes):
    """nutils.check_on_layer (nuscenes_utils.py:266-298): cars (B,4) and lw (B,2) UNNORMALISED, mapixes (B,) -> (B,) float
    fraction of the car footprint on pixels of `layer` that are 1."""
    env = _env(map_env)
    dev = env.nusc_raster.device
    cars, lw = _f32(cars, dev), _f32(lw, dev)
    B = cars.size(0)
    if B == 0:
        return torch.zeros(0, dtype=torch.float32, device=dev)
    mdx = torch.mean(env.nusc_dx)                                 # :277 float64, all maps and both axes
    mlw = torch.mean(lw, dim=0)                                   # :278
    L = int(torch.round(mlw[0] / mdx).int().item())               # :279-280
    W = int(torch.round(mlw[1] / mdx).int().item())
    if L < 1 or W < 1:
        raise RuntimeError('strive_b200.check_on_layer: degenerate footprint grid %d x %d' % (L, W))
    lin_l = torch.linspace(-1.0, 1.0, L, device=dev)              # gen_car_coords ls/ws branch, :222-224
    lin_w = torch.linspace(-1.0, 1.0, W, device=dev)
    out = torch.empty(B, dtype=torch.float32, device=dev)
    mo = mapixes.detach().to(dev, torch.int32).contiguous()
    with torch.cuda.device(out.device):
        _cabi.check(_cabi.lib().strive_on_layer_frac(C.byref(env.cstruct), int(layer), _cabi.dptr(cars), _cabi.dptr(lw), _cabi.dptr(mo),
                                                     _cabi.dptr(lin_l), _cabi.dptr(lin_w), L, W, B, _cabi.dptr(out), _cabi.stream_ptr()))
    return out


def compute_coll_rate_env(scene_graph, map_idx, pred, map_env, state_normalizer, att_normalizer, ego_only=False):
    """losses/traffic_model.py:366-419: which (agent, sample) rollouts leave the drivable layer by more than 5 % of the
    footprint at any step.  pred: (NA,NS,FT,4) NORMALISED (or a dict with 'future_pred').  NaN frames never collide."""
    env = _env(map_env)
    dev = env.nusc_raster.device
    pred_future = pred if isinstance(pred, torch.Tensor) else pred['future_pred']
    NA, NS, FT, _ = pred_future.size()
    veh_att = scene_graph.lw
    mapixes = map_idx[scene_graph.batch]
    if ego_only:
        ego_inds = scene_graph.ptr[:-1].long()                                         # datasets/utils.py get_ego_inds
        pred_future, veh_att, mapixes = pred_future[ego_inds], veh_att[ego_inds], mapixes[ego_inds]
        NA = pred_future.size(0)
    pred_un = state_normalizer.unnormalize(pred_future).reshape(NA * NS * FT, 4)
    att_un = att_normalizer.unnormalize(veh_att).view(NA, 1, 1, 2).expand(NA, NS, FT, 2).reshape(NA * NS * FT, 2)
    mapixes = mapixes.view(NA, 1, 1).expand(NA, NS, FT).reshape(NA * NS * FT)
    valid = ~torch.isnan(pred_un.sum(-1))                                               # :399-400
    frac = torch.ones(NA * NS * FT, dtype=torch.float32, device=dev)
    if bool(valid.any()):
        # the reference hands ONLY the valid rows to check_on_layer, so the batch-global grid comes from their mean lw (:401-405)
        frac[valid.to(dev)] = check_on_layer(env, 0, pred_un[valid], att_un[valid], mapixes[valid])
    frac = frac.view(NA, NS, FT)
    coll_frame = frac < (1.0 - ENV_COLL_THRESH)                                         # :410
    map_coll = torch.sum(coll_frame, dim=2) >= 1
    return {'num_coll_map': float(torch.sum(map_coll).item()), 'num_traj_map': float(NS * NA), 'did_collide': map_coll}


def check_line_layer(map_env, layer, start, end, mapixes):
    """nutils.check_line_layer (nuscenes_utils.py:300-333): (B,) bool, True where the segment start->end (UNNORMALISED xy)
    touches a 0 pixel of `layer`."""
    env = _env(map_env)
    dev = env.nusc_raster.device
    start, end = _f32(start, dev), _f32(end, dev)
    B = start.size(0)
    if B == 0:
        return torch.zeros(0, dtype=torch.bool, device=dev)
    line_len = torch.norm(start - end, dim=-1)
    mdx = torch.mean(env.nusc_dx)
    NL = int(torch.max(torch.round(line_len / mdx).int()).item())                       # :316-318
    lin01 = torch.linspace(0.0, 1.0, max(NL, 1), device=dev)                            # :320
    hit = torch.zeros(B, dtype=torch.uint8, device=dev)
    oob = torch.zeros(1, dtype=torch.int32, device=dev)
    mo = mapixes.detach().to(dev, torch.int32).contiguous()
    with torch.cuda.device(hit.device):
        _cabi.check(_cabi.lib().strive_line_layer(C.byref(env.cstruct), int(layer), _cabi.dptr(start), _cabi.dptr(end), _cabi.dptr(mo),
                                                  _cabi.dptr(lin01), NL, B, _cabi.dptr(hit), _cabi.dptr(oob), _cabi.stream_ptr()))
    if int(oob.item()) != 0:
        raise RuntimeError('strive_b200.check_line_layer: %d samples index outside the raster '
                           '(the reference raises IndexError here, nuscenes_utils.py:329)' % int(oob.item()))
    return hit.bool()


def _iou_hits(traj_a, lw_a, traj_b, lw_b, want_iou=False):
    dev = traj_a.device if traj_a.is_cuda else torch.device('cuda', torch.cuda.current_device())
    ta, la, tb, lb = _f32(traj_a, dev), _f32(lw_a, dev), _f32(traj_b, dev), _f32(lw_b, dev)
    na, T = ta.size(0), ta.size(1)
    nb = tb.size(0)
    hit = torch.zeros((na, nb, T), dtype=torch.uint8, device=dev)
    iou = torch.zeros((na, nb, T), dtype=torch.float32, device=dev) if want_iou else None
    with torch.cuda.device(hit.device):
        _cabi.check(_cabi.lib().strive_veh_iou_hits(_cabi.dptr(ta), _cabi.dptr(la), na, _cabi.dptr(tb), _cabi.dptr(lb), nb, T,
                                                    float(VEH_COLL_THRESH), _cabi.dptr(hit), _cabi.dptr(iou), _cabi.stream_ptr()))
    return hit, iou


def check_single_veh_coll(traj_tgt, lw_tgt, traj_others, lw_others):
    """adv_gen_nusc.py:517-565.  traj_tgt (T,4), lw_tgt (2,), traj_others (N,T,4), lw_others (N,2), all UNNORMALISED.
    Returns numpy (veh_coll (N,) bool, coll_time (N,) int): first step at which IoU > 0.02, FT if none; NaN frames skipped."""
    N, FT, _ = traj_others.size()
    if N == 0:
        return np.zeros((0,), dtype=bool), np.zeros((0,), dtype=int)
    hit, _ = _iou_hits(traj_tgt.unsqueeze(0), lw_tgt.reshape(1, 2), traj_others, lw_others)
    hit = hit[0].bool()                                                                 # (N, T)
    any_hit = hit.any(dim=1)
    first = torch.where(any_hit, hit.float().argmax(dim=1), torch.full((N,), FT, device=hit.device, dtype=torch.long))
    return any_hit.cpu().numpy().astype(bool), first.cpu().numpy().astype(int)


def check_pairwise_veh_coll(traj, lw):
    """adv_gen_nusc.py:567-623.  traj (N,T,4), lw (N,2) UNNORMALISED.  The reference scans pairs (ai < aj) and flags only ai,
    once (:585-586, :609-612): did_collide[ai] = any collision with a LATER agent; num_coll_veh = number of flagged agents."""
    N = traj.size(0)
    if N == 0:
        return {'num_coll_veh': 0.0, 'num_traj_veh': 0.0, 'did_collide': np.zeros((0,), dtype=bool)}
    hit, _ = _iou_hits(traj, lw, traj, lw)
    pair = hit.bool().any(dim=2)                                                        # (N, N)
    later = torch.triu(torch.ones(N, N, dtype=torch.bool, device=pair.device), diagonal=1)
    did = (pair & later).any(dim=1)
    return {'num_coll_veh': float(did.sum().item()), 'num_traj_veh': float(N), 'did_collide': did.cpu().numpy().astype(bool)}


def determine_feasibility_nusc(samples, normalizer, feasibility_thresh, feasibility_time=0, feasibility_vel=0.0,
                               feasibility_infront_min=None, check_non_drivable_separation=True, map_env=None, map_idx=None):
    """utils/scenario_gen.py:30-107 for ONE scene graph (row 0 = ego).  samples (NA,NS,FT,4) NORMALISED.
    Returns (feasible (NA-1,) bool, feasible_time_step (NA-1,), feasible_dist (NA-1,)) or (None, None, None) for an ego-only scene.
    Device tensor arithmetic in the reference's order; the drivable-separation test runs in csrc/metrics.cu."""
    if samples.size(0) == 1:
        return None, None, None
    samples = normalizer.unnormalize(samples)
    ego, agents = samples[0:1], samples[1:]
    NA, NS, FT, _ = agents.size()
    d = torch.norm(ego[:, :, :, :2] - agents[:, :, :, :2], dim=-1)[:, :, feasibility_time:]
    if feasibility_infront_min is not None:
        assert -1 <= feasibility_infront_min <= 1
        ego_h = ego[:, :, feasibility_time:, 2:4]
        e2a = agents[:, :, feasibility_time:, :2] - ego[:, :, feasibility_time:, :2]
        e2a = e2a / torch.norm(e2a, dim=-1, keepdim=True)
        infront = torch.sum(e2a * ego_h, dim=-1) >= feasibility_infront_min
        d = d.clone()
        d[~infront] = float('inf')
    min_samp_d, min_samp_i = torch.min(d, dim=1)                                         # (NA, T')
    feasible_dist, ft_step = torch.min(min_samp_d, dim=1)
    ft_step = ft_step + feasibility_time
    feasible = (d < feasibility_thresh).sum(dim=[1, 2]) > 0
    if check_non_drivable_separation:
        ar = torch.arange(NA, device=samples.device)
        msi = min_samp_i[ar, ft_step - feasibility_time]
        a_xy = agents[ar, msi][ar, ft_step][:, :2]
        e_xy = ego.expand(NA, NS, FT, 4)[ar, msi][ar, ft_step][:, :2]
        sep = check_line_layer(map_env, 0, a_xy, e_xy, map_idx.expand(NA))
        feasible = torch.logical_and(feasible, ~sep.to(feasible.device))
    vel = torch.norm(agents[:, :, 1:, :2] - agents[:, :, :-1, :2], dim=-1)
    max_vel = torch.max(torch.max(vel, dim=1)[0], dim=1)[0]
    feasible = torch.logical_and(feasible, max_vel > feasibility_vel)
    return feasible, ft_step, feasible_dist
