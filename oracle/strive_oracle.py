"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the STRIVE latent-optimisation hot path.

A from-scratch PyTorch (CPU, autograd) restatement of what the reference computes on the path
BASELINE.json's north_star names: decode(z) -> losses -> dL/dz -> Adam.  Each function cites the
reference file:line it restates (paths relative to /root/reference/src).  It is the checker for
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg ONLY; nothing
under strive_b200/ may import it, and the product path raises if the CUDA library is missing.

Pinning: the reference ships no tests / golden vectors, so this oracle is pinned against
outputs of the UNMODIFIED reference code run in the build container behind third-party shims
(oracle/ref_shims.py, generator oracle/gen_golden.py, fixtures tests/golden/*.npz); see
tests/test_oracle_golden.py.  Third-party arithmetic not under /root/reference (torch 1.9 ops,
torch-geometric 1.7.1 propagate / torch-scatter 2.0.7 max) is restated with this image's torch 2.11.

Works in float32 (reference precision) or float64 (tight reference for tolerance budgeting).
"""
import math

import torch
import torch.nn.functional as F

STATE_MEAN = (0.0, 0.0, 0.0, 0.0, 1.802009, -0.000037)   # datasets/utils.py:131-140
STATE_STD = (15.0, 15.0, 1.0, 1.0, 3.507907, 0.055684)
ATT_MEAN = (4.844294, 2.021752)
ATT_STD = (1.084860, 0.299647)
A_STATS = (0.409074, 1.045530)                           # datasets/utils.py:121-127
DDH_STATS = (0.000046, 0.075032)
DT, MAXHDOT, MAXS = 0.5, 2.0 * math.pi, 50.0
BOUNDS = (-17.0, -38.5, 60.0, 38.5)                      # datasets/map_env.py:23
CONV_K = (7, 5, 5, 3, 3, 3)                              # models/traffic_model.py:32-34


def _t(vals, like):
    return torch.tensor(vals, dtype=like.dtype, device=like.device)


def unnorm_state(x):
    """MeanStdNormalizer.unnormalize on the first x.size(-1) state dims (datasets/utils.py:91-104)."""
    d = x.size(-1)
    return x * _t(STATE_STD[:d], x) + _t(STATE_MEAN[:d], x)


def norm_state(x):
    d = x.size(-1)
    return (x - _t(STATE_MEAN[:d], x)) / _t(STATE_STD[:d], x)


def unnorm_att(x):
    return x * _t(ATT_STD, x) + _t(ATT_MEAN, x)


# --------------------------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------------------------

def transform2frame(frame, poses, inverse=False):
    """utils/transforms.py:78-139 for the (x,y,hx,hy) case. frame (B,4), poses (B,4) -> (B,4)."""
    c, s = frame[:, 2], frame[:, 3]
    pc, ps = poses[:, 2], poses[:, 3]
    if inverse:
        hx = pc * c - ps * s
        hy = ps * c + pc * s
        tx = c * poses[:, 0] - s * poses[:, 1] + frame[:, 0]
        ty = s * poses[:, 0] + c * poses[:, 1] + frame[:, 1]
    else:
        hx = pc * c + ps * s
        hy = ps * c - pc * s
        dx = poses[:, 0] - frame[:, 0]
        dy = poses[:, 1] - frame[:, 1]
        tx = c * dx + s * dy
        ty = -s * dx + c * dy
    return torch.stack([tx, ty, hx, hy], dim=1)


# --------------------------------------------------------------------------------------------
# network pieces
# --------------------------------------------------------------------------------------------

def mlp(sd, prefix, x, nlayers):
    """models/common.py:8-44: Linear, then (LayerNorm -> ReLU -> Linear) per further layer; LN eps 1e-5."""
    x = F.linear(x, sd[prefix + '.net.0.weight'], sd[prefix + '.net.0.bias'])
    idx = 1
    for _ in range(nlayers - 1):
        w, b = sd[prefix + '.net.%d.weight' % idx], sd[prefix + '.net.%d.bias' % idx]
        x = F.layer_norm(x, (x.size(-1),), w, b, 1e-5)
        x = F.relu(x)
        idx += 2
        x = F.linear(x, sd[prefix + '.net.%d.weight' % idx], sd[prefix + '.net.%d.bias' % idx])
        idx += 1
    return x


def decoder_net(sd, feat, pos, sem, edge_index, taps=None):
    """models/interaction_net.py:52-77 (SceneInteractionNet.forward) with one AgentInteractionConv round
    (:121-218) and PyG 1.7.1 max aggregation at the target node (zeros for nodes with no in-edges,
    comment at :187-188).  edge_index[0]=source j, edge_index[1]=target i."""
    p = 'decoder_net'
    x = mlp(sd, p + '.mlp_in', feat, 3)
    src, dst = edge_index[0], edge_index[1]
    N = x.size(0)
    if src.numel() > 0:
        rel = transform2frame(pos[dst], pos[src])                       # :160
        rel = torch.where(torch.isnan(rel), torch.zeros_like(rel), rel)  # :162
        msg_in = torch.cat([x[dst], x[src], sem[dst], sem[src], rel], dim=-1)   # :175
        msg = mlp(sd, p + '.msg.0.edge_mlp', msg_in, 3)
        idx = dst.view(-1, 1).expand(-1, msg.size(1))
        aggr = torch.zeros((N, msg.size(1)), dtype=x.dtype).scatter_reduce(0, idx, msg, 'amax', include_self=False)
        has_in = torch.zeros(N, dtype=torch.bool)
        has_in[dst] = True
        aggr = torch.where(has_in.view(-1, 1), aggr, torch.zeros_like(aggr))
    else:
        aggr = torch.zeros((N, 64), dtype=x.dtype)
    xu = mlp(sd, p + '.msg.0.update_mlp', torch.cat([x, aggr, sem], dim=-1), 2)   # :199,218
    out = mlp(sd, p + '.mlp_out', xu, 3)
    if taps is not None:
        taps.update(x=x, aggr=aggr, xu=xu, out=out)
        if src.numel() > 0:
            # arg-max routing of the max aggregation: per (target, channel) the GLOBAL index of the winning source, and the
            # margin to the runner-up (diagnostics: a kernel may legitimately pick the other one when the margin is at rounding level)
            dense = torch.full((N, N, msg.size(1)), float('-inf'), dtype=x.dtype)
            dense[dst, src] = msg.detach()
            top2 = dense.topk(min(2, N), dim=1)
            taps['arg'] = top2.indices[:, 0]
            taps['arg_margin'] = (top2.values[:, 0] - top2.values[:, 1]) if N > 1 else torch.full_like(top2.values[:, 0], float('inf'))
            taps['msg_dense'] = dense
    return out


def gru3_step(sd, x, h):
    """One time-step of nn.GRU(4,64,3,batch_first) (models/traffic_model.py:152-156, 686-688).
    x (N,4), h (3,N,64) -> top-layer output (N,64), new h. Gate order r,z,n (torch.nn.GRU)."""
    new_h = []
    inp = x
    for l in range(3):
        wi, wh = sd['decoder_memory.weight_ih_l%d' % l], sd['decoder_memory.weight_hh_l%d' % l]
        bi, bh = sd['decoder_memory.bias_ih_l%d' % l], sd['decoder_memory.bias_hh_l%d' % l]
        gi = F.linear(inp, wi, bi)
        gh = F.linear(h[l], wh, bh)
        i_r, i_z, i_n = gi.chunk(3, dim=1)
        h_r, h_z, h_n = gh.chunk(3, dim=1)
        r = torch.sigmoid(i_r + h_r)
        zg = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        hn = (1.0 - zg) * n + zg * h[l]
        new_h.append(hn)
        inp = hn
    return inp, torch.stack(new_h, dim=0)


def bicycle_step(state_un, a, ddh, veh_len):
    """models/traffic_model.py:714-733 (sim_traj, one step) + models/common.py:47-67 (car_dynamics) +
    utils/transforms.py:8-29 (kinematics2angle/vec). state_un (N,6) unnormalised -> (N,6)."""
    x, y, hx, hy, s, hdot = state_un.unbind(dim=1)
    h = torch.atan2(hy, hx)
    newhdot = (hdot + ddh * DT).clamp(-MAXHDOT, MAXHDOT)
    newh = h + DT * s.abs() / veh_len * newhdot
    news = (s + a * DT).clamp(0.0, MAXS)
    newy = y + news * newh.sin() * DT
    newx = x + news * newh.cos() * DT
    return torch.stack([newx, newy, newh.cos(), newh.sin(), news, newhdot], dim=1)


# --------------------------------------------------------------------------------------------
# once-per-batch producers (embed side)
# --------------------------------------------------------------------------------------------

def interaction_net(sd, prefix, feat, pos, sem, edge_index):
    """models/interaction_net.py:52-77 for any of prior_net / posterior_net / decoder_net (single-sample branch)."""
    x = mlp(sd, prefix + '.mlp_in', feat, 3)
    src, dst = edge_index[0], edge_index[1]
    N = x.size(0)
    D = x.size(1)
    if src.numel() > 0:
        rel = transform2frame(pos[dst], pos[src])
        rel = torch.where(torch.isnan(rel), torch.zeros_like(rel), rel)
        msg = mlp(sd, prefix + '.msg.0.edge_mlp', torch.cat([x[dst], x[src], sem[dst], sem[src], rel], dim=-1), 3)
        idx = dst.view(-1, 1).expand(-1, msg.size(1))
        aggr = torch.zeros((N, msg.size(1)), dtype=x.dtype).scatter_reduce(0, idx, msg, 'amax', include_self=False)
        has_in = torch.zeros(N, dtype=torch.bool)
        has_in[dst] = True
        aggr = torch.where(has_in.view(-1, 1), aggr, torch.zeros_like(aggr))
    else:
        aggr = torch.zeros((N, D), dtype=x.dtype)
    xu = mlp(sd, prefix + '.msg.0.update_mlp', torch.cat([x, aggr, sem], dim=-1), 2)
    return mlp(sd, prefix + '.mlp_out', xu, 3)


def _encode_traj(sd, prefix, frame, traj, vis, lw, sem):
    """models/traffic_model.py:453-522 (mlp encoders): trajectory in the frame of the last past step, unobserved frames
    zeroed (whole row), visibility and vehicle attributes appended per frame, semantic class appended once."""
    NA, T, _ = traj.shape
    fr = frame.unsqueeze(1).expand(NA, T, 4).reshape(NA * T, 4)
    local = transform2frame(fr, traj[:, :, :4].reshape(NA * T, 4)).view(NA, T, 4)
    local = torch.cat([local, traj[:, :, 4:]], dim=2)
    local = torch.where((vis == 0.0).unsqueeze(-1), torch.zeros_like(local), local)
    local = torch.cat([local, vis.unsqueeze(-1)], dim=-1)
    enc_in = torch.cat([local, lw.unsqueeze(1).expand(NA, T, 2)], dim=-1)
    enc_in = torch.cat([enc_in.reshape(NA, -1), sem], dim=1)
    return mlp(sd, prefix, enc_in, 4)


def encode_past(sd, past, past_vis, lw, sem):
    """:453-486"""
    return _encode_traj(sd, 'past_encoder', past[:, -1, :4], past, past_vis, lw, sem)


def encode_future(sd, past, future, future_vis, lw, sem):
    """:488-522"""
    return _encode_traj(sd, 'future_encoder', past[:, -1, :4], future, future_vis, lw, sem)


def prior(sd, past, map_feat, past_feat, sem, edge_index):
    """:545-565 -> (mean, var)"""
    out = interaction_net(sd, 'prior_net', torch.cat([past_feat, map_feat, sem], dim=-1), past[:, -1, :4], sem, edge_index)
    return out[:, :32], torch.exp(out[:, 32:])


def posterior(sd, past, map_feat, past_feat, future_feat, sem, edge_index):
    """:524-543 -> (mean, var)"""
    out = interaction_net(sd, 'posterior_net', torch.cat([past_feat, future_feat, map_feat, sem], dim=-1), past[:, -1, :4], sem,
                          edge_index)
    return out[:, :32], torch.exp(out[:, 32:])


def embed(sd, scene, raster, dx, past_vis, future=None, future_vis=None):
    """:372-403"""
    mapixes = scene['map_idx'][scene['batch']]
    map_feat = encode_map(sd, raster, dx, scene['past'][:, -1, :4], mapixes)
    past_feat = encode_past(sd, scene['past'], past_vis, scene['lw'], scene['sem'])
    out = {'map_feat': map_feat, 'past_feat': past_feat,
           'prior_out': prior(sd, scene['past'], map_feat, past_feat, scene['sem'], scene['edge_index'])}
    if future is not None:
        ff = encode_future(sd, scene['past'], future, future_vis, scene['lw'], scene['sem'])
        out['posterior_out'] = posterior(sd, scene['past'], map_feat, past_feat, ff, scene['sem'], scene['edge_index'])
    return out


# --------------------------------------------------------------------------------------------
# map crop + CNN
# --------------------------------------------------------------------------------------------

def map_crop(raster, dx, pose_un, mapixes, L=256, W=256, bounds=BOUNDS):
    """datasets/map_env.py:168-203 + datasets/nuscenes_utils.py:205-264 (gen_car_coords, get_map_obs).
    pose_un (B,4) unnormalised (x,y,hx,hy) -> (B,C,L,W) uint8.  Coordinates are built in the pose dtype,
    divided by the float64 dx (type-promoting to float64), rounded half-to-even; outside -> pixel (0,0);
    note x is divided by dx[:,0] and y by dx[:,1] exactly as the reference does."""
    B = pose_un.size(0)
    C = raster.size(1)
    lwise = torch.linspace(bounds[0], bounds[2], L, dtype=pose_un.dtype).view(1, L, 1)
    wwise = torch.linspace(bounds[1], bounds[3], W, dtype=pose_un.dtype).view(1, 1, W)
    hcos = pose_un[:, 2].view(B, 1, 1)
    hsin = pose_un[:, 3].view(B, 1, 1)
    gx = (lwise * hcos - wwise * hsin) + pose_un[:, 0].view(B, 1, 1)
    gy = (lwise * hsin + wwise * hcos) + pose_un[:, 1].view(B, 1, 1)
    gx = torch.where(torch.isnan(gx), torch.zeros_like(gx), gx)
    gy = torch.where(torch.isnan(gy), torch.zeros_like(gy), gy)
    px = torch.round(gx / dx[mapixes, 0].view(B, 1, 1)).long()
    py = torch.round(gy / dx[mapixes, 1].view(B, 1, 1)).long()
    outside = (py < 0) | (py >= raster.shape[2]) | (px < 0) | (px >= raster.shape[3])
    px = torch.where(outside, torch.zeros_like(px), px)
    py = torch.where(outside, torch.zeros_like(py), py)
    m = mapixes.view(B, 1, 1).expand(B, L, W)
    return torch.stack([raster[m, c, py, px] for c in range(C)], dim=1)


def map_cnn(sd, crop):
    """models/traffic_model.py:69-87, 437-440: 6 x [Conv2d(stride 2, pad 0) -> GroupNorm(1,C) -> ReLU] -> Linear."""
    x = crop
    for li in range(6):
        x = F.conv2d(x, sd['map_conv.%d.weight' % (3 * li)], sd['map_conv.%d.bias' % (3 * li)], stride=2)
        x = F.group_norm(x, 1, sd['map_conv.%d.weight' % (3 * li + 1)], sd['map_conv.%d.bias' % (3 * li + 1)], 1e-5)
        x = F.relu(x)
    return F.linear(x.reshape(x.size(0), -1), sd['map_feature.weight'], sd['map_feature.bias'])


def encode_map(sd, raster, dx, pos_norm, mapixes):
    """models/traffic_model.py:416-451. pos_norm (N,4) normalised -> map_feat (N,64)."""
    pose_un = unnorm_state(pos_norm)
    crop = map_crop(raster, dx, pose_un, mapixes).to(pos_norm.dtype)
    return map_cnn(sd, crop)


# --------------------------------------------------------------------------------------------
# decoder rollout
# --------------------------------------------------------------------------------------------

def decode(sd, z, map_feat, past_feat, past_last, lw, sem, ptr, edge_index, map_idx, raster, dx,
           FT, ext_future=None, taps=None, map_feat_override=None):
    """models/traffic_model.py:589-704 (autoregressive_decoder, output_bicycle=True, single-sample branch).
    All inputs NORMALISED.  Returns future_pred (NA,FT,4) normalised global (x,y,hx,hy).
    `taps` (optional dict) collects per-step intermediates for kernel-level diagnostics.
    `map_feat_override` (optional list of (NA,64) for t=1..FT-1) replaces the map re-encode (diagnostics)."""
    NA = z.size(0)
    batch = torch.repeat_interleave(torch.arange(ptr.numel() - 1), ptr[1:] - ptr[:-1])
    mapixes = map_idx[batch]
    prev_state = past_last                                  # :595
    cur_map_feat, cur_past_feat = map_feat, past_feat
    veh_len = unnorm_att(lw)[:, 0]                          # :601
    ego = ptr[:-1]
    pos = past_last[:, :4]                                  # :604
    mem = past_feat.unsqueeze(0).expand(3, NA, past_feat.size(1)).contiguous()   # :625
    traj = []
    for t in range(FT):
        feat = torch.cat([cur_past_feat, cur_map_feat, sem, z, lw], dim=-1)      # :628-629
        step_taps = {} if taps is not None else None
        out = decoder_net(sd, feat, pos, sem, edge_index, step_taps)             # :632
        a = out[:, 0] * A_STATS[1] + A_STATS[0]                                  # :645
        ddh = out[:, 1] * DDH_STATS[1] + DDH_STATS[0]                            # :646
        cur = norm_state(bicycle_step(unnorm_state(prev_state), a, ddh, veh_len))   # :648-650
        glob = cur[:, :4]
        loc = transform2frame(prev_state[:, :4], glob)                           # :654
        traj.append(glob)                                                        # :665
        if ext_future is not None:                                               # :667-675
            glob = glob.clone()
            glob[ego] = ext_future[:, t]
            loc = loc.clone()
            loc[ego] = transform2frame(prev_state[ego][:, :4], glob[ego])
        if step_taps is not None:
            step_taps.update(feat=feat, cur=cur, loc=loc, pos_in=pos, past_feat=cur_past_feat,
                             map_feat=cur_map_feat, mem=mem)
            taps.setdefault('steps', []).append(step_taps)
        prev_state = cur                                                         # :680
        if t < FT - 1:
            cur_past_feat, mem = gru3_step(sd, loc, mem)                         # :686-688
            if map_feat_override is not None:
                cur_map_feat = map_feat_override[t]
            else:
                cur_map_feat = encode_map(sd, raster, dx, glob.detach(), mapixes)   # :694-695
            pos = glob                                                           # :698
    return torch.stack(traj, dim=1)


# --------------------------------------------------------------------------------------------
# losses (losses/adv_gen_nusc.py)
# --------------------------------------------------------------------------------------------

def interp_traj(traj, scale=3):
    """losses/adv_gen_nusc.py:625-644: F.interpolate(linear, scale_factor=3) then heading renormalised."""
    out = F.interpolate(traj.transpose(1, 2), scale_factor=scale, mode='linear').transpose(1, 2)
    h = out[:, :, 2:4]
    return torch.cat([out[:, :, :2], h / torch.norm(h, dim=-1, keepdim=True)], dim=-1)


def circle_offsets(lw_un, num_circ=5):
    """VehCollLoss.__init__ (:432-437): 5 centres linspace(-l/2+w/2, l/2-w/2) along the long axis, radius w/2."""
    rad = lw_un[:, 1] / 2.0
    cmin = -(lw_un[:, 0] / 2.0) + rad
    cmax = (lw_un[:, 0] / 2.0) - rad
    cx = torch.stack([torch.linspace(cmin[i].item(), cmax[i].item(), num_circ) for i in range(lw_un.size(0))], 0)
    return cx.to(lw_un.dtype), rad


def veh_coll_scene(traj, cx, rad, buffer):
    """VehCollLoss.forward (:464-512) for ONE scene (the block-diagonal block that survives the batch mask
    :447-451).  traj (n,T,4) unnormalised -> pen (T,n,n), mask (T,n,n) [colliding & i!=j]."""
    n, T, _ = traj.shape
    c, s = traj[:, :, 2], traj[:, :, 3]
    wx = c.unsqueeze(-1) * cx.unsqueeze(1) + traj[:, :, 0:1]       # inverse transform2frame, centre (cx,0)
    wy = s.unsqueeze(-1) * cx.unsqueeze(1) + traj[:, :, 1:2]
    cent = torch.stack([wx, wy], dim=-1).transpose(0, 1).reshape(T, n * 5, 2)    # (T,n*5,2)
    # torch.cdist as in the reference (:487); exact (non-matmul) mode, zero-distance-safe backward
    dist = torch.cdist(cent, cent, compute_mode='donot_use_mm_for_euclid_dist')
    dist = dist.view(T, n, 5, n, 5).permute(0, 1, 3, 2, 4).reshape(T, n, n, 25)
    mind = dist.min(dim=-1)[0]
    pdist = rad.view(n, 1) + rad.view(1, n) + buffer
    mask = (mind <= pdist) & (~torch.eye(n, dtype=torch.bool)).view(1, n, n)
    pen = 1.0 - mind / pdist
    return pen, mask


def coll_point(drivable, dx, cars, lw_un, mapixes, L, W):
    """datasets/nuscenes_utils.py:334-390 (get_coll_point) with the batch-global L, W passed in."""
    B = cars.size(0)
    lwise = torch.linspace(-1.0, 1.0, L, dtype=cars.dtype).view(1, L, 1) * lw_un[:, 0].view(B, 1, 1) / 2
    wwise = torch.linspace(-1.0, 1.0, W, dtype=cars.dtype).view(1, 1, W) * lw_un[:, 1].view(B, 1, 1) / 2
    hcos, hsin = cars[:, 2].view(B, 1, 1), cars[:, 3].view(B, 1, 1)
    gx = (lwise * hcos - wwise * hsin) + cars[:, 0].view(B, 1, 1)
    gy = (lwise * hsin + wwise * hcos) + cars[:, 1].view(B, 1, 1)
    px = torch.round(gx / dx[mapixes, 0].view(B, 1, 1)).long()
    py = torch.round(gy / dx[mapixes, 1].view(B, 1, 1)).long()
    outside = (py < 0) | (py >= drivable.shape[1]) | (px < 0) | (px >= drivable.shape[2])
    px = torch.where(outside, torch.zeros_like(px), px)
    py = torch.where(outside, torch.zeros_like(py), py)
    nd = drivable[mapixes.view(B, 1, 1).expand(B, L, W), py, px] == 0
    num = nd.sum(dim=(1, 2))
    # same tensor layout as the reference's reduction, (B,L,W,2) summed over (1,2) (:376-379): the mean of ~600 world coordinates
    # of magnitude 1e2..1e3 m in float32 is only good to ~1e-4 m and the summation order shows in d(penalty)/d(centre)
    xyw = torch.stack([gx, gy], dim=-1)
    pt = (xyw * nd.unsqueeze(-1)).sum(dim=(1, 2)) / num.view(B, 1)
    pt[num == L * W] = float('nan')
    return pt, num


def env_grid_size(dx, lw_un_rows):
    """nuscenes_utils.py:351-354: L,W from the mean of ALL dx entries and the mean lw of the rows passed in."""
    mdx = torch.mean(dx) * 0.5
    mlw = torch.mean(lw_un_rows, dim=0)
    L = torch.round(mlw[0] / mdx).int().item()
    W = torch.round(mlw[1] / mdx).int().item()
    return L, W


def env_coll(traj, lw_un, mapixes, raster, dx):
    """EnvCollLoss.forward (:374-403). traj (N,T,4) unnormalised -> penalties (n_valid,) or None if none."""
    N, T, _ = traj.shape
    flat = traj.reshape(N * T, 4)
    att = lw_un.view(N, 1, 2).expand(N, T, 2).reshape(N * T, 2)
    mix = mapixes.view(N, 1).expand(N, T).reshape(N * T)
    L, W = env_grid_size(dx, att)
    pt, _ = coll_point(raster[:, 0], dx, flat.detach(), att, mix, L, W)
    valid = ~torch.isnan(pt.sum(dim=1))
    if valid.sum() == 0:
        return None
    pen_d = torch.sqrt(lw_un[:, 0] ** 2 / 4.0 + lw_un[:, 1] ** 2 / 4.0).view(N, 1).expand(N, T).reshape(N * T)
    dist = torch.norm(flat[:, :2][valid] - pt[valid], dim=1)
    return 1.0 - dist / pen_d[valid]


def motion_prior_nll(z, mu, var):
    """MotionPriorLoss (:343-364) = -log_normal (losses/common.py:26-41)."""
    lp = -torch.log(torch.sqrt(var)) - math.log(math.sqrt(2 * math.pi)) - ((z - mu) ** 2 / (2 * var))
    return -lp.sum(dim=-1)


def avoid_coll_loss(future_un, z, prior, init_z, weights, lw_un, mapixes, ptr, raster, dx,
                    veh_coll_buffer=0.0, single_veh_idx=None):
    """AvoidCollLoss.forward (:303-341) for one reference batch (= one loss-normalisation group).
    ptr=None reproduces the refine driver (refine_traffic_optim.py:176-181 passes no ptr, so VehCollLoss
    :443-451 treats the WHOLE batch as one block: agents of different scenes can collide).
    With single_veh_idx (sol_optim.py:57-63) z/prior/init_z are (B,D) for that agent of each scene."""
    if ptr is None:
        ptr = torch.tensor([0, future_un.size(0)])
    out = {}
    loss = torch.zeros((), dtype=future_un.dtype)
    fi = interp_traj(future_un)
    cx, rad = circle_offsets(lw_un)
    S = ptr.numel() - 1
    if weights['coll_veh'] > 0.0:
        pens = []
        for s in range(S):
            a, b = int(ptr[s]), int(ptr[s + 1])
            pen, mask = veh_coll_scene(fi[a:b], cx[a:b], rad[a:b], veh_coll_buffer)
            if single_veh_idx is not None:
                sm = torch.zeros(b - a, dtype=torch.bool)
                sm[single_veh_idx] = True
                mask = mask & (sm.view(1, -1, 1) | sm.view(1, 1, -1))
            pens.append(pen[mask])
        pens = torch.cat(pens)
        if pens.numel() == 0:
            pens = torch.zeros(1, dtype=future_un.dtype)     # :502-503
        loss = loss + weights['coll_veh'] * pens.mean()
        out['coll_veh_loss'] = pens
    if weights['coll_env'] > 0.0:
        if single_veh_idx is not None:
            sel = ptr[:-1] + single_veh_idx
            pen = env_coll(fi[sel], lw_un[sel], mapixes[sel], raster, dx)
        else:
            pen = env_coll(fi, lw_un, mapixes, raster, dx)
        if pen is None:
            pen = torch.zeros(1, dtype=future_un.dtype)      # :393-394
        loss = loss + weights['coll_env'] * pen.mean()
        out['coll_env_loss'] = pen
    if weights['motion_prior'] > 0.0:
        nll = motion_prior_nll(z, prior[0], prior[1])
        loss = loss + weights['motion_prior'] * nll.mean()
        out['motion_prior_loss'] = nll
    if weights.get('init_z', 0.0) > 0.0:
        il = torch.sum((init_z - z) ** 2, dim=1)
        loss = loss + weights['init_z'] * il.mean()
        out['init_loss'] = il
    out['loss'] = loss
    return out


def tgt_matching_loss(future_un, tgt_un, weights):
    """TgtMatchingLoss.forward (:27-51) INCLUDING the reference's bug at :46 (the prior NLL is computed but
    the matching mean is what gets added a second time)."""
    loss = torch.zeros((), dtype=future_un.dtype)
    out = {}
    tl = torch.sum((future_un - tgt_un) ** 2, dim=-1)
    if weights['match_ext'] > 0.0:
        loss = loss + weights['match_ext'] * tl.mean()
        out['match_ext_loss'] = tl
    if weights['motion_prior_ext'] > 0.0:
        loss = loss + weights['motion_prior_ext'] * tl.mean()
    out['loss'] = loss
    return out


def adv_gen_loss(future_un, tgt_un, z_other, prior_other, init_z_other, weights, lw_un, mapixes, ptr,
                 raster, dx, veh_coll_buffer=0.0, crash_min_t=0, crash_min_infront=None, attack_agt_idx=None):
    """AdvGenLoss.forward (:93-262) + check_behind (:646-673), one reference batch.
    future_un (NA,T,4), tgt_un (B,T,4); z_other/prior/init for the NA-B non-ego agents in graph order."""
    NA, T, _ = future_un.shape
    B = ptr.numel() - 1
    dt = future_un.dtype
    ego_mask = torch.zeros(NA, dtype=torch.bool)
    ego_mask[ptr[:-1]] = True
    sizes = ptr[1:] - ptr[:-1]
    nonego_ptr = ptr - torch.arange(B + 1)
    scene_of_other = torch.repeat_interleave(torch.arange(B), sizes - 1)
    atk = future_un[~ego_mask][:, crash_min_t:, :]
    tgt_e = tgt_un[:, crash_min_t:, :4][scene_of_other]
    dist = torch.norm(atk[:, :, :2] - tgt_e[:, :, :2], dim=-1)
    dist_in = dist
    NT = T - crash_min_t
    if crash_min_infront is not None:
        v = atk[:, :, :2].detach() - tgt_e[:, :, :2].detach()
        v = v / torch.norm(v, dim=-1, keepdim=True)
        cossim = torch.sum(v * tgt_e[:, :, 2:4].detach(), dim=-1)
        behind = cossim < crash_min_infront
        behind_traj = (behind.sum(dim=1, keepdim=True) == NT).expand_as(behind)
        if behind_traj.all():
            behind_traj = torch.zeros_like(behind_traj)
        dist_in = torch.where(behind_traj, torch.full_like(dist_in, float('inf')), dist_in)
    if attack_agt_idx is not None:
        am = torch.zeros(NA, dtype=torch.bool)
        am[attack_agt_idx] = True
        am = am[~ego_mask].unsqueeze(1).expand_as(dist_in)
        dist_in = torch.where(~am, torch.full_like(dist_in, float('inf')), dist_in)
    soft, crash = [], []
    min_agt, min_t = [], []
    for b in range(B):
        a0, a1 = int(nonego_ptr[b]), int(nonego_ptr[b + 1])
        sm = F.softmin(dist_in[a0:a1].reshape(-1), dim=0)
        if torch.isnan(sm[0]):
            sm = torch.zeros_like(sm)
        am_ = int(torch.max(sm, dim=0)[1])
        min_agt.append(am_ // NT + 1)
        min_t.append(am_ % NT + crash_min_t)
        soft.append(sm)
        crash.append(torch.sum(sm * dist[a0:a1].reshape(-1) ** 2))
    soft = torch.cat(soft)
    crash = torch.stack(crash)
    rew = 1.0 - soft.detach().reshape(NA - B, NT).sum(dim=1)             # :151-152
    out = {}
    loss = torch.zeros((), dtype=dt)
    if weights.get('init_z', 0.0) > 0.0:
        il = torch.sum((init_z_other - z_other) ** 2, dim=1)
        coeff = rew * weights['init_z'] + (1.0 - rew) * weights['init_z_atk']
        il = torch.sum(il * coeff)                                        # :222 (a scalar; .mean() is a no-op)
        loss = loss + il
        out['init_loss'] = il
    if weights.get('motion_prior', 0.0) > 0.0:
        nll = motion_prior_nll(z_other, prior_other[0], prior_other[1])
        coeff = rew * weights['motion_prior'] + (1.0 - rew) * weights['motion_prior_atk']
        nll = nll * coeff
        loss = loss + nll.mean()
        out['motion_prior_loss'] = nll
    fi = interp_traj(future_un)
    cx, rad = circle_offsets(lw_un)
    ego_w = torch.ones(NA, dtype=dt)
    ego_w[~ego_mask] = rew
    veh_pens, plan_pens = [], []
    for s in range(B):
        a, b = int(ptr[s]), int(ptr[s + 1])
        pen, mask = veh_coll_scene(fi[a:b], cx[a:b], rad[a:b], veh_coll_buffer)
        n = b - a
        is_ego = torch.zeros(n, dtype=torch.bool)
        is_ego[0] = True
        ego_pair = is_ego.view(1, n, 1) | is_ego.view(1, 1, n)
        veh_pens.append(pen[mask & ~ego_pair])                            # :179-187
        wmat = torch.ones(n, n, dtype=dt)
        wmat[0, :] = ego_w[a:b]
        wmat[:, 0] = ego_w[a:b]
        plan_pens.append((pen * wmat.view(1, n, n))[mask & ego_pair])     # :192-208
    if weights.get('coll_veh', 0.0) > 0.0:
        v = torch.cat(veh_pens)
        if v.numel() == 0:
            v = torch.zeros(1, dtype=dt)
        loss = loss + weights['coll_veh'] * v.mean()
        out['coll_veh_loss'] = v
    if weights.get('coll_veh_plan', 0.0) > 0.0:
        v = torch.cat(plan_pens)
        if v.numel() == 0:
            v = torch.zeros(1, dtype=dt)
        loss = loss + weights['coll_veh_plan'] * v.mean()
        out['coll_veh_plan_loss'] = v
    if weights.get('coll_env', 0.0) > 0.0:
        pen = env_coll(fi[~ego_mask], lw_un[~ego_mask], mapixes[~ego_mask], raster, dx)
        if pen is None:
            pen = torch.zeros(1, dtype=dt)
        loss = loss + weights['coll_env'] * pen.mean()
        out['coll_env_loss'] = pen
    loss = loss + weights['adv_crash'] * crash.mean()
    out['adv_crash_loss'] = crash
    out['loss'] = loss
    out['min_agt'] = min_agt
    out['min_t'] = min_t
    return out


# --------------------------------------------------------------------------------------------
# Adam on z (torch.optim.Adam defaults used by the drivers: betas (0.9,0.999), eps 1e-8, no weight decay)
# --------------------------------------------------------------------------------------------

def adam_step(z, g, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8):
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    z = z - (lr / bc1) * (m / denom)
    return z, m, v


def refine_loop(sd, scene, raster, dx, weights, iters, lr, FT, veh_coll_buffer=0.2, groups=None, record=None):
    """refine_traffic_optim.py:146-226 latent Adam loop (decode -> AvoidCollLoss -> backward -> Adam.step).
    `groups` = list of scene-index lists forming independent reference batches (loss-normalisation groups)."""
    z = scene['z'].clone().requires_grad_(True)
    init_z = scene['z'].clone()
    opt = torch.optim.Adam([z], lr=lr)
    ptr = scene['ptr']
    S = ptr.numel() - 1
    if groups is None:
        groups = [list(range(S))]
    batch = scene['batch']
    lw_un = unnorm_att(scene['lw'])
    mapixes = scene['map_idx'][batch]
    for it in range(iters):
        opt.zero_grad()
        fut = decode(sd, z, scene['map_feat'], scene['past_feat'], scene['past'][:, -1, :], scene['lw'],
                     scene['sem'], ptr, scene['edge_index'], scene['map_idx'], raster, dx, FT)
        fut_un = unnorm_state(fut)
        total = 0.0
        for g in groups:
            idx = torch.cat([torch.arange(int(ptr[s]), int(ptr[s + 1])) for s in g])
            ld = avoid_coll_loss(fut_un[idx], z[idx], (scene['prior_mu'][idx], scene['prior_var'][idx]),
                                 init_z[idx], weights, lw_un[idx], mapixes[idx], None, raster, dx,
                                 veh_coll_buffer=veh_coll_buffer)
            total = total + ld['loss']
        total.backward()
        if record is not None:
            record.append({'loss': float(total.detach()), 'grad': z.grad.detach().clone(), 'traj': fut.detach().clone()})
        opt.step()
    return z.detach()


# --------------------------------------------------------------------------------------------
# adversarial / solution loops, literal restatements (two decodes per iteration as the reference)
# --------------------------------------------------------------------------------------------
def _collate(ptr, tgt_z, other_z):
    """utils/adv_gen_optim.py:19-36"""
    out, prev = [], 0
    for b in range(ptr.numel() - 1):
        n = int(ptr[b + 1] - ptr[b]) - 1
        out += [tgt_z[b:b + 1], other_z[prev:prev + n]]
        prev += n
    return torch.cat(out, 0)


def adv_loop(sd, scene, raster, dx, weights, iters, lr, FT, planner_fut_n, crash_min_t=0, crash_min_infront=None, veh_coll_buffer=0.1,
             record=None):
    """utils/adv_gen_optim.py:39-175, planner_name == 'ego' (open-loop replay, planner_inject_traj=True)."""
    ptr = scene['ptr']
    NA = int(ptr[-1])
    ego = torch.zeros(NA, dtype=torch.bool)
    ego[ptr[:-1]] = True
    tgt_z = scene['z'][ego].clone().requires_grad_(True)
    other_z = scene['z'][~ego].clone().requires_grad_(True)
    init_o = scene['z'][~ego].clone()
    opt = torch.optim.Adam([tgt_z, other_z], lr=lr)
    lw_un = unnorm_att(scene['lw'])
    mapixes = scene['map_idx'][scene['batch']]
    tprior = (scene['prior_mu'][ego], scene['prior_var'][ego])
    oprior = (scene['prior_mu'][~ego], scene['prior_var'][~ego])
    pf_un = unnorm_state(planner_fut_n)

    def dec(z):
        return decode(sd, z, scene['map_feat'], scene['past_feat'], scene['past'][:, -1, :], scene['lw'], scene['sem'], ptr,
                      scene['edge_index'], scene['map_idx'], raster, dx, FT, ext_future=planner_fut_n)
    for it in range(iters):
        opt.zero_grad()
        f_t = dec(_collate(ptr, tgt_z, other_z.detach()))
        f_o = dec(_collate(ptr, tgt_z.detach(), other_z))
        lt = tgt_matching_loss(unnorm_state(f_t)[ego], pf_un, weights)
        la = adv_gen_loss(unnorm_state(f_o), pf_un, other_z, oprior, init_o, weights, lw_un, mapixes, ptr, raster, dx,
                          veh_coll_buffer=veh_coll_buffer, crash_min_t=crash_min_t, crash_min_infront=crash_min_infront)
        loss = lt['loss'] + la['loss']
        loss.backward()
        if record is not None:
            record.append({'loss': float(loss.detach()), 'g_tgt': tgt_z.grad.clone(), 'g_other': other_z.grad.clone()})
        opt.step()
    return _collate(ptr, tgt_z.detach(), other_z.detach())


def sol_loop(sd, scene, raster, dx, weights, iters, lr, future_len, FT, other_match_un, record=None):
    """utils/sol_optim.py:19-123 (weights already stripped of the 'sol_' prefix)."""
    ptr = scene['ptr']
    NA = int(ptr[-1])
    ego = torch.zeros(NA, dtype=torch.bool)
    ego[ptr[:-1]] = True
    tgt_z = scene['prior_mu'][ego].clone().requires_grad_(True)
    other_z = scene['z'][~ego].clone().requires_grad_(True)
    init_t = tgt_z.detach().clone()
    opt = torch.optim.Adam([tgt_z, other_z], lr=lr)
    lw_un = unnorm_att(scene['lw'])
    mapixes = scene['map_idx'][scene['batch']]
    tprior = (scene['prior_mu'][ego], scene['prior_var'][ego])

    def dec(z, n):
        return decode(sd, z, scene['map_feat'], scene['past_feat'], scene['past'][:, -1, :], scene['lw'], scene['sem'], ptr,
                      scene['edge_index'], scene['map_idx'], raster, dx, n)
    for it in range(iters):
        opt.zero_grad()
        f_t = dec(_collate(ptr, tgt_z, other_z.detach()), future_len)
        f_o = dec(_collate(ptr, tgt_z.detach(), other_z), FT)
        lt = avoid_coll_loss(unnorm_state(f_t), tgt_z, tprior, init_t, weights, lw_un, mapixes, ptr, raster, dx, veh_coll_buffer=0.5,
                             single_veh_idx=0)
        lo = tgt_matching_loss(unnorm_state(f_o)[~ego], other_match_un, weights)
        loss = lt['loss'] + lo['loss']
        loss.backward()
        if record is not None:
            record.append({'loss': float(loss.detach()), 'g_tgt': tgt_z.grad.clone(), 'g_other': other_z.grad.clone()})
        opt.step()
    return _collate(ptr, tgt_z.detach(), other_z.detach())


def init_loop(sd, scene, raster, dx, weights, iters, lr, FT, init_traj_n, traj_vis, record=None):
    """utils/init_optim.py:11-68: Adam([z]) on TgtMatchingLoss (init_* weights) between the decoded future and the observed
    one at the visible (agent, step) entries."""
    z = scene['z'].clone().requires_grad_(True)
    opt = torch.optim.Adam([z], lr=lr)
    vis = traj_vis == 1.0
    tgt_un = unnorm_state(init_traj_n)[vis]
    w = {k[5:]: v for k, v in weights.items() if k[:5] == 'init_'}
    for it in range(iters):
        opt.zero_grad()
        fut = decode(sd, z, scene['map_feat'], scene['past_feat'], scene['past'][:, -1, :], scene['lw'], scene['sem'], scene['ptr'],
                     scene['edge_index'], scene['map_idx'], raster, dx, FT)
        ld = tgt_matching_loss(unnorm_state(fut)[vis], tgt_un, w)
        ld['loss'].backward()
        if record is not None:
            record.append({'loss': float(ld['loss'].detach()), 'grad': z.grad.clone()})
        opt.step()
    return z.detach()
