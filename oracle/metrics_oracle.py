"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the success / plausibility checks around the latent loop:
  check_on_layer, check_line_layer     /root/reference/src/datasets/nuscenes_utils.py:266-333
  compute_coll_rate_env                /root/reference/src/losses/traffic_model.py:366-419
  determine_feasibility_nusc           /root/reference/src/utils/scenario_gen.py:30-107
  check_single_veh_coll / check_pairwise_veh_coll   /root/reference/src/losses/adv_gen_nusc.py:517-623

Pinning: the raster functions and determine_feasibility_nusc are checked against outputs of the unmodified reference
(tests/golden/metrics.npz, written by oracle/gen_golden_metrics.py).  The two vehicle-collision checks call
shapely (`Polygon.intersection/union`), a third-party dependency that is absent from /root/reference and from this image
(requirements.txt pins shapely==1.7.1): PARITY UNPINNED for the polygon arithmetic itself -- rect_iou below restates the
published definition (area of intersection / area of union of two rotated rectangles) with convex clipping in float64 and is
anchored on closed-form cases in tests/test_oracle_golden.py; the loop structure around it (first colliding step, NaN
skip, one-sided pair flagging) follows the reference lines cited per function.
"""
import numpy as np
import torch

VEH_COLL_THRESH = 0.02
ENV_COLL_THRESH = 0.05


def gen_car_coords_lw(xy, hs, L, W, ls, ws):
    """nuscenes_utils.py:205-232, ls/ws branch, C = 1."""
    B = hs.size(0)
    lwise = torch.linspace(-1.0, 1.0, L).view(1, L, 1).expand(B, L, W) * ls.view(B, 1, 1) / 2
    wwise = torch.linspace(-1.0, 1.0, W).view(1, 1, W).expand(B, L, W) * ws.view(B, 1, 1) / 2
    hcos, hsin = hs[:, 0].view(B, 1, 1), hs[:, 1].view(B, 1, 1)
    return torch.stack((lwise * hcos - wwise * hsin, lwise * hsin + wwise * hcos), 3) + xy.view(B, 1, 1, 2)


def check_on_layer(layer, dx, cars, lw, mapixes):
    """nuscenes_utils.py:266-298.  layer (M,H,W) uint8."""
    mdx = torch.mean(dx)
    mlw = torch.mean(lw, dim=0)
    L = torch.round(mlw[0] / mdx).int().item()
    W = torch.round(mlw[1] / mdx).int().item()
    B = cars.size(0)
    xys = gen_car_coords_lw(cars[:, :2], cars[:, 2:], L, W, lw[:, 0], lw[:, 1])
    xys = torch.round(xys / dx[mapixes].view(B, 1, 1, 2)).long()
    outside = (xys[..., 1] < 0) | (xys[..., 1] >= layer.shape[1]) | (xys[..., 0] < 0) | (xys[..., 0] >= layer.shape[2])
    xys[outside] = 0
    pix = layer[mapixes.view(B, 1, 1).expand(B, L, W), xys[..., 1], xys[..., 0]]
    return torch.sum(pix.float(), dim=[1, 2]) / (L * W)


def compute_coll_rate_env(lw_norm, mapixes_agent, pred_un, lw_un, raster, dx):
    """losses/traffic_model.py:366-419 after the normaliser calls: pred_un (NA,NS,FT,4), lw_un (NA,2)."""
    NA, NS, FT, _ = pred_un.shape
    flat = pred_un.reshape(NA * NS * FT, 4)
    att = lw_un.view(NA, 1, 1, 2).expand(NA, NS, FT, 2).reshape(NA * NS * FT, 2)
    mix = mapixes_agent.view(NA, 1, 1).expand(NA, NS, FT).reshape(NA * NS * FT)
    valid = ~torch.isnan(flat.sum(-1))
    frac = torch.ones(NA * NS * FT)
    frac[valid] = check_on_layer(raster[:, 0], dx, flat[valid], att[valid], mix[valid])
    coll = (frac.view(NA, NS, FT) < (1.0 - ENV_COLL_THRESH)).sum(dim=2) >= 1
    return coll


def check_line_layer(layer, dx, start, end, mapixes):
    """nuscenes_utils.py:300-333."""
    B = start.size(0)
    line_len = torch.norm(start - end, dim=-1)
    mdx = torch.mean(dx)
    L = torch.max(torch.round(line_len / mdx).int()).item()
    w = torch.linspace(0.0, 1.0, L).view(1, L, 1).expand(B, L, 2)
    pts = start.view(B, 1, 2) * (1.0 - w) + end.view(B, 1, 2) * w
    xys = torch.round(pts / dx[mapixes].view(B, 1, 2)).long()
    pix = layer[mapixes.view(B, 1).expand(B, L), xys[:, :, 1], xys[:, :, 0]]
    return torch.sum(pix == 0, dim=-1) > 0


def determine_feasibility(samples_un, thresh, ftime, fvel, infront_min, sep, layer, dx, map_idx):
    """utils/scenario_gen.py:63-107 on UNNORMALISED samples (NA,NS,FT,4)."""
    ego, ag = samples_un[0:1], samples_un[1:]
    NA, NS, FT, _ = ag.shape
    d = torch.norm(ego[..., :2] - ag[..., :2], dim=-1)[:, :, ftime:]
    if infront_min is not None:
        e2a = ag[:, :, ftime:, :2] - ego[:, :, ftime:, :2]
        e2a = e2a / torch.norm(e2a, dim=-1, keepdim=True)
        infront = torch.sum(e2a * ego[:, :, ftime:, 2:4], dim=-1) >= infront_min
        d[~infront] = float('inf')
    msd, msi = torch.min(d, dim=1)
    fdist, fstep = torch.min(msd, dim=1)
    fstep = fstep + ftime
    feas = (d < thresh).sum(dim=[1, 2]) > 0
    if sep:
        ar = torch.arange(NA)
        m = msi[ar, fstep - ftime]
        a_xy = ag[ar, m][ar, fstep][:, :2]
        e_xy = ego.expand(NA, NS, FT, 4)[ar, m][ar, fstep][:, :2]
        feas = feas & ~check_line_layer(layer, dx, a_xy, e_xy, map_idx.expand(NA))
    vel = torch.norm(ag[:, :, 1:, :2] - ag[:, :, :-1, :2], dim=-1)
    feas = feas & (vel.max(dim=1)[0].max(dim=1)[0] > fvel)
    return feas, fstep, fdist


# ----------------------------------------------------------------------------------------------------------
# rotated rectangles
# ----------------------------------------------------------------------------------------------------------
def get_corners(box, lw):
    """nuscenes_utils.py:416-428 with numpy float32 inputs (as check_*_veh_coll passes them): float32 arithmetic."""
    l, w = lw
    simple_box = np.array([[-l / 2., -w / 2.], [l / 2., -w / 2.], [l / 2., w / 2.], [-l / 2., w / 2.]])
    h = np.arctan2(box[3], box[2])
    rot = np.array([[np.cos(h), np.sin(h)], [-np.sin(h), np.cos(h)]])
    return np.dot(simple_box, rot) + box[:2]


def _area(p):
    x, y = p[:, 0], p[:, 1]
    return 0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def rect_iou(ca, cb):
    """intersection / union area of two convex CCW quadrilaterals (float64 Sutherland-Hodgman clipping)."""
    ca, cb = np.asarray(ca, dtype=np.float64), np.asarray(cb, dtype=np.float64)
    cur = [tuple(p) for p in ca]
    for e in range(4):
        a, b = cb[e], cb[(e + 1) % 4]
        ex, ey = b[0] - a[0], b[1] - a[1]
        nxt = []
        n = len(cur)
        for k in range(n):
            p, q = cur[k], cur[(k + 1) % n]
            dp = ex * (p[1] - a[1]) - ey * (p[0] - a[0])
            dq = ex * (q[1] - a[1]) - ey * (q[0] - a[0])
            if dp >= 0.0:
                nxt.append(p)
            if (dp > 0.0 and dq < 0.0) or (dp < 0.0 and dq > 0.0):
                t = dp / (dp - dq)
                nxt.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
        cur = nxt
        if not cur:
            break
    inter = abs(_area(np.array(cur))) if len(cur) >= 3 else 0.0
    union = abs(_area(ca)) + abs(_area(cb)) - inter
    return inter / union if union > 0.0 else 0.0


def check_single_veh_coll(traj_tgt, lw_tgt, traj_others, lw_others):
    """adv_gen_nusc.py:517-565 (numpy float32 inputs)."""
    NA, FT, _ = traj_others.shape
    veh_coll = np.zeros(NA, dtype=bool)
    coll_time = np.ones(NA, dtype=int) * FT
    for aj in range(NA):
        for t in range(FT):
            if np.sum(np.isnan(traj_others[aj, t])) > 0:
                continue
            if np.sum(np.isnan(traj_tgt[t])) > 0:       # shapely would produce an invalid polygon / NaN IoU: never > thresh
                continue
            iou = rect_iou(get_corners(traj_tgt[t], lw_tgt), get_corners(traj_others[aj, t], lw_others[aj]))
            if iou > VEH_COLL_THRESH:
                veh_coll[aj] = True
                coll_time[aj] = t
                break
    return veh_coll, coll_time


def check_pairwise_veh_coll(traj, lw):
    """adv_gen_nusc.py:567-623."""
    NA, FT, _ = traj.shape
    veh_coll = np.zeros(NA, dtype=bool)
    count = 0
    for ai in range(NA):
        for aj in range(ai + 1, NA):
            if veh_coll[ai]:
                break
            for t in range(FT):
                if np.isnan(traj[ai, t]).any() or np.isnan(traj[aj, t]).any():
                    continue
                if rect_iou(get_corners(traj[ai, t], lw[ai]), get_corners(traj[aj, t], lw[aj])) > VEH_COLL_THRESH:
                    count += 1
                    veh_coll[ai] = True
                    break
    return {'num_coll_veh': float(count), 'num_traj_veh': float(NA), 'did_collide': veh_coll}
