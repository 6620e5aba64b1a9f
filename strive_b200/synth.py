"""Deterministic synthetic inputs shaped like the reference's nuScenes pipeline.

There is no dataset and no checkpoint in this environment, so tests, bench.py and the
golden-vector generator all draw from here:

  * make_raster  -- stands in for NuScenesMapEnv.nusc_raster / nusc_dx
                    (reference src/datasets/map_env.py:79-166: (M,4,H,W) uint8 + (M,2) float64 m/px).
  * make_scenes  -- stands in for NuScenesDataset's collated scene graph
                    (reference src/datasets/nuscenes_dataset.py:609-687: node 0 of each scene = ego,
                    fully connected directed edges inside a scene, `past` (NA,PT,6) normalised).
  * make_weights -- seeded state_dict for the decode-path modules (decoder_net, decoder_memory,
                    map_conv, map_feature) with the reference's key names/shapes.

Everything is numpy RandomState / torch.Generator seeded so the GPU box regenerates bit-identical
inputs; fixtures carry checksums to prove it.
"""
import math

import numpy as np
import torch

# reference src/datasets/utils.py:121-140 (car/truck stats), nuscenes_dataset.py:213-222
STATE_MEAN = (0.0, 0.0, 0.0, 0.0, 1.802009, -0.000037)
STATE_STD = (15.0, 15.0, 1.0, 1.0, 3.507907, 0.055684)
ATT_MEAN = (4.844294, 2.021752)
ATT_STD = (1.084860, 0.299647)
BIKE_PARAMS = {'maxs': 50.0, 'maxhdot': 2.0 * math.pi, 'dt': 0.5,
               'a_stats': (0.409074, 1.045530), 'ddh_stats': (0.000046, 0.075032)}
CROP_BOUNDS = (-17.0, -38.5, 60.0, 38.5)

ROAD_PITCH_M = 60.0
ROAD_HALF_W_M = 7.0


def make_raster(seed=0, M=1, H=4096, W=4096, dx=None):
    """(M,4,H,W) uint8 raster + (M,2) float64 metres/pixel.

    Layer 0 = drivable: union of axis-aligned road strips 14 m wide on a 60 m grid.
    Layers 1-3 = blocky 30 %-density noise (stand-ins for carpark / dividers).
    """
    rng = np.random.RandomState(seed)
    if dx is None:
        dx = np.tile(np.array([[0.25, 0.25]], dtype=np.float64), (M, 1))
        for m in range(1, M):
            # distinct, non-square resolutions exercise the reference's x/dx[:,0], y/dx[:,1] convention
            dx[m] = [0.25 + 0.0007 * m, 0.25 - 0.0004 * m]
    dx = np.asarray(dx, dtype=np.float64).reshape(M, 2)
    raster = np.zeros((M, 4, H, W), dtype=np.uint8)
    for m in range(M):
        xs = np.arange(W) * dx[m, 0]
        ys = np.arange(H) * dx[m, 1]
        on_v = np.abs(((xs - ROAD_PITCH_M / 2) % ROAD_PITCH_M) - 0.0)
        on_v = np.minimum(on_v, ROAD_PITCH_M - on_v) <= ROAD_HALF_W_M
        on_h = np.abs(((ys - ROAD_PITCH_M / 2) % ROAD_PITCH_M) - 0.0)
        on_h = np.minimum(on_h, ROAD_PITCH_M - on_h) <= ROAD_HALF_W_M
        raster[m, 0] = (on_h[:, None] | on_v[None, :]).astype(np.uint8)
        cell = 8
        for c in range(1, 4):
            noise = (rng.random_sample((H // cell + 1, W // cell + 1)) < 0.3).astype(np.uint8)
            raster[m, c] = np.kron(noise, np.ones((cell, cell), dtype=np.uint8))[:H, :W]
    return torch.from_numpy(raster), torch.from_numpy(dx)


def clique_edges(ptr):
    """Fully connected directed edges inside each scene (nuscenes_dataset.py:678-687)."""
    src, dst = [], []
    for s in range(len(ptr) - 1):
        a, b = int(ptr[s]), int(ptr[s + 1])
        n = b - a
        if n < 2:
            continue
        ii, jj = np.meshgrid(np.arange(a, b), np.arange(a, b), indexing='ij')
        mask = ii != jj
        src.append(jj[mask])
        dst.append(ii[mask])
    if not src:
        return torch.zeros((2, 0), dtype=torch.long)
    return torch.from_numpy(np.stack([np.concatenate(src), np.concatenate(dst)], 0)).long()


def make_scenes(seed, sizes, map_extent_m=(200.0, 800.0), M=1, PT=4, FT=20,
                collide_frac=0.25, offroad_frac=0.25, dtype=torch.float32):
    """Synthetic batched scene graph + latent-loop inputs.

    Returns a dict of CPU tensors:
      past (NA,PT,6) normalised, lw (NA,2) normalised, sem (NA,2) one-hot, ptr (S+1) int64,
      batch (NA) int64, map_idx (S) int64, edge_index (2,E), z / prior_mu / prior_var (NA,32),
      map_feat / past_feat (NA,64), ext_future (S,FT,4) normalised constant-velocity ego futures.
    """
    rng = np.random.RandomState(seed)
    S = len(sizes)
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    NA = int(ptr[-1])
    lo, hi = map_extent_m
    pos = np.zeros((NA, 2))
    head = np.zeros(NA)
    speed = rng.uniform(2.0, 10.0, NA)
    l = np.clip(rng.normal(ATT_MEAN[0], ATT_STD[0], NA), 3.2, 9.0)
    w = np.clip(rng.normal(ATT_MEAN[1], ATT_STD[1], NA), 1.5, 2.9)
    w = np.minimum(w, l - 0.6)
    truck = rng.random_sample(NA) < 0.2
    map_idx = rng.randint(0, M, S).astype(np.int64)
    k_lo = int(math.ceil((lo - ROAD_PITCH_M / 2) / ROAD_PITCH_M))
    k_hi = int(math.floor((hi - ROAD_PITCH_M / 2) / ROAD_PITCH_M))
    for s in range(S):
        gx = ROAD_PITCH_M * rng.randint(k_lo, k_hi + 1) + ROAD_PITCH_M / 2
        gy = ROAD_PITCH_M * rng.randint(k_lo, k_hi + 1) + ROAD_PITCH_M / 2
        for a in range(int(ptr[s]), int(ptr[s + 1])):
            horiz = rng.random_sample() < 0.5
            along = rng.uniform(-70.0, 70.0)
            side = 1.0 if rng.random_sample() < 0.5 else -1.0
            lat = side * 3.5 + rng.normal(0.0, 0.3)
            other = ROAD_PITCH_M * rng.randint(-1, 2) if rng.random_sample() < 0.3 else 0.0
            if horiz:
                pos[a] = [gx + along, gy + other + lat]
                head[a] = (0.0 if side < 0 else math.pi) + rng.normal(0.0, 0.1)
            else:
                pos[a] = [gx + other + lat, gy + along]
                head[a] = (math.pi / 2 if side > 0 else -math.pi / 2) + rng.normal(0.0, 0.1)
        n = int(ptr[s + 1] - ptr[s])
        a0 = int(ptr[s])
        pair = ()
        if n >= 2 and rng.random_sample() < collide_frac:
            # an overlapping pair so the vehicle-collision term is active
            i = rng.randint(0, n)
            j = a0 + (i + 1 + rng.randint(0, n - 1)) % n
            i = a0 + i
            pos[j] = pos[i] + rng.normal(0.0, 0.8, 2)
            head[j] = head[i] + rng.normal(0.0, 0.2)
            speed[j] = speed[i]
            pair = (i, j)
        if rng.random_sample() < offroad_frac:
            # one agent straddling the road edge so the drivable-area term is active
            i = a0 + rng.randint(0, n)
            if n > 2:
                while i in pair:
                    i = a0 + rng.randint(0, n)
            gy2 = ROAD_PITCH_M * round((pos[i, 1] - ROAD_PITCH_M / 2) / ROAD_PITCH_M) + ROAD_PITCH_M / 2
            pos[i, 1] = gy2 + ROAD_HALF_W_M + rng.uniform(-0.8, 0.8)
            head[i] = rng.normal(0.0, 0.1)
            pos[i, 0] = ROAD_PITCH_M * round(pos[i, 0] / ROAD_PITCH_M) + rng.uniform(-8, 8)
    # constant-velocity past, last past step at `pos`
    past = np.zeros((NA, PT, 6))
    for t in range(PT):
        back = (PT - 1 - t) * BIKE_PARAMS['dt']
        past[:, t, 0] = pos[:, 0] - back * speed * np.cos(head)
        past[:, t, 1] = pos[:, 1] - back * speed * np.sin(head)
        past[:, t, 2] = np.cos(head)
        past[:, t, 3] = np.sin(head)
        past[:, t, 4] = speed
        past[:, t, 5] = 0.0
    fut = np.zeros((S, FT, 4))
    for t in range(FT):
        fwd = (t + 1) * BIKE_PARAMS['dt']
        e = ptr[:-1]
        fut[:, t, 0] = pos[e, 0] + fwd * speed[e] * np.cos(head[e])
        fut[:, t, 1] = pos[e, 1] + fwd * speed[e] * np.sin(head[e])
        fut[:, t, 2] = np.cos(head[e])
        fut[:, t, 3] = np.sin(head[e])
    mean = np.array(STATE_MEAN)
    std = np.array(STATE_STD)
    past_n = (past - mean) / std
    fut_n = (fut - mean[:4]) / std[:4]
    lw = np.stack([l, w], 1)
    lw_n = (lw - np.array(ATT_MEAN)) / np.array(ATT_STD)
    sem = np.zeros((NA, 2))
    sem[np.arange(NA), truck.astype(np.int64)] = 1.0
    batch = np.repeat(np.arange(S), sizes).astype(np.int64)
    g = torch.Generator().manual_seed(seed + 7919)
    prior_mu = 0.3 * torch.randn(NA, 32, generator=g)
    prior_var = torch.exp(0.4 * torch.randn(NA, 32, generator=g))
    z = prior_mu + torch.sqrt(prior_var) * torch.randn(NA, 32, generator=g)
    map_feat = 0.5 * torch.randn(NA, 64, generator=g)
    past_feat = 0.5 * torch.randn(NA, 64, generator=g)
    out = {
        'past': torch.from_numpy(past_n).to(dtype),
        'lw': torch.from_numpy(lw_n).to(dtype),
        'sem': torch.from_numpy(sem).to(dtype),
        'ptr': torch.from_numpy(ptr),
        'batch': torch.from_numpy(batch),
        'map_idx': torch.from_numpy(map_idx),
        'edge_index': clique_edges(ptr),
        'z': z.to(dtype), 'prior_mu': prior_mu.to(dtype), 'prior_var': prior_var.to(dtype),
        'map_feat': map_feat.to(dtype), 'past_feat': past_feat.to(dtype),
        'ext_future': torch.from_numpy(fut_n).to(dtype),
    }
    return out


def _mlp_spec(prefix, sizes):
    """Key/shape list of reference models/common.py:8-44 MLP (Linear, then [LayerNorm, ReLU, Linear]*)."""
    spec = [(prefix + '.net.0.weight', (sizes[1], sizes[0])), (prefix + '.net.0.bias', (sizes[1],))]
    idx = 1
    for li in range(1, len(sizes) - 1):
        spec += [(prefix + '.net.%d.weight' % idx, (sizes[li],)), (prefix + '.net.%d.bias' % idx, (sizes[li],))]
        idx += 2
        spec += [(prefix + '.net.%d.weight' % idx, (sizes[li + 1], sizes[li])),
                 (prefix + '.net.%d.bias' % idx, (sizes[li + 1],))]
        idx += 1
    return spec


def decode_path_spec(NC=2):
    """(key, shape) for every tensor the decode path reads (SURVEY.md 8a / 5 checkpoint row)."""
    spec = []
    chans = [4, 16, 32, 64, 64, 128, 128]
    ks = [7, 5, 5, 3, 3, 3]
    for li in range(6):
        spec += [('map_conv.%d.weight' % (3 * li), (chans[li + 1], chans[li], ks[li], ks[li])),
                 ('map_conv.%d.bias' % (3 * li), (chans[li + 1],)),
                 ('map_conv.%d.weight' % (3 * li + 1), (chans[li + 1],)),
                 ('map_conv.%d.bias' % (3 * li + 1), (chans[li + 1],))]
    spec += [('map_feature.weight', (64, 512)), ('map_feature.bias', (64,))]
    dec_in = 32 + 64 + 64 + NC + 2
    spec += _mlp_spec('decoder_net.mlp_in', [dec_in, 128, 128, 64])
    spec += _mlp_spec('decoder_net.msg.0.edge_mlp', [2 * (64 + NC) + 4, 128, 128, 64])
    spec += _mlp_spec('decoder_net.msg.0.update_mlp', [64 + 64 + NC, 128, 64])
    spec += _mlp_spec('decoder_net.mlp_out', [64, 128, 128, 2])
    for layer in range(3):
        kin = 4 if layer == 0 else 64
        spec += [('decoder_memory.weight_ih_l%d' % layer, (192, kin)),
                 ('decoder_memory.weight_hh_l%d' % layer, (192, 64)),
                 ('decoder_memory.bias_ih_l%d' % layer, (192,)),
                 ('decoder_memory.bias_hh_l%d' % layer, (192,))]
    return spec


def make_weights(seed=0, NC=2, dtype=torch.float32):
    """Seeded decode-path state_dict. Norm scales ~1+-0.2, norm shifts/biases +-0.1, matrices U(+-1/sqrt(fan_in))
    (the reference's default nn init family), so every affine term is exercised."""
    g = torch.Generator().manual_seed(1000003 * (seed + 1))
    sd = {}
    norm_w = set()
    for k, shp in decode_path_spec(NC):
        is_norm = (len(shp) == 1 and k.endswith('.weight'))
        if is_norm:
            t = 1.0 + 0.4 * (torch.rand(shp, generator=g) - 0.5)
        elif len(shp) == 1:
            t = 0.2 * (torch.rand(shp, generator=g) - 0.5)
        else:
            fan_in = int(np.prod(shp[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            t = (2.0 * torch.rand(shp, generator=g) - 1.0) * bound
        sd[k] = t.to(dtype)
    return sd


def checksum(t):
    """Order-sensitive float64 checksum used to prove fixtures/inputs regenerate identically."""
    a = t.detach().to(torch.float64).reshape(-1)
    w = torch.arange(1, a.numel() + 1, dtype=torch.float64)
    return float((a * torch.cos(w * 0.37)).sum())


def host_producer_spec(NC=2, PT=4, FT=20):
    """(key, shape) of the once-per-batch producer modules (reference src/models/traffic_model.py:95-148): past / future MLP
    encoders and the prior / posterior interaction nets."""
    spec = []
    spec += _mlp_spec('past_encoder', [NC + PT * 9, 128, 128, 128, 64])
    spec += _mlp_spec('future_encoder', [NC + FT * 9, 128, 128, 128, 64])
    for name, din in (('prior_net', 64 + 64 + NC), ('posterior_net', 64 + 64 + 64 + NC)):
        spec += _mlp_spec(name + '.mlp_in', [din, 128, 128, 128])
        spec += _mlp_spec(name + '.msg.0.edge_mlp', [2 * (128 + NC) + 4, 128, 128, 128])
        spec += _mlp_spec(name + '.msg.0.update_mlp', [128 + 128 + NC, 128, 128])
        spec += _mlp_spec(name + '.mlp_out', [128, 128, 128, 64])
    return spec


def make_host_weights(seed=0, NC=2, PT=4, FT=20, dtype=torch.float32):
    """Seeded state_dict entries for `host_producer_spec` (same init family as make_weights; a separate generator so the
    decode-path weights -- and the fixtures pinned to them -- are unchanged)."""
    g = torch.Generator().manual_seed(7000003 * (seed + 1) + FT)
    sd = {}
    for k, shp in host_producer_spec(NC, PT, FT):
        is_norm = (len(shp) == 1 and k.endswith('.weight'))
        if is_norm:
            t = 1.0 + 0.4 * (torch.rand(shp, generator=g) - 0.5)
        elif len(shp) == 1:
            t = 0.2 * (torch.rand(shp, generator=g) - 0.5)
        else:
            bound = 1.0 / math.sqrt(int(np.prod(shp[1:])))
            t = (2.0 * torch.rand(shp, generator=g) - 1.0) * bound
        sd[k] = t.to(dtype)
    return sd


def make_future(seed, sc, FT, invis_frac=0.2, dtype=torch.float32):
    """Observed futures for the posterior / init paths: constant-velocity continuation of `past` (normalised, (NA,FT,6)) with
    noise, plus future_vis (NA,FT) and a past_vis (NA,PT) with a few unobserved frames (never the last past frame)."""
    g = torch.Generator().manual_seed(seed)
    past = sc['past']
    NA, PT = past.size(0), past.size(1)
    last = past[:, -1, :]
    vel = past[:, -1, :2] - past[:, -2, :2]
    steps = torch.arange(1, FT + 1).view(1, FT, 1).to(dtype)
    xy = last[:, None, :2] + vel[:, None, :] * steps + 0.02 * torch.randn(NA, FT, 2, generator=g).to(dtype)
    fut = torch.cat([xy, last[:, None, 2:].expand(NA, FT, 4)], dim=2).contiguous()
    fvis = (torch.rand(NA, FT, generator=g) > invis_frac).to(dtype)
    fvis[:, 0] = 1.0
    pvis = (torch.rand(NA, PT, generator=g) > invis_frac).to(dtype)
    pvis[:, -1] = 1.0
    return fut, fvis, pvis
