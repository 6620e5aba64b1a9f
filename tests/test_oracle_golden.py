"""CPU: pins oracle/strive_oracle.py against golden vectors produced by the UNMODIFIED reference
(tests/golden/*.npz, generator oracle/gen_golden.py).  The tolerances are fp32 re-association noise
amplified by the rollout (see DESIGN.md 'Tolerances'); single-function cases are tight."""
import numpy as np
import torch

from oracle import strive_oracle as O
from tests.common import world, golden, scene_for, REFINE_W, ADV_W, SOL_W


def _decode(sd, raster, dx, sc, z, FT, ext=None):
    return O.decode(sd, z, sc['map_feat'], sc['past_feat'], sc['past'][:, -1, :], sc['lw'], sc['sem'], sc['ptr'],
                    sc['edge_index'], sc['map_idx'], raster, dx, FT, ext_future=ext)


def test_encode_map_bit_exact_crop_and_feature():
    raster, dx, sd = world()
    g = golden('encode_map')
    pos_n = torch.from_numpy(g['pos_n'])
    mapix = torch.from_numpy(g['mapix'])
    crop = O.map_crop(raster, dx, O.unnorm_state(pos_n), mapix)
    assert np.array_equal(crop.long().sum(dim=3).numpy(), g['crop_rowsum'])      # integer work: bit exact
    feat = O.map_cnn(sd, crop.float())
    assert np.abs(feat.numpy() - g['map_feat']).max() < 2e-6


def test_decode_small_and_ext():
    raster, dx, sd = world()
    for name, tol in (('decode_small', 2e-4), ('decode_ext', 2e-4)):
        g = golden(name)
        sc = scene_for(g)
        assert abs(float(g['z_sum']) - __import__('strive_b200').synth.checksum(sc['z'])) < 1e-9
        with torch.no_grad():
            traj = _decode(sd, raster, dx, sc, sc['z'], int(g['FT']), sc['ext_future'] if int(g['with_ext']) else None)
        assert np.abs(traj.numpy() - g['traj']).max() < tol, name
        # first step has no amplification at all
        assert np.abs(traj.numpy()[:, 0] - g['traj'][:, 0]).max() < 1e-6


def test_decode_c1_config():
    """BASELINE.json configs[0]: 1 scene x 8 agents, 4 past / 20 future, CPU plumbing."""
    raster, dx, sd = world()
    g = golden('decode_c1')
    sc = scene_for(g)
    with torch.no_grad():
        traj = _decode(sd, raster, dx, sc, sc['z'], 20)
    d = np.abs(traj.numpy() - g['traj']).max(axis=(0, 2))
    assert d[0] < 1e-6 and d[:5].max() < 1e-4 and d.max() < 0.1      # rounding noise grows ~1.6x / step


def test_refine_loop_matches_reference_adam_trajectory():
    raster, dx, sd = world()
    g = golden('refine')
    sc = scene_for(g)
    rec = []
    z = O.refine_loop(sd, sc, raster, dx, REFINE_W, int(g['iters']), float(g['lr']), int(g['FT']), veh_coll_buffer=0.2, record=rec)
    assert abs(rec[0]['loss'] - g['loss'][0]) < 1e-3 * abs(g['loss'][0])
    assert np.abs(rec[0]['grad'].numpy() - g['grad'][0]).max() < 5e-4
    assert np.abs(z.numpy() - g['z'][-1]).max() < 5e-4
    assert g['terms'][0][4] > 1 and g['terms'][0][5] > 1      # both collision terms active in the fixture


def test_loss_modules():
    raster, dx, sd = world()
    g = golden('losses')
    sc = scene_for(g)
    ptr = sc['ptr']
    NA = int(ptr[-1])
    ego = torch.zeros(NA, dtype=torch.bool)
    ego[ptr[:-1]] = True
    FT = int(g['FT'])
    fut_n = torch.from_numpy(g['fut_n'])
    lw_un = O.unnorm_att(sc['lw'])
    mapixes = sc['map_idx'][sc['batch']]
    tgt = O.unnorm_state(sc['ext_future'][:, :FT])
    fut = O.unnorm_state(fut_n).requires_grad_(True)
    z_o = sc['z'][~ego].clone().requires_grad_(True)
    prior_o = (sc['prior_mu'][~ego], sc['prior_var'][~ego])
    ld = O.adv_gen_loss(fut, tgt, z_o, prior_o, sc['z'][~ego] + 0.05, ADV_W, lw_un, mapixes, ptr, raster, dx,
                        veh_coll_buffer=0.1, crash_min_t=2, crash_min_infront=-0.5)
    ld['loss'].backward()
    assert abs(float(ld['loss']) - float(g['adv_loss'])) < 1e-4 * abs(float(g['adv_loss']))
    assert np.abs(fut.grad.numpy() - g['adv_d_fut']).max() < 1e-3 * np.abs(g['adv_d_fut']).max()
    assert np.abs(z_o.grad.numpy() - g['adv_d_z']).max() < 1e-5
    assert list(ld['min_agt']) == list(g['adv_min_agt']) and list(ld['min_t']) == list(g['adv_min_t'])
    futm = O.unnorm_state(fut_n)[ego].requires_grad_(True)
    lm = O.tgt_matching_loss(futm, tgt, ADV_W)
    lm['loss'].backward()
    assert abs(float(lm['loss']) - float(g['match_loss'])) < 1e-5 * abs(float(g['match_loss']))
    assert np.abs(futm.grad.numpy() - g['match_d_fut']).max() < 1e-5
    futs = O.unnorm_state(fut_n).requires_grad_(True)
    zs = sc['prior_mu'][ego].clone().requires_grad_(True)
    ls = O.avoid_coll_loss(futs, zs, (sc['prior_mu'][ego], sc['prior_var'][ego]), zs.detach().clone(), SOL_W, lw_un,
                           mapixes, ptr, raster, dx, veh_coll_buffer=0.5, single_veh_idx=0)
    ls['loss'].backward()
    assert abs(float(ls['loss']) - float(g['sol_loss'])) < 1e-4 * abs(float(g['sol_loss']))
    assert np.abs(futs.grad.numpy() - g['sol_d_fut']).max() < 1e-4


# ----------------------------------------------------------------------------------------------------------
# success / plausibility checks: oracle/metrics_oracle.py against the unmodified reference
# ----------------------------------------------------------------------------------------------------------
def _metric_case():
    from tests.common import metric_inputs
    mi = metric_inputs()
    g = golden('metrics')
    raster, dx, _ = world()
    return mi, g, raster, dx


def test_metrics_oracle_raster_checks_match_reference():
    from oracle import metrics_oracle as MO
    mi, g, raster, dx = _metric_case()
    sc, nrm, att = mi['sc'], mi['nrm'], mi['att']
    NA, NS, FT = mi['NA'], mi['NS'], mi['FT']
    un = nrm.unnormalize(mi['samples'])
    cars = un[:, 0].reshape(NA * FT, 4)
    lw_un = att.unnormalize(sc['lw'])
    lw = lw_un.view(NA, 1, 2).expand(NA, FT, 2).reshape(NA * FT, 2)
    mix_a = sc['map_idx'][sc['batch']]
    mix = mix_a.view(NA, 1).expand(NA, FT).reshape(NA * FT)
    ok = ~torch.isnan(cars.sum(-1))
    assert np.array_equal(MO.check_on_layer(raster[:, 0], dx, cars[ok], lw[ok], mix[ok]).numpy(), g['on_layer_frac'])
    assert np.array_equal(MO.check_on_layer(raster[:, 2], dx, cars[ok], lw[ok], mix[ok]).numpy(), g['on_layer_frac_l2'])
    assert np.array_equal(MO.compute_coll_rate_env(sc['lw'], mix_a, un, lw_un, raster, dx).numpy(), g['env_did_collide'])
    assert 0 < int(g['env_did_collide'].sum()) < NA * NS
    hit = MO.check_line_layer(raster[:, 0], dx, torch.from_numpy(g['line_start']), torch.from_numpy(g['line_end']), mix_a)
    assert np.array_equal(hit.numpy(), g['line_hit'])


def test_metrics_oracle_feasibility_matches_reference():
    from oracle import metrics_oracle as MO
    mi, g, raster, dx = _metric_case()
    sc, nrm = mi['sc'], mi['nrm']
    n0 = int(sc['ptr'][1])
    s0 = mi['samples'][:n0].clone()
    s0[torch.isnan(s0)] = 0.0
    for name, (ft, fv, fi, sep) in (('feas_a', (0, 0.0, None, True)), ('feas_b', (2, 1.0, -0.5, True)), ('feas_c', (1, 0.5, 0.0, False))):
        f, ts, dist = MO.determine_feasibility(nrm.unnormalize(s0.clone()), 10.0, ft, fv, fi, sep, raster[:, 0], dx, sc['map_idx'][0:1])
        assert np.array_equal(f.numpy(), g[name + '_feasible'])
        assert np.array_equal(ts.numpy(), g[name + '_step'])
        assert np.allclose(dist.numpy(), g[name + '_dist'], rtol=0, atol=0)


def test_rect_iou_closed_form_cases():
    """Known answers anchoring the polygon arithmetic that the reference delegates to shapely (absent: parity unpinned)."""
    from oracle import metrics_oracle as MO

    def rect(x, y, ang, l, w):
        return MO.get_corners(np.array([x, y, np.cos(ang), np.sin(ang)], dtype=np.float32), np.array([l, w], dtype=np.float32))
    assert abs(MO.rect_iou(rect(0, 0, 0, 4, 2), rect(0, 0, 0, 4, 2)) - 1.0) < 1e-12
    assert abs(MO.rect_iou(rect(0, 0, 0, 2, 1), rect(1, 0, 0, 2, 1)) - 1.0 / 3.0) < 1e-6
    assert MO.rect_iou(rect(0, 0, 0, 2, 1), rect(5, 0, 0.3, 2, 1)) == 0.0
    assert abs(MO.rect_iou(rect(0, 0, 0, 2, 2), rect(0, 0, np.pi / 4, np.sqrt(2.0), np.sqrt(2.0))) - 0.5) < 1e-6
    assert abs(MO.rect_iou(rect(0, 0, 0, 4, 1), rect(0, 0, np.pi / 2, 4, 1)) - 1.0 / 7.0) < 1e-6
    assert abs(MO.rect_iou(rect(3, -2, 1.1, 4.5, 2), rect(3, -2, 1.1 + np.pi, 4.5, 2)) - 1.0) < 1e-5      # same box, heading flipped
    # loop semantics: first colliding step; NaN frames skipped; pairwise flags only the earlier agent of a pair, once
    T = 4
    tgt = np.array([[t * 1.0, 0.0, 1.0, 0.0] for t in range(T)], dtype=np.float32)
    others = np.stack([np.array([[3.0, 0.0, 1.0, 0.0]] * T, dtype=np.float32),
                       np.array([[0.0, 9.0, 1.0, 0.0]] * T, dtype=np.float32),
                       np.array([[0.0, 0.5, 1.0, 0.0]] * T, dtype=np.float32)])
    others[2, 0] = np.nan
    lw = np.array([[2.0, 1.0]] * 3, dtype=np.float32)
    coll, when = MO.check_single_veh_coll(tgt, np.array([2.0, 1.0], dtype=np.float32), others, lw)
    assert coll.tolist() == [True, False, True] and when.tolist() == [2, T, 1]
    pw = MO.check_pairwise_veh_coll(np.concatenate([tgt[None], others]), np.array([[2.0, 1.0]] * 4, dtype=np.float32))
    assert pw['did_collide'].tolist() == [True, False, False, False] and pw['num_coll_veh'] == 1.0 and pw['num_traj_veh'] == 4.0


def test_init_loop_matches_reference():
    """run_init_optim (utils/init_optim.py:11-68): 3 Adam iterations with a partial visibility mask."""
    from tests.common import init_case, INIT_W
    g = golden('init_loop')
    raster, dx, sd = world()
    FT, iters, lr = int(g['FT']), int(g['iters']), float(g['lr'])
    sc, init_traj, vis = init_case(FT)
    assert float(sc['z'].double().sum()) == float(g['z_in_sum']) and float(vis.sum()) == float(g['vis_sum'])
    z = O.init_loop(sd, sc, raster, dx, INIT_W, iters, lr, FT, init_traj, vis)
    assert np.abs(z.numpy() - g['z']).max() < 1e-4


def test_adv_loss_option_branches_match_reference():
    """AdvGenLoss branches the main fixture does not take (adv_gen_nusc.py:118-123 attack_agt_idx, :116 crash_loss_min_infront=None,
    crash_loss_min_time 0 / 3, veh_coll_buffer 0 / 0.2): oracle against the unmodified reference's loss, gradients and arg-mins."""
    g = golden('losses2')
    raster, dx, sd = world()
    FT = int(g['FT'])
    sc = scene_for(g)
    ptr = sc['ptr']
    NA = int(ptr[-1])
    ego = torch.zeros(NA, dtype=torch.bool)
    ego[ptr[:-1]] = True
    lw_un = O.unnorm_att(sc['lw'])
    mapixes = sc['map_idx'][sc['batch']]
    tgt = O.unnorm_state(sc['ext_future'][:, :FT]).clone()
    prior_o = (sc['prior_mu'][~ego], sc['prior_var'][~ego])
    init_o = (sc['z'][~ego] - 0.03).clone()
    fut_n = torch.from_numpy(g['fut_n'])
    for name, kw in (('atk', dict(veh_coll_buffer=0.0, crash_min_t=0, crash_min_infront=None, attack_agt_idx=(ptr[:-1] + torch.tensor([2, 1, 3])))),
                     ('noinfront', dict(veh_coll_buffer=0.2, crash_min_t=3, crash_min_infront=None))):
        fut = O.unnorm_state(fut_n).clone().requires_grad_(True)
        z = sc['z'][~ego].clone().requires_grad_(True)
        ld = O.adv_gen_loss(fut, tgt, z, prior_o, init_o, ADV_W, lw_un, mapixes, ptr, raster, dx, **kw)
        ld['loss'].backward()
        assert abs(float(ld['loss']) - float(g[name + '_loss'])) < 1e-5 * abs(float(g[name + '_loss']))
        gf = g[name + '_d_fut']
        assert np.abs(fut.grad.numpy() - gf).max() < 1e-5 * max(1.0, np.abs(gf).max())
        assert np.abs(z.grad.numpy() - g[name + '_d_z']).max() < 1e-6
        assert [int(v) for v in ld['min_agt']] == [int(v) for v in g[name + '_min_agt']]
        assert [int(v) for v in ld['min_t']] == [int(v) for v in g[name + '_min_t']]


def test_rect_iou_against_an_independent_construction():
    """Second, structurally different computation of the intersection area the reference gets from shapely (absent here): the
    intersection of two convex quadrilaterals is the convex hull of {corners of A inside B} + {corners of B inside A} + {edge x edge
    crossing points} (scipy.spatial.ConvexHull), against the oracle's Sutherland-Hodgman clipping, on random rotated rectangles
    of vehicle proportions incl. touching / containing / disjoint configurations."""
    from scipy.spatial import ConvexHull
    from oracle import metrics_oracle as MO

    def inside(p, quad):      # CCW quad, closed
        for e in range(4):
            a, b = quad[e], quad[(e + 1) % 4]
            if (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0]) < -1e-12:
                return False
        return True

    def seg_x(p, q, a, b):
        r, s = q - p, b - a
        den = r[0] * s[1] - r[1] * s[0]
        if abs(den) < 1e-14:
            return None
        t = ((a[0] - p[0]) * s[1] - (a[1] - p[1]) * s[0]) / den
        u = ((a[0] - p[0]) * r[1] - (a[1] - p[1]) * r[0]) / den
        return p + t * r if (0.0 <= t <= 1.0 and 0.0 <= u <= 1.0) else None

    def area(quad):
        x, y = quad[:, 0], quad[:, 1]
        return 0.5 * abs(float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y)))

    rng = np.random.RandomState(11)
    worst, nonzero = 0.0, 0
    for it in range(400):
        l1, w1, l2, w2 = rng.uniform(3.5, 6.0), rng.uniform(1.6, 2.4), rng.uniform(3.5, 12.0), rng.uniform(1.6, 3.0)
        a1, a2 = rng.uniform(-np.pi, np.pi, 2)
        c2 = rng.uniform(-5.0, 5.0, 2) if it % 5 else np.zeros(2)          # every fifth pair concentric (containment / crossings)
        A = np.asarray(MO.get_corners(np.array([0.0, 0.0, np.cos(a1), np.sin(a1)], dtype=np.float32), np.array([l1, w1], dtype=np.float32)), dtype=np.float64)
        B = np.asarray(MO.get_corners(np.array([c2[0], c2[1], np.cos(a2), np.sin(a2)], dtype=np.float32), np.array([l2, w2], dtype=np.float32)), dtype=np.float64)
        if 0.5 * float(np.sum(A[:, 0] * np.roll(A[:, 1], -1) - np.roll(A[:, 0], -1) * A[:, 1])) < 0:      # make both CCW for inside()
            A = A[::-1]
        if 0.5 * float(np.sum(B[:, 0] * np.roll(B[:, 1], -1) - np.roll(B[:, 0], -1) * B[:, 1])) < 0:
            B = B[::-1]
        pts = [p for p in A if inside(p, B)] + [p for p in B if inside(p, A)]
        for i in range(4):
            for j in range(4):
                x = seg_x(A[i], A[(i + 1) % 4], B[j], B[(j + 1) % 4])
                if x is not None:
                    pts.append(x)
        inter = 0.0
        if len(pts) >= 3:
            P = np.unique(np.round(np.array(pts), 12), axis=0)
            if len(P) >= 3:
                try:
                    inter = float(ConvexHull(P).volume)
                except Exception:          # degenerate (collinear) point set: zero area
                    inter = 0.0
        union = area(A) + area(B) - inter
        ref = inter / union if union > 0 else 0.0
        got = MO.rect_iou(A, B)
        worst = max(worst, abs(got - ref))
        nonzero += ref > 0
    assert nonzero > 150, nonzero
    assert worst < 1e-9, worst
