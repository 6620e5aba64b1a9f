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
