"""AdvGenLoss option branches (attack_agt_idx, no in-front filter, crash_min_t / buffer variants) on the GPU against the
unmodified reference's outputs (tests/golden/losses2.npz).

Written after this round's GPU minutes were spent: it has not run on a B200 yet (the CPU oracle is pinned on the same
fixture in tests/test_oracle_golden.py).  The file name sorts last so that `pytest -x` reaches it after every validated test.
"""
import numpy as np
import pytest
import torch

from oracle import strive_oracle as O
from tests.common import golden, scene_for, ADV_W
from tests.test_gpu_parity import ctx, diag

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', ['atk', 'noinfront'])
def test_adv_loss_option_branches_vs_reference(name):
    from strive_b200.losses import AdvGenLoss
    dev, model, env = ctx()
    g = golden('losses2')
    sc = scene_for(g)
    ptr = sc['ptr']
    NA = int(ptr[-1])
    ego = torch.zeros(NA, dtype=torch.bool)
    ego[ptr[:-1]] = True
    FT = int(g['FT'])
    fut_n = torch.from_numpy(g['fut_n'])
    lw_un = O.unnorm_att(sc['lw']).to(dev)
    mapixes = sc['map_idx'][sc['batch']].to(dev)
    tgt = O.unnorm_state(sc['ext_future'][:, :FT]).to(dev)
    prior_o = (sc['prior_mu'][~ego].to(dev), sc['prior_var'][~ego].to(dev))
    init_o = (sc['z'][~ego] - 0.03).to(dev)
    if name == 'atk':
        kw = dict(veh_coll_buffer=0.0, crash_loss_min_time=0, crash_loss_min_infront=None)
        fkw = dict(attack_agt_idx=(ptr[:-1] + torch.tensor([2, 1, 3])).to(dev))
    else:
        kw = dict(veh_coll_buffer=0.2, crash_loss_min_time=3, crash_loss_min_infront=None)
        fkw = {}
    fut = O.unnorm_state(fut_n).to(dev).requires_grad_(True)
    z_o = sc['z'][~ego].clone().to(dev).requires_grad_(True)
    adv = AdvGenLoss(ADV_W, lw_un, mapixes, env, init_o, ptr.to(dev), **kw)
    ld = adv(fut, tgt, z_o, prior_o, return_mins=True, **fkw)
    ld['loss'].backward()
    gl, gf, gz = float(g[name + '_loss']), g[name + '_d_fut'], g[name + '_d_z']
    e_f = np.abs(fut.grad.cpu().numpy() - gf).max()
    e_z = np.abs(z_o.grad.cpu().numpy() - gz).max()
    diag('adv[%s]: loss gpu %.4f golden %.4f | |d_fut| err %.3e (max %.3e) |d_z| err %.3e | mins %s %s vs %s %s' % (
        name, float(ld['loss']), gl, e_f, np.abs(gf).max(), e_z, ld['min_agt'], ld['min_t'], g[name + '_min_agt'], g[name + '_min_t']))
    assert abs(float(ld['loss']) - gl) < 1e-4 * abs(gl)
    assert e_f < 1e-3 * np.abs(gf).max() and e_z < 1e-4 * max(1.0, np.abs(gz).max())
    assert list(ld['min_agt']) == [int(v) for v in g[name + '_min_agt']] and list(ld['min_t']) == [int(v) for v in g[name + '_min_t']]
