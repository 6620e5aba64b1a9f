"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/losses2.npz: the UNMODIFIED reference AdvGenLoss
(/root/reference/src/losses/adv_gen_nusc.py:57-262) on the option branches the main fixture does not take:
attack_agt_idx given (:118-123), crash_loss_min_infront=None (:124), crash_loss_min_time=0, veh_coll_buffer=0.
Run in the build container only:   python oracle/gen_golden_losses2.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import gen_golden as GG                # noqa: E402
from oracle import strive_oracle as O              # noqa: E402
from strive_b200 import synth                      # noqa: E402


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    raster, dx, env, model, sd = GG.build()
    from losses.adv_gen_nusc import AdvGenLoss
    seed, sizes, FT = 23, (5, 3, 4), 7
    scene = synth.make_scenes(seed, list(sizes), map_extent_m=GG.EXTENT, M=2, FT=FT, collide_frac=1.0, offroad_frac=1.0)
    ptr = scene['ptr']
    NA = int(ptr[-1])
    ego = torch.zeros(NA, dtype=torch.bool)
    ego[ptr[:-1]] = True
    with torch.no_grad():
        fut_n = GG.decode_ref(model, env, scene, scene['z'], FT)
    lw_un = model.get_att_normalizer().unnormalize(scene['lw'])
    mapixes = scene['map_idx'][scene['batch']]
    tgt = O.unnorm_state(scene['ext_future'][:, :FT]).clone()
    prior_o = (scene['prior_mu'][~ego], scene['prior_var'][~ego])
    init_o = (scene['z'][~ego] - 0.03).clone()
    out = dict(seed=seed, sizes=np.array(sizes), FT=FT, fut_n=fut_n.numpy())
    for name, kw, fkw in (('atk', dict(veh_coll_buffer=0.0, crash_loss_min_time=0, crash_loss_min_infront=None),
                           dict(attack_agt_idx=(ptr[:-1] + torch.tensor([2, 1, 3])))),      # GLOBAL agent indices, as adv_gen_optim.py:54-55 passes them
                          ('noinfront', dict(veh_coll_buffer=0.2, crash_loss_min_time=3, crash_loss_min_infront=None), dict())):
        fut = O.unnorm_state(fut_n).clone().requires_grad_(True)
        z_o = scene['z'][~ego].clone().requires_grad_(True)
        adv = GG.quiet(AdvGenLoss, GG.ADV_W, lw_un, mapixes, env, init_o, ptr, **kw)
        ld = GG.quiet(adv, fut, tgt, z_o, prior_o, return_mins=True, **fkw)
        ld['loss'].backward()
        fut2 = O.unnorm_state(fut_n).clone().requires_grad_(True)
        z2 = scene['z'][~ego].clone().requires_grad_(True)
        okw = dict(veh_coll_buffer=kw['veh_coll_buffer'], crash_min_t=kw['crash_loss_min_time'], crash_min_infront=kw['crash_loss_min_infront'])
        if 'attack_agt_idx' in fkw:
            okw['attack_agt_idx'] = fkw['attack_agt_idx']
        md = O.adv_gen_loss(fut2, tgt, z2, prior_o, init_o, GG.ADV_W, lw_un, mapixes, ptr, raster, dx, **okw)
        md['loss'].backward()
        print('%s: loss ref %.6f oracle %.6f | d_fut maxdiff %.3e | d_z maxdiff %.3e | mins %s %s vs %s %s' % (
            name, float(ld['loss']), float(md['loss']), GG.maxdiff(fut.grad, fut2.grad), GG.maxdiff(z_o.grad, z2.grad),
            list(ld['min_agt']), list(ld['min_t']), md['min_agt'], md['min_t']))
        out[name + '_loss'] = float(ld['loss'])
        out[name + '_d_fut'] = fut.grad.numpy()
        out[name + '_d_z'] = z_o.grad.numpy()
        out[name + '_min_agt'] = np.array(ld['min_agt'])
        out[name + '_min_t'] = np.array(ld['min_t'])
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'losses2.npz'), **out)
    print('wrote losses2.npz')


if __name__ == '__main__':
    main()
