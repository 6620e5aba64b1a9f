"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/init_loop.npz: the UNMODIFIED reference run_init_optim
(/root/reference/src/utils/init_optim.py:11-68) for 3 Adam iterations on the seeded case of tests/common.py init_case().
Run in the build container only:   python oracle/gen_golden_init.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import gen_golden as GG                # noqa: E402
from oracle import strive_oracle as O              # noqa: E402
from tests.common import init_case, INIT_W         # noqa: E402


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    raster, dx, env, model, sd = GG.build()
    from utils.init_optim import run_init_optim
    FT, iters, lr = 6, 3, 0.1                                                     # adv_scenario_gen.py:289 (init lr 0.1)
    sc, init_traj, vis = init_case(FT)
    model.FT = FT
    sg = GG.ref_graph(sc)
    embed = {'map_feat': sc['map_feat'], 'past_feat': sc['past_feat']}
    z, traj, _ = GG.quiet(run_init_optim, sc['z'], init_traj, vis, lr, INIT_W, model, sg, env, sc['map_idx'], iters, embed,
                          (sc['prior_mu'], sc['prior_var']))
    rec = []
    z_mine = O.init_loop(sd, sc, raster, dx, INIT_W, iters, lr, FT, init_traj, vis, record=rec)
    print('init loop: |z_ref - z_oracle| = %.3e, moved %.3e, loss0 %.5f' % (float((z - z_mine).abs().max()), float((z - sc['z']).abs().max()),
                                                                             rec[0]['loss']))
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'init_loop.npz'), z=z.detach().numpy(), traj=traj.numpy(), FT=FT, iters=iters, lr=lr,
                        z_in_sum=float(sc['z'].double().sum()), vis_sum=float(vis.sum()))
    print('wrote init_loop.npz')


if __name__ == '__main__':
    main()
