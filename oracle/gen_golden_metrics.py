"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/metrics.npz: outputs of the UNMODIFIED reference functions
  datasets/nuscenes_utils.py check_on_layer (:266-298), check_line_layer (:300-333)
  losses/traffic_model.py compute_coll_rate_env (:366-419)
  utils/scenario_gen.py determine_feasibility_nusc (:30-107)
on the seeded inputs of tests/common.py metric_inputs().  Run in the build container only:
    python oracle/gen_golden_metrics.py
(check_single_veh_coll / check_pairwise_veh_coll need shapely, which is not installed: no golden, see oracle/metrics_oracle.py.)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims                      # noqa: E402
from oracle import metrics_oracle as MO            # noqa: E402
from strive_b200 import synth                      # noqa: E402
from tests.common import RASTER_KW, metric_inputs  # noqa: E402


class G(object):
    pass


def main():
    ref_shims.install()
    import datasets.nuscenes_utils as nutils
    from losses.traffic_model import compute_coll_rate_env
    from utils.scenario_gen import determine_feasibility_nusc
    raster, dx = synth.make_raster(**RASTER_KW)
    env = ref_shims.make_map_env(raster, dx)
    mi = metric_inputs()
    sc, samples, nrm, att = mi['sc'], mi['samples'], mi['nrm'], mi['att']
    NA, NS, FT = mi['NA'], mi['NS'], mi['FT']
    out = {}
    # check_on_layer on the valid frames of sample 0
    un = nrm.unnormalize(samples)
    cars = un[:, 0].reshape(NA * FT, 4)
    lw = att.unnormalize(sc['lw']).view(NA, 1, 2).expand(NA, FT, 2).reshape(NA * FT, 2)
    mix = sc['map_idx'][sc['batch']].view(NA, 1).expand(NA, FT).reshape(NA * FT)
    ok = ~torch.isnan(cars.sum(-1))
    out['on_layer_frac'] = nutils.check_on_layer(env.nusc_raster[:, 0], env.nusc_dx, cars[ok], lw[ok], mix[ok]).numpy()
    out['on_layer_frac_l2'] = nutils.check_on_layer(env.nusc_raster[:, 2], env.nusc_dx, cars[ok], lw[ok], mix[ok]).numpy()
    g = G()
    g.lw, g.batch, g.ptr = sc['lw'], sc['batch'], sc['ptr']
    cd = compute_coll_rate_env(g, sc['map_idx'], samples, env, nrm, att)
    out['env_did_collide'] = cd['did_collide'].numpy()
    out['env_num_coll'] = np.array(cd['num_coll_map'])
    # line / layer
    gen = torch.Generator().manual_seed(5)
    start = un[:, 1, 0, :2].clone()
    end = start + 30.0 * (torch.rand(NA, 2, generator=gen) - 0.5)
    out['line_start'], out['line_end'] = start.numpy(), end.numpy()
    out['line_hit'] = nutils.check_line_layer(env.nusc_raster[:, 0], env.nusc_dx, start, end, mix.view(NA, FT)[:, 0]).numpy()
    # feasibility of scene 0 (row 0 = ego)
    n0 = int(sc['ptr'][1])
    s0 = samples[:n0].clone()
    s0[torch.isnan(s0)] = 0.0
    for name, kw in (('feas_a', dict(feasibility_time=0, feasibility_vel=0.0, feasibility_infront_min=None, check_non_drivable_separation=True)),
                     ('feas_b', dict(feasibility_time=2, feasibility_vel=1.0, feasibility_infront_min=-0.5, check_non_drivable_separation=True)),
                     ('feas_c', dict(feasibility_time=1, feasibility_vel=0.5, feasibility_infront_min=0.0, check_non_drivable_separation=False))):
        f, ts, dist = determine_feasibility_nusc(s0.clone(), nrm, 10.0, map_env=env, map_idx=sc['map_idx'][0:1], **kw)
        out[name + '_feasible'], out[name + '_step'], out[name + '_dist'] = f.numpy(), ts.numpy(), dist.numpy()
    # oracle vs reference, right here
    d = float(np.abs(MO.check_on_layer(raster[:, 0], dx, cars[ok], lw[ok], mix[ok]).numpy() - out['on_layer_frac']).max())
    print('oracle check_on_layer vs reference: max diff %.3e' % d)
    print('env collisions: %d of %d' % (int(out['env_did_collide'].sum()), NA * NS), '| line hits', out['line_hit'].astype(int),
          '| feasible', out['feas_a_feasible'].astype(int), out['feas_b_feasible'].astype(int), out['feas_c_feasible'].astype(int))
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'metrics.npz'), **out)
    print('wrote metrics.npz')


if __name__ == '__main__':
    main()
