"""Shared fixtures for the parity tests: the small synthetic world the golden vectors were generated on
(must mirror oracle/gen_golden.py: RASTER_KW / EXTENT / weight seed 0)."""
import os

import numpy as np
import torch

from strive_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
RASTER_KW = dict(seed=3, M=2, H=1280, W=1280)
EXTENT = (90.0, 230.0)
REFINE_W = {'coll_veh': 100.0, 'coll_env': 100.0, 'motion_prior': 1.0, 'init_z': 0.01}
ADV_W = {'coll_veh': 20.0, 'coll_veh_plan': 20.0, 'coll_env': 20.0, 'init_z': 0.5, 'init_z_atk': 0.05,
         'motion_prior': 1.0, 'motion_prior_atk': 0.005, 'motion_prior_ext': 0.0001, 'match_ext': 10.0,
         'adv_crash': 2.0}
SOL_W = {'motion_prior': 0.005, 'coll_veh': 10.0, 'coll_env': 10.0, 'motion_prior_ext': 0.001,
         'match_ext': 10.0, 'init_z': 0.0}

_cache = {}


def world():
    """raster, dx, weights for the golden world; verifies the regenerated inputs against stored checksums."""
    if 'w' not in _cache:
        raster, dx = synth.make_raster(**RASTER_KW)
        sd = synth.make_weights(0)
        meta = np.load(os.path.join(GOLD, 'meta.npz'))
        assert float(raster.double().sum()) == float(meta['raster_sum'])
        wsum = synth.checksum(torch.cat([v.reshape(-1) for v in sd.values()]))
        assert abs(wsum - float(meta['weights_sum'])) < 1e-9, 'seeded weights did not regenerate identically'
        assert np.array_equal(dx.numpy(), meta['dx'])
        _cache['w'] = (raster, dx, sd)
    return _cache['w']


def golden(name):
    return np.load(os.path.join(GOLD, name + '.npz'))


def scene_for(g, **kw):
    sizes = [int(v) for v in g['sizes']]
    args = dict(map_extent_m=EXTENT, M=2, FT=int(g['FT']), collide_frac=1.0, offroad_frac=1.0)
    args.update(kw)
    return synth.make_scenes(int(g['seed']), sizes, **args)


def metric_inputs():
    """Seeded inputs of the success / plausibility checks (tests/golden/metrics.npz holds the reference's outputs for them):
    two scenes on the golden world, NS sampled futures per agent = constant-velocity rollouts with per-sample heading / speed
    noise, a few NaN frames, some agents steered off the road."""
    from strive_b200.traffic_model import MeanStdNormalizer, STATE_MEAN, STATE_STD, ATT_MEAN, ATT_STD
    sc = synth.make_scenes(41, [5, 3], map_extent_m=EXTENT, M=2, FT=8, collide_frac=1.0, offroad_frac=1.0)
    NA, NS, FT = sc['past'].size(0), 4, 8
    g = torch.Generator().manual_seed(77)
    nrm = MeanStdNormalizer(torch.tensor(STATE_MEAN), torch.tensor(STATE_STD))
    att = MeanStdNormalizer(torch.tensor(ATT_MEAN), torch.tensor(ATT_STD))
    last = nrm.unnormalize(sc['past'][:, -1, :])                      # (NA,6) x,y,hx,hy,s,hdot
    h0 = torch.atan2(last[:, 3], last[:, 2]).view(NA, 1, 1)
    h = h0 + 0.5 * (torch.rand(NA, NS, 1, generator=g) - 0.5) + 0.05 * torch.arange(FT).view(1, 1, FT) * (torch.rand(NA, NS, 1, generator=g) - 0.5)
    spd = (last[:, 4].view(NA, 1, 1) * (0.5 + torch.rand(NA, NS, 1, generator=g))).expand(NA, NS, FT)
    step = 0.5 * spd
    x = last[:, 0].view(NA, 1, 1) + torch.cumsum(step * torch.cos(h), dim=2)
    y = last[:, 1].view(NA, 1, 1) + torch.cumsum(step * torch.sin(h), dim=2)
    fut_un = torch.stack([x, y, torch.cos(h), torch.sin(h)], dim=3)
    fut_un[1, 2, 3:] = float('nan')
    fut_un[6, 0, 0] = float('nan')
    samples = nrm.normalize(fut_un)
    return dict(sc=sc, samples=samples, nrm=nrm, att=att, NA=NA, NS=NS, FT=FT)


INIT_W = {'init_match_ext': 10.0, 'init_motion_prior_ext': 0.1}       # configs/adv_gen_rule_based.cfg:28-30


def init_case(FT=6):
    """Seeded inputs of the init-loop case (tests/golden/init_loop.npz): observed futures = constant-velocity continuations of
    the past with noise, ~25 % of the (agent, step) entries invisible."""
    sc = synth.make_scenes(61, [4, 2, 3], map_extent_m=EXTENT, M=2, FT=FT, collide_frac=0.0, offroad_frac=0.0)
    NA = sc['past'].size(0)
    g = torch.Generator().manual_seed(62)
    last = sc['past'][:, -1, :4]
    vel = sc['past'][:, -1, :2] - sc['past'][:, -2, :2]
    steps = torch.arange(1, FT + 1).view(1, FT, 1).float()
    xy = last[:, None, :2] + vel[:, None, :] * steps + 0.02 * torch.randn(NA, FT, 2, generator=g)
    init_traj = torch.cat([xy, last[:, None, 2:4].expand(NA, FT, 2)], dim=2).contiguous()
    vis = (torch.rand(NA, FT, generator=g) > 0.25).float()
    vis[:, 0] = 1.0
    return sc, init_traj, vis


class StubPlanner(object):
    """Deterministic numpy stand-in for the reference's rule-based planner with its call surface (reset / rollout,
    src/planners/hardcode_goalcond_nusc.py:109,178): the ego keeps its initial heading and speed and brakes while any other
    agent's PREDICTED position lies in a corridor ahead of it -- so its answer reacts smoothly to the latents being optimised,
    which is all the closed-loop adversarial loop needs from a planner.  Used unchanged by oracle/gen_golden_r2.py (with the
    unmodified reference loop) and by the GPU tests (with the strive_b200 loop)."""

    def __init__(self):
        self.calls = 0

    def reset(self, init_state, vehicle_atts, batch_mask, batch_size, map_idx, ego_idx=0):
        st = init_state.detach().cpu().numpy()
        bm = batch_mask.detach().cpu().numpy()
        self.B = int(batch_size)
        self.ego = np.stack([st[bm == b][ego_idx] for b in range(self.B)])          # (B,6) x,y,hx,hy,s,hdot

    def rollout(self, agent_obs, agent_t, agent_ptr, planner_t, init_state=None, control_all=False, viz=None, coll_t=None):
        self.calls += 1
        obs = np.asarray(agent_obs, dtype=np.float64)                                # (NA-B, T, 4) unnormalised
        T = obs.shape[1]
        dt = float(planner_t[0])
        out = np.zeros((self.B, T, 4))
        for b in range(self.B):
            x, y, hx, hy, s, _ = [float(v) for v in self.ego[b]]
            others = obs[int(agent_ptr[b]):int(agent_ptr[b + 1])]
            for t in range(T):
                brake = 0.0
                for o in others:
                    dx, dy = o[t, 0] - x, o[t, 1] - y
                    ahead, side = dx * hx + dy * hy, -dx * hy + dy * hx
                    # smooth proximity: 1 when the agent sits right in front, fading over 15 m ahead / 2.5 m to the side
                    brake = max(brake, float(np.exp(-max(ahead, 0.0) / 15.0 - (side / 2.5) ** 2)) if ahead > -2.0 else 0.0)
                s = max(0.0, s - 3.0 * dt * brake)
                x, y = x + hx * s * dt, y + hy * s * dt
                out[b, t] = [x, y, hx, hy]
        return torch.from_numpy(out).float()


def scenario_inputs():
    """Seeded inputs of the scenario writer (tests/golden/scenario.json holds the reference writer's output for them)."""
    sc = synth.make_scenes(77, [4], map_extent_m=EXTENT, M=2, FT=5, collide_frac=1.0, offroad_frac=0.0)
    g = torch.Generator().manual_seed(78)
    NA = sc['z'].size(0)

    class SG(object):
        pass
    sg = SG()
    sg.past_gt, sg.lw, sg.sem = sc['past'], sc['lw'], sc['sem']
    kw = dict(init_fut_traj=torch.randn(NA, 5, 4, generator=g), adv_fut_traj=torch.randn(NA, 5, 4, generator=g),
              sol_fut_traj=torch.randn(NA, 5, 4, generator=g), attack_agt=2, attack_t=3, adv_z=torch.randn(NA, 32, generator=g),
              sol_z=torch.randn(NA, 32, generator=g), prior_distrib=(sc['prior_mu'], sc['prior_var']),
              internal_ego_traj=torch.randn(5, 4, generator=g))

    class Env(object):
        map_list = ['map-a', 'map-b']
    return sg, kw, Env()
