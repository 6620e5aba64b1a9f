"""BASELINE.json configs[2] on one GPU: adv_scenario_gen latent loops in planner-replay mode on ragged synthetic scenes
(~512 agents, 4..40 per scene): init (100 iters, FT 12), adversarial (400 iters, FT 12), solution (200 iters, FT 16), through
the drop-in modules (strive_b200.optim.run_*_optim = reference utils/{init,adv_gen,sol}_optim.py).  Prints wall-clock
throughput in agent*timestep*iter/s per phase (host timer around the whole loop, device synchronised on both sides)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import strive_b200
from strive_b200 import synth
from strive_b200.optim import run_init_optim, run_adv_gen_optim, run_find_solution_optim

ADV_W = {'coll_veh': 20.0, 'coll_veh_plan': 20.0, 'coll_env': 20.0, 'init_z': 0.5, 'init_z_atk': 0.05, 'motion_prior': 1.0,
         'motion_prior_atk': 0.005, 'motion_prior_ext': 0.0001, 'match_ext': 10.0, 'adv_crash': 2.0}          # configs/adv_gen_rule_based.cfg:34-43
SOL_W = {'sol_motion_prior': 0.005, 'sol_coll_veh': 10.0, 'sol_coll_env': 10.0, 'sol_motion_prior_ext': 0.001, 'sol_match_ext': 10.0,
         'sol_init_z': 0.0}                                                                                       # :45-50
INIT_W = {'init_match_ext': 10.0, 'init_motion_prior_ext': 0.1}                                                 # :28-30
ITERS = [int(x) for x in os.environ.get('ITERS', '100,400,200').split(',')]
dev = torch.device('cuda:0')
rng = np.random.RandomState(5)
sizes = []
while sum(sizes) < 512:
    sizes.append(int(rng.randint(4, 41)))
FT, FTs = 12, 16
raster, dx = synth.make_raster(seed=1, M=1, H=4096, W=4096)
model = strive_b200.make_model(nfuture=FT, state_dict=synth.make_weights(0), device=dev)
env = strive_b200.MapEnv(raster, dx, device=dev)
sc = synth.make_scenes(3000, sizes, map_extent_m=(200.0, 800.0), M=1, FT=FTs, collide_frac=0.5, offroad_frac=0.25)
NA = int(sc['ptr'][-1])


class G(object):
    pass


g = G()
for k in ('past', 'lw', 'sem', 'ptr', 'batch', 'edge_index'):
    setattr(g, k, sc[k].to(dev))
ego = torch.zeros(NA, dtype=torch.bool)
ego[sc['ptr'][:-1]] = True
pf = sc['ext_future'][:, :FT].contiguous()
fg = torch.zeros(NA, FT, 6)
fg[ego, :, :4] = pf
g.future_gt = fg.to(dev)
midx = sc['map_idx'].to(dev)
embed = {'map_feat': sc['map_feat'].to(dev), 'past_feat': sc['past_feat'].to(dev)}
tp = (sc['prior_mu'][ego].to(dev), sc['prior_var'][ego].to(dev))
op = (sc['prior_mu'][~ego].to(dev), sc['prior_var'][~ego].to(dev))
ap = (sc['prior_mu'].to(dev), sc['prior_var'].to(dev))
# observed futures for the init phase: constant-velocity continuation of the past
vel = sc['past'][:, -1, :2] - sc['past'][:, -2, :2]
steps = torch.arange(1, FT + 1).view(1, FT, 1).float()
init_traj = torch.cat([sc['past'][:, -1:, :2] + vel[:, None] * steps, sc['past'][:, -1:, 2:4].expand(NA, FT, 2)], 2).contiguous().to(dev)
vis = torch.ones(NA, FT, device=dev)


def timed(name, fn, units):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('%-10s %4d iters  %7.2f s  %8.1f ms/iter  %9.0f agent*timestep*iter/s' % (name, units[1], dt, 1000 * dt / units[1], units[0] * units[1] / dt))
    return out


print('configs[2]: %d scenes (%d..%d agents), NA %d' % (len(sizes), min(sizes), max(sizes), NA))
for _ in range(2):      # warm-up (allocator, lazy module loading)
    run_init_optim(sc['z'].to(dev), init_traj, vis, 0.1, INIT_W, model, g, env, midx, 1, embed, ap)
z, _, _ = timed('init', lambda: run_init_optim(sc['z'].to(dev), init_traj, vis, 0.1, INIT_W, model, g, env, midx, ITERS[0], embed, ap), (NA * FT, ITERS[0]))
z2, traj, _, min_agt, min_t = timed('adv', lambda: run_adv_gen_optim(z.detach(), 0.05, ADV_W, model, g, env, midx, ITERS[1], embed, 'ego', tp, op, 1, -0.5,
                                                                      future_len=FT, veh_coll_buffer=0.1), (NA * FT, ITERS[1]))
z3, sol, _ = timed('solution', lambda: run_find_solution_optim(z2, traj, FTs, 0.05, SOL_W, model, g, env, midx, ITERS[2], embed, tp, op), (NA * FTs, ITERS[2]))
print('finite:', bool(torch.isfinite(z3).all()), bool(torch.isfinite(sol).all()))
