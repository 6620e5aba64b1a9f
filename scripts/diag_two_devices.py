"""Development: device-resident loops with CUDA-graph replay on two devices of ONE process."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import strive_b200
from strive_b200 import synth
from strive_b200.optim import RefineLoop, AdvLoop
raster, dx = synth.make_raster(seed=3, M=2, H=1280, W=1280)
sd = synth.make_weights(0)
FT = 4
sc = synth.make_scenes(5, [3, 1, 9, 2, 18, 33], map_extent_m=(90.0, 230.0), M=2, FT=FT, collide_frac=1.0, offroad_frac=1.0)
w = {'coll_veh': 100.0, 'coll_env': 100.0, 'motion_prior': 1.0, 'init_z': 0.01}
adv_w = {'coll_veh': 20.0, 'coll_veh_plan': 20.0, 'coll_env': 20.0, 'init_z': 0.5, 'init_z_atk': 0.05, 'motion_prior': 1.0, 'motion_prior_atk': 0.005,
         'motion_prior_ext': 0.0001, 'match_ext': 10.0, 'adv_crash': 2.0}
class G(object):
    pass
res = []
for d in range(min(2, torch.cuda.device_count())):
    dev = torch.device('cuda:%d' % d)
    model = strive_b200.make_model(nfuture=FT, state_dict=sd, device=dev)
    env = strive_b200.MapEnv(raster, dx, device=dev)
    g = G()
    for k in ('past', 'lw', 'sem', 'ptr', 'batch', 'edge_index'):
        setattr(g, k, sc[k].to(dev))
    embed = {'map_feat': sc['map_feat'].to(dev), 'past_feat': sc['past_feat'].to(dev), 'prior_out': (sc['prior_mu'].to(dev), sc['prior_var'].to(dev))}
    loop = RefineLoop(model, g, sc['map_idx'].to(dev), env, embed, sc['z'].to(dev), w, 0.05, FT, veh_coll_buffer=0.2, group_scene_ptr=[0, 3, 6])
    loop.run(3)
    adv = AdvLoop(model, g, sc['map_idx'].to(dev), env, embed, sc['z'].to(dev), sc['ext_future'][:, :FT].to(dev), adv_w, 0.05, FT, embed['prior_out'],
                  veh_coll_buffer=0.1, crash_min_t=1, crash_min_infront=-0.5)
    adv.run(3)
    torch.cuda.synchronize(dev)
    res.append((loop.traj.cpu(), adv.traj.cpu(), loop.z.cpu()))
    print('device %d: refine loss %.4f adv loss %.4f graphs %s %s' % (d, float(loop.terms[:, 0].sum()), float(adv.terms[:, 0].sum()), loop.graph is not None, adv.graph is not None))
if len(res) == 2:
    print('two devices, one process: |traj diff| refine %.3e adv %.3e, |z diff| %.3e' % (
        float((res[0][0] - res[1][0]).abs().max()), float((res[0][1] - res[1][1]).abs().max()), float((res[0][2] - res[1][2]).abs().max())))
