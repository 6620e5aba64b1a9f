"""Tiny refine / adversarial iterations for compute-sanitizer (memcheck / racecheck / synccheck): ragged scenes incl. 1- and 2-agent
ones, every kernel of the rollout (tcgen05 edge kernels included), the losses and the device Adam, eager launches (no graph)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import strive_b200
from strive_b200 import synth
from strive_b200.optim import RefineLoop, AdvLoop
from strive_b200 import _cabi
_cabi.lib().strive_mapenc_set_pair(2)          # conv3 on CTA pairs whatever residency the tool leaves
dev = torch.device('cuda:0')
raster, dx = synth.make_raster(seed=3, M=2, H=1280, W=1280)
sd = synth.make_weights(0)
FT = int(os.environ.get('FT', '3'))
model = strive_b200.make_model(nfuture=FT, state_dict=sd, device=dev)
env = strive_b200.MapEnv(raster, dx, device=dev)
sc = synth.make_scenes(5, [3, 1, 9, 2, 18, 33], map_extent_m=(90.0, 230.0), M=2, FT=FT, collide_frac=1.0, offroad_frac=1.0)
class G(object):
    pass
g = G()
for k in ('past', 'lw', 'sem', 'ptr', 'batch', 'edge_index'):
    setattr(g, k, sc[k].to(dev))
embed = {'map_feat': sc['map_feat'].to(dev), 'past_feat': sc['past_feat'].to(dev), 'prior_out': (sc['prior_mu'].to(dev), sc['prior_var'].to(dev))}
w = {'coll_veh': 100.0, 'coll_env': 100.0, 'motion_prior': 1.0, 'init_z': 0.01}
loop = RefineLoop(model, g, sc['map_idx'].to(dev), env, embed, sc['z'].to(dev), w, 0.05, FT, veh_coll_buffer=0.2, group_scene_ptr=[0, 3, 6], use_graph=False)
loop.run(2)
adv_w = {'coll_veh': 20.0, 'coll_veh_plan': 20.0, 'coll_env': 20.0, 'init_z': 0.5, 'init_z_atk': 0.05, 'motion_prior': 1.0, 'motion_prior_atk': 0.005,
         'motion_prior_ext': 0.0001, 'match_ext': 10.0, 'adv_crash': 2.0}
adv = AdvLoop(model, g, sc['map_idx'].to(dev), env, embed, sc['z'].to(dev), sc['ext_future'][:, :FT].to(dev), adv_w, 0.05, FT, embed['prior_out'], veh_coll_buffer=0.1,
              crash_min_t=1, crash_min_infront=-0.5, use_graph=False)
adv.run(1)
torch.cuda.synchronize()
print('sanitize_small: refine loss %.4f, adv loss %.4f, finite %s' % (float(loop.terms[:, 0].sum()), float(adv.terms[:, 0].sum()),
                                                                     bool(torch.isfinite(loop.z).all() and torch.isfinite(adv.z).all())))
