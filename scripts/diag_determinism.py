"""Runs the forward rollout of the bench batch several times and reports whether the trajectories are bitwise identical
(development: programmatic-dependent-launch bisect, STRIVE_PDL=<mask>)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import strive_b200
import bench
from strive_b200.optim import RefineLoop
dev = torch.device('cuda:0')
raster, dx, sd, sc = bench.make_workload(0)
model = strive_b200.make_model(nfuture=20, state_dict=sd, device=dev)
env = strive_b200.MapEnv(raster, dx, device=dev)
graph = bench.to_graph(sc, dev)
embed = {'map_feat': sc['map_feat'].to(dev), 'past_feat': sc['past_feat'].to(dev), 'prior_out': (sc['prior_mu'].to(dev), sc['prior_var'].to(dev))}
loop = RefineLoop(model, graph, sc['map_idx'].to(dev), env, embed, sc['z'].to(dev), bench.REFINE_W, 0.05, 20, veh_coll_buffer=0.2,
                  group_scene_ptr=list(range(0, 65, 4)))
ref = None
bad = 0
for i in range(6):
    loop._forward()
    torch.cuda.synchronize()
    t = loop.traj.clone()
    if ref is None:
        ref = t
    elif not torch.equal(ref, t):
        bad += 1
        d = (ref - t).abs().amax(dim=(0, 2))
        first = int((d > 0).nonzero()[0]) if bool((d > 0).any()) else -1
        print('  run %d differs: first differing step %d, max %.3e' % (i, first, float(d.max())))
print('STRIVE_PDL=%s: %d of 5 repeats differ' % (os.environ.get('STRIVE_PDL', 'default'), bad))
