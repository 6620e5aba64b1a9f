"""One short refine iteration (FT small) of the bench-sized batch: a cheap target for `ncu -k regex:<kernel>` captures of the
rollout kernels.  FT=3 by default: two GRU steps, three node/edge/post steps forward and backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import strive_b200
import bench
from strive_b200.optim import RefineLoop
FT = int(os.environ.get('FT', '3'))
dev = torch.device('cuda:0')
bench.WORK['FT'] = FT
raster, dx, sd, sc = bench.make_workload(0)
model = strive_b200.make_model(nfuture=FT, state_dict=sd, device=dev)
env = strive_b200.MapEnv(raster, dx, device=dev)
graph = bench.to_graph(sc, dev)
embed = {'map_feat': sc['map_feat'].to(dev), 'past_feat': sc['past_feat'].to(dev), 'prior_out': (sc['prior_mu'].to(dev), sc['prior_var'].to(dev))}
loop = RefineLoop(model, graph, sc['map_idx'].to(dev), env, embed, sc['z'].to(dev), bench.REFINE_W, 0.05, FT, veh_coll_buffer=0.2,
                  group_scene_ptr=list(range(0, 65, 4)))
for _ in range(int(os.environ.get('ITERS', '2'))):
    loop.step()
torch.cuda.synchronize()
print('loss %.4f' % float(loop.terms[:, 0].sum()))
