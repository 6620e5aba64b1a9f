"""Per-kernel CUDA-event times of one refine iteration of the bench workload (BASELINE configs[1]) through KPROF.
STRIVE_LIB=<path> selects an experimental build of the library (development only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from strive_b200 import _cabi
if os.environ.get('STRIVE_LIB'):
    _cabi.LIB_PATH = os.environ['STRIVE_LIB']
import strive_b200
if os.environ.get('EDGE_IMPL'):
    _cabi.set_edge_impl(int(os.environ['EDGE_IMPL']) != 0)
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
for _ in range(2):
    loop.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    loop.step()
e1.record()
torch.cuda.synchronize()
print('%s: %.2f ms/step, loss %.4f' % (os.environ.get('STRIVE_LIB', 'default'), e0.elapsed_time(e1) / 3, float(loop.terms[:, 0].sum())))
_cabi.profile_enable(True)
loop.eager_step()
torch.cuda.synchronize()
rep = _cabi.profile_report()
tot = sum(v[1] for v in rep.values())
print('  ' + '  '.join('%s %.2f' % (k, v[1]) for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1])[:18]) + '  | total %.2f ms' % tot)
print('  per launch (us): ' + '  '.join('%s %.1f' % (k, 1000 * rep[k][1] / rep[k][0]) for k in
                                        ('node_fwd', 'edge_fwd', 'post_fwd', 'gru_fwd', 'gru_bwd', 'post_bwd', 'edge_bwd', 'node_bwd') if k in rep))
