"""Where does a sub-batch rollout first differ from the same scenes inside the full batch?  Prints max |diff| of every tape
tensor per step (bench world, BASELINE configs[1] scenes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import strive_b200
from strive_b200 import synth, shard, _cabi
from strive_b200.optim import RefineLoop
import bench
dev = torch.device('cuda:0')
FT, S, n = 6, 64, 32
raster, dx = synth.make_raster(seed=1, M=1, H=4096, W=4096)
model = strive_b200.make_model(nfuture=FT, state_dict=synth.make_weights(0), device=dev)
env = strive_b200.MapEnv(raster, dx, device=dev)
sc = synth.make_scenes(1000, [n] * S, map_extent_m=(200.0, 800.0), M=1, FT=FT, collide_frac=0.25, offroad_frac=0.25)
gptr = list(range(0, S + 1, 4))

def loop_for(s, gp):
    g = bench.to_graph(s, dev)
    embed = {'map_feat': s['map_feat'].to(dev), 'past_feat': s['past_feat'].to(dev), 'prior_out': (s['prior_mu'].to(dev), s['prior_var'].to(dev))}
    return RefineLoop(model, g, s['map_idx'].to(dev), env, embed, s['z'].to(dev), bench.REFINE_W, 0.05, FT, veh_coll_buffer=0.2, group_scene_ptr=gp)

def tape(loop, name, t, width):
    out = torch.empty((loop.NA, width), dtype=torch.float32, device=dev)
    _cabi.check(_cabi.lib().strive_decode_tape_read(_cabi.dptr(loop.tape), loop.NA, FT, name.encode(), t, _cabi.dptr(out), _cabi.stream_ptr()))
    return out.cpu()

full = loop_for(sc, gptr); full._forward()
for groups in ([0], [0, 1, 2], [11, 2, 7]):
    sub, lg, idx = shard.shard_scenes(sc, gptr, groups)
    lp = loop_for(sub, lg); lp._forward(); torch.cuda.synchronize()
    print('groups', groups, 'traj equal', torch.equal(lp.traj.cpu(), full.traj.cpu()[idx]))
    for t in range(FT):
        row = []
        for name, w in (('past_feat', 64), ('map_feat', 64), ('x', 64), ('P', 128), ('Q', 128), ('aggr', 64), ('pos', 4), ('loc', 4), ('mem', 192)):
            d = (tape(lp, name, t, w) - tape(full, name, t, w)[idx]).abs().max().item()
            row.append('%s %.1e' % (name, d))
        print('  t=%d ' % t + ' '.join(row))
