"""GPU diagnostic: isolates loss-kernel vs BPTT-kernel error in one refine iteration against the fp32 oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import strive_b200
from strive_b200 import synth
from strive_b200.optim import RefineLoop
from oracle import strive_oracle as O
from tests.common import world, golden, scene_for, REFINE_W
from tests.test_gpu_parity import to_graph

dev = torch.device('cuda:0')
raster, dx, sd = world()
g = golden('refine')
sc = scene_for(g)
FT, lr = int(g['FT']), float(g['lr'])
model = strive_b200.make_model(nfuture=20, state_dict=sd, device=dev)
env = strive_b200.MapEnv(raster, dx, device=dev)
graph = to_graph(sc, dev)
embed = {'map_feat': sc['map_feat'].to(dev), 'past_feat': sc['past_feat'].to(dev), 'prior_out': (sc['prior_mu'].to(dev), sc['prior_var'].to(dev))}
loop = RefineLoop(model, graph, sc['map_idx'].to(dev), env, embed, sc['z'].to(dev), REFINE_W, lr, FT, veh_coll_buffer=0.2)
loop._forward(); loop._loss(); loop._backward(); torch.cuda.synchronize()
traj_g = loop.traj.cpu(); dtraj_g = loop.d_traj.cpu(); dzb_g = loop.d_z_bptt.cpu(); dzd_g = loop.d_z_direct.cpu()
print('traj gpu vs golden traj0', (traj_g - torch.from_numpy(g['traj0'])).abs().max().item())
# (b) oracle loss on the GPU trajectory
lw_un = O.unnorm_att(sc['lw']); mapixes = sc['map_idx'][sc['batch']]
tn = traj_g.clone().requires_grad_(True)
z = sc['z'].clone().requires_grad_(True)
ld = O.avoid_coll_loss(O.unnorm_state(tn), z, (sc['prior_mu'], sc['prior_var']), sc['z'].clone(), REFINE_W, lw_un, mapixes, None, raster, dx, veh_coll_buffer=0.2)
ld['loss'].backward()
print('loss gpu %.6f oracle(on gpu traj) %.6f' % (float(loop.terms[:, 0].sum()), float(ld['loss'])), 'terms', loop.terms[0, :8].cpu().numpy())
e = (dtraj_g - tn.grad).abs()
print('d_traj(normalised) gpu vs oracle-on-gpu-traj: max err %.3e (max %.3e)' % (e.max().item(), tn.grad.abs().max().item()))
idx = torch.nonzero(e > 1e-3 * tn.grad.abs().max())
print('  mismatching entries', idx[:20].tolist())
print('d_z_direct err %.3e' % (dzd_g - z.grad).abs().max().item())
# (c) oracle BPTT with the GPU's d_traj as the seed
z2 = sc['z'].clone().requires_grad_(True)
tr = O.decode(sd, z2, sc['map_feat'], sc['past_feat'], sc['past'][:, -1, :], sc['lw'], sc['sem'], sc['ptr'], sc['edge_index'], sc['map_idx'], raster, dx, FT)
tr.backward(dtraj_g)
e2 = (dzb_g - z2.grad).abs()
print('d_z_bptt gpu vs oracle-autograd(seed = gpu d_traj): max err %.3e (max %.3e)' % (e2.max().item(), z2.grad.abs().max().item()))
print('  per-agent max err', e2.amax(dim=1).numpy().round(5))
print('  traj gpu vs oracle32', (traj_g - tr.detach()).abs().amax(dim=(0, 2)).numpy())
# per-step seeds: which rollout step's adjoint disagrees
for t in range(FT):
    z3 = sc['z'].clone().requires_grad_(True)
    tr3 = O.decode(sd, z3, sc['map_feat'], sc['past_feat'], sc['past'][:, -1, :], sc['lw'], sc['sem'], sc['ptr'], sc['edge_index'], sc['map_idx'], raster, dx, FT)
    seed = torch.zeros_like(tr3); seed[:, t] = dtraj_g[:, t]
    tr3.backward(seed)
    zz = sc['z'].clone().to(dev).requires_grad_(True)
    out = model.decode_embedding(zz, embed, graph, sc['map_idx'].to(dev), env, nfuture=FT)['future_pred']
    out.backward(seed.to(dev))
    e3 = (zz.grad.cpu() - z3.grad).abs()
    print('  seed only at step %d: max err %.3e (max %.3e) worst agent %d' % (t, e3.max().item(), z3.grad.abs().max().item(), int(e3.amax(dim=1).argmax())))
