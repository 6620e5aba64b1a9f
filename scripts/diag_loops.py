"""GPU diagnostic for the adv / sol loops: first-iteration gradients per row vs the oracle's literal two-decode loops."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import strive_b200
from strive_b200.optim import collate_tgt_other_z
from strive_b200.losses import AdvGenLoss, TgtMatchingLoss, AvoidCollLoss
from oracle import strive_oracle as O
from tests.common import world, golden, scene_for, ADV_W, SOL_W
from tests.test_gpu_parity import to_graph

dev = torch.device('cuda:0')
raster, dx, sd = world()
model = strive_b200.make_model(nfuture=20, state_dict=sd, device=dev)
env = strive_b200.MapEnv(raster, dx, device=dev)
g = golden('losses')
FT = 5
sc = scene_for(g, FT=FT)
NA = sc['z'].size(0)
ego = torch.zeros(NA, dtype=torch.bool); ego[sc['ptr'][:-1]] = True
pf = sc['ext_future'][:, :FT].contiguous()
rec = []
O.adv_loop(sd, sc, raster, dx, ADV_W, 1, 0.05, FT, pf, crash_min_t=1, crash_min_infront=-0.5, veh_coll_buffer=0.1, record=rec)
graph = to_graph(sc, dev)
embed = {'map_feat': sc['map_feat'].to(dev), 'past_feat': sc['past_feat'].to(dev)}
nrm = model.get_normalizer()
egod = ego.to(dev)
tgt_z = sc['z'][ego].to(dev).requires_grad_(True); other_z = sc['z'][~ego].to(dev).requires_grad_(True)
z_all = collate_tgt_other_z(graph, tgt_z.detach(), other_z.detach()).requires_grad_(True)
print('collate ok', (z_all.detach().cpu() - sc['z']).abs().max().item())
fut = model.decode_embedding(z_all, embed, graph, sc['map_idx'].to(dev), env, ext_future=pf.to(dev), nfuture=FT)['future_pred']
with torch.no_grad():
    fo = O.decode(sd, sc['z'], sc['map_feat'], sc['past_feat'], sc['past'][:, -1, :], sc['lw'], sc['sem'], sc['ptr'], sc['edge_index'], sc['map_idx'], raster, dx, FT, ext_future=pf)
print('traj diff', (fut.detach().cpu() - fo).abs().max().item())
fu = nrm.unnormalize(fut)
pun = nrm.unnormalize(pf.to(dev))
lt = TgtMatchingLoss(ADV_W)(fu[egod], pun, tgt_z, (sc['prior_mu'][ego].to(dev), sc['prior_var'][ego].to(dev)))
adv = AdvGenLoss(ADV_W, model.get_att_normalizer().unnormalize(graph.lw), sc['map_idx'].to(dev)[graph.batch], env, other_z.clone().detach(), graph.ptr,
                 veh_coll_buffer=0.1, crash_loss_min_time=1, crash_loss_min_infront=-0.5)
la = adv(fu, pun, other_z, (sc['prior_mu'][~ego].to(dev), sc['prior_var'][~ego].to(dev)))
g_t = torch.autograd.grad(lt['loss'], z_all, retain_graph=True)[0]
g_a, g_o = torch.autograd.grad(la['loss'], [z_all, other_z])
gt = g_t[egod].cpu(); go = (g_a[~egod] + g_o).cpu()
print('loss', float(lt['loss'] + la['loss']), rec[0]['loss'])
print('g_tgt  : max ref %.3e  per-row err' % rec[0]['g_tgt'].abs().max().item(), (gt - rec[0]['g_tgt']).abs().amax(dim=1).numpy())
print('g_other: max ref %.3e  per-row err' % rec[0]['g_other'].abs().max().item(), (go - rec[0]['g_other']).abs().amax(dim=1).numpy())
print('g_t on non-ego rows (should be ignored):', g_t[~egod].abs().max().item(), ' g_a on ego rows:', g_a[egod].abs().max().item())
# oracle split: BPTT-only part of other grads
