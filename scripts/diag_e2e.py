"""Per-phase host times and device time of the drop-in API step (the e2e arm of bench.py) over many steps: which phase
stalls when a step is slow?"""
import os, sys, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import strive_b200
import bench
from strive_b200.losses import AvoidCollLoss
dev = torch.device('cuda:0')
raster, dx, sd, sc = bench.make_workload(0)
model = strive_b200.make_model(nfuture=20, state_dict=sd, device=dev)
env = strive_b200.MapEnv(raster, dx, device=dev)
graph = bench.to_graph(sc, dev)
midx = sc['map_idx'].to(dev)
gptr = list(range(0, 65, 4))
host = {k: sc[k].clone().pin_memory() for k in ('z', 'map_feat', 'past_feat', 'prior_mu', 'prior_var')}
devb = {k: torch.empty_like(v, device=dev) for k, v in host.items()}
z_dev = devb['z'].requires_grad_(True)
opt = torch.optim.Adam([z_dev], lr=0.05)
lossm = AvoidCollLoss(bench.REFINE_W, model.get_att_normalizer().unnormalize(graph.lw), midx[graph.batch], env, sc['z'].to(dev),
                      veh_coll_buffer=0.2, group_scene_ptr=gptr, ptr_for_groups=sc['ptr'])
loss_host = torch.zeros((len(gptr) - 1, 16)).pin_memory()
z_host_out = torch.zeros_like(host['z']).pin_memory()
rows = []
N = int(os.environ.get('N', '40'))
if os.environ.get('NOGC'):
    gc.disable()
for it in range(N):
    t = [time.perf_counter()]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with torch.no_grad():
        for k, v in host.items():
            devb[k].copy_(v, non_blocking=True)
    em = {'map_feat': devb['map_feat'], 'past_feat': devb['past_feat'], 'prior_out': (devb['prior_mu'], devb['prior_var'])}
    opt.zero_grad()
    fut = model.get_normalizer().unnormalize(model.decode_embedding(z_dev, em, graph, midx, env, nfuture=20)['future_pred'])
    t.append(time.perf_counter())
    ld = lossm(fut, z_dev, em['prior_out'])
    t.append(time.perf_counter())
    ld['loss'].backward()
    t.append(time.perf_counter())
    opt.step()
    loss_host.copy_(lossm.last_terms.detach(), non_blocking=True)
    z_host_out.copy_(z_dev.detach(), non_blocking=True)
    e1.record()
    t.append(time.perf_counter())
    torch.cuda.synchronize()
    t.append(time.perf_counter())
    host['z'].copy_(z_host_out)
    t.append(time.perf_counter())
    rows.append([1000 * (t[i + 1] - t[i]) for i in range(6)] + [e0.elapsed_time(e1)])
    if os.environ.get('DELREF'):
        del fut, ld, em
    if os.environ.get('GCEACH'):
        gc.collect()
names = ['h2d+fwd_call', 'loss_call', 'bwd_call', 'adam+d2h', 'sync', 'hostcopy', 'gpu_ms']
tot = [sum(r[:6]) for r in rows]
med = sorted(tot)[len(tot) // 2]
print('median step %.1f ms; slow steps (> 1.3x median):' % med)
for i, r in enumerate(rows):
    if tot[i] > 1.3 * med and i >= 3:
        print('  step %2d total %.1f | ' % (i, tot[i]) + ' '.join('%s %.1f' % (n, v) for n, v in zip(names, r)))
print('typical: ' + ' '.join('%s %.1f' % (n, v) for n, v in zip(names, rows[len(rows) // 2])))
from strive_b200.traffic_model import _TapeLease
print('tapes made %d, free lists %s' % (_TapeLease.made, {k: len(v) for k, v in _TapeLease._free.items()}))
print('alloc: reserved %.1f GB, num_alloc_retries %d, cudaMalloc segments %d' % (torch.cuda.memory_reserved() / 1e9, torch.cuda.memory_stats()['num_alloc_retries'],
                                                                                torch.cuda.memory_stats()['segment.all.allocated']))
