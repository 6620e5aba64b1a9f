"""Runs the map encoder alone on N random poses (profiling target for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import strive_b200
from strive_b200 import synth
N = int(os.environ.get('N', '512'))
REPS = int(os.environ.get('REPS', '3'))
dev = torch.device('cuda:0')
raster, dx = synth.make_raster(seed=1, M=1, H=4096, W=4096)
sd = synth.make_weights(0)
model = strive_b200.make_model(nfuture=20, state_dict=sd, device=dev)
env = strive_b200.MapEnv(raster, dx, device=dev)
g = torch.Generator().manual_seed(0)
xy = torch.rand(N, 2, generator=g) * 600 + 200
ang = torch.rand(N, generator=g) * 6.2831853
pose = torch.cat([xy, torch.cos(ang)[:, None], torch.sin(ang)[:, None]], 1).to(dev).contiguous()
mapix = torch.zeros(N, dtype=torch.int32, device=dev)
for _ in range(REPS):
    f = model.encode_map_poses(pose, mapix, env)
torch.cuda.synchronize()
from strive_b200 import _cabi
_cabi.lib().strive_tc_debug(int(os.environ.get('DBG', '0')))
_cabi.tc_trace(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(REPS):
    f = model.encode_map_poses(pose, mapix, env)
e1.record()
torch.cuda.synchronize()
print('mapenc: %d crops, %.3f ms per call, %.3f us per crop, feat checksum %.6f' % (N, e0.elapsed_time(e1) / REPS, 1000 * e0.elapsed_time(e1) / REPS / N, float(f.double().sum())))
tr = _cabi.tc_trace(True)
for k, name in enumerate(['conv1', 'conv2', 'conv3', 'conv4']):
    t = tr[k]
    ct = max(t[7], 1)
    print('%s per CTA-launch kcycles: producer wait-empty %.0f / total %.0f | mma wait-full %.0f wait-acc %.0f / total %.0f | epi wait %.0f / total %.0f (CTAs %d)' % (
        name, t[0] / ct / 1e3, t[1] / ct / 1e3, t[2] / ct / 1e3, t[3] / ct / 1e3, t[4] / ct / 1e3, t[5] / ct / 1e3, t[6] / ct / 1e3, t[7]))
