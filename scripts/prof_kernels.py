"""Per-kernel CUDA-event times of the map encoder alone (N crops) through the library's KPROF hooks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import strive_b200
from strive_b200 import synth, _cabi
N = int(os.environ.get('N', '2048'))
dev = torch.device('cuda:0')
raster, dx = synth.make_raster(seed=1, M=1, H=4096, W=4096)
model = strive_b200.make_model(nfuture=20, state_dict=synth.make_weights(0), device=dev)
env = strive_b200.MapEnv(raster, dx, device=dev)
g = torch.Generator().manual_seed(0)
xy = torch.rand(N, 2, generator=g) * 600 + 200
ang = torch.rand(N, generator=g) * 6.2831853
pose = torch.cat([xy, torch.cos(ang)[:, None], torch.sin(ang)[:, None]], 1).to(dev).contiguous()
mapix = torch.zeros(N, dtype=torch.int32, device=dev)
for _ in range(2):
    f = model.encode_map_poses(pose, mapix, env)
torch.cuda.synchronize()
_cabi.profile_enable(True)
for _ in range(3):
    f = model.encode_map_poses(pose, mapix, env)
torch.cuda.synchronize()
tot = 0.0
for k, v in sorted(_cabi.profile_report().items(), key=lambda kv: -kv[1][1]):
    print('%-12s n=%d avg %.1f us' % (k, v[0], 1000 * v[1] / v[0]))
    tot += 1000 * v[1] / v[0]
print('sum %.1f us per %d crops; feat checksum %.6f' % (tot, N, float(f.double().sum())))
