"""Map encoder alone on N crops under each timing-experiment flag set of strive_tc_debug (development only): per-kernel CUDA-event
times via KPROF.  Flags: 1 no output stores, 2 no operand stores, 4 no input loads, 8 no MMAs,
32 no weight copies (tc_gemm).  Results with any flag set are garbage by design."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hashlib
import torch
from strive_b200 import _cabi
if os.environ.get('STRIVE_LIB'):            # A/B against another build of the library (development only)
    _cabi.LIB_PATH = os.environ['STRIVE_LIB']
import strive_b200
from strive_b200 import synth
N = int(os.environ.get('N', '2048'))
FLAGS = [int(v) for v in os.environ.get('FLAGS', '0,4,2,8,32,12,14,6').split(',')]
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
for _ in range(2):
    f = model.encode_map_poses(pose, mapix, env)
torch.cuda.synchronize()
ref = f.clone()
if os.environ.get('STRIVE_MAPENC_PAIR', '0') != '0':      # A/B of the CTA-pair kernels against the single-CTA kernels, same process
    mask = int(os.environ['STRIVE_MAPENC_PAIR'])
    _cabi.lib().strive_mapenc_set_pair(0)
    f1 = model.encode_map_poses(pose, mapix, env)
    _cabi.lib().strive_mapenc_set_pair(mask)
    f2 = model.encode_map_poses(pose, mapix, env)
    torch.cuda.synchronize()
    print('pair kernels (mask %d) vs single-CTA kernels: max |feature diff| %.3e (max |feature| %.3f); pair kernels run to run %.1e' % (
        mask, float((f1 - ref).abs().max()), float(ref.abs().max()), float((f2 - ref).abs().max())))
print('%s: feature sha1 %s' % (os.environ.get('STRIVE_LIB', 'in-tree library'), hashlib.sha1(ref.cpu().numpy().tobytes()).hexdigest()))
names = ('crop_pack', 'tc_conv1', 'tc_conv2', 'tc_conv3', 'tc_conv4', 'tc_conv5', 'tc_conv6', 'tc_fc')
for fl in FLAGS:
    _cabi.lib().strive_tc_debug(fl)
    f = model.encode_map_poses(pose, mapix, env)
    torch.cuda.synchronize()
    _cabi.profile_enable(True)
    for _ in range(3):
        f = model.encode_map_poses(pose, mapix, env)
    torch.cuda.synchronize()
    rep = _cabi.profile_report()
    _cabi.profile_enable(False)
    print('flags %2d: ' % fl + '  '.join('%s %.1f' % (k.replace('tc_', ''), 1000 * rep[k][1] / rep[k][0]) for k in names if k in rep)
          + '  | us per launch; |feat - flags0| = %.3g' % float((f - ref).abs().max()))
_cabi.lib().strive_tc_debug(0)
if os.environ.get('TRACE'):
    _cabi.tc_trace(True)
    f = model.encode_map_poses(pose, mapix, env)
    torch.cuda.synchronize()
    tr = _cabi.tc_trace(True)
    t = tr[2]
    ct = max(t[7], 1)
    print('conv3 trace (kcycles per traced CTA): producer wait-empty %.0f / total %.0f | mma wait-full %.0f wait-acc %.0f wait-weights(pair kernel)/epi-wait %.0f / total %.0f | CTAs %d' % (
        t[0] / ct / 1e3, t[1] / ct / 1e3, t[2] / ct / 1e3, t[3] / ct / 1e3, t[5] / ct / 1e3, t[4] / ct / 1e3, t[7]))
