"""GPU: pins the tcgen05 / TMEM primitives (csrc/tc.cuh) the tensor-core map encoder is built from."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_tcgen05_selftest_gemm_and_shifted_window():
    from strive_b200 import _cabi
    L = _cabi.lib()
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(1)
    bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
    A, B = bf(torch.randn(128, 32, generator=g)), bf(torch.randn(32, 32, generator=g))
    X0, X1 = bf(torch.randn(144, 8, generator=g)), bf(torch.randn(144, 8, generator=g))
    D0 = torch.full((128, 32), float('nan'), device=dev)
    D1 = torch.full((128, 32), float('nan'), device=dev)
    d = [t.to(dev).contiguous() for t in (A, B, X0, X1)]
    _cabi.check(L.strive_tc_selftest(_cabi.dptr(d[0]), _cabi.dptr(d[1]), _cabi.dptr(d[2]), _cabi.dptr(d[3]), _cabi.dptr(D0), _cabi.dptr(D1),
                                     _cabi.stream_ptr()))
    torch.cuda.synchronize()
    ref0 = A.double() @ B.double().t()
    A1 = torch.cat([X0[1:129], X1[3:131]], dim=1)
    ref1 = A1.double() @ B[:, :16].double().t()
    e0 = (D0.cpu().double() - ref0).abs().max().item()
    e1 = (D1.cpu().double() - ref1).abs().max().item()
    msg = 'tcgen05 selftest: gemm err %.3e, shifted-window err %.3e' % (e0, e1)
    print(msg)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, 'diag_gpu.txt'), 'a') as f:
        f.write(msg + '\n')
        if not (e0 < 1e-3 and e1 < 1e-3):
            f.write('D0[:4,:8]=%s\nref0[:4,:8]=%s\nD1[:4,:8]=%s\nref1[:4,:8]=%s\n' % (D0[:4, :8].cpu(), ref0[:4, :8], D1[:4, :8].cpu(), ref1[:4, :8]))
    assert e0 < 1e-4 and e1 < 1e-4


def test_tcgen05_cta_pair_selftest():
    """cta_group::2: one M = 256 MMA over a cluster of two CTAs (each stages 128 rows of A and HALF of B), completion multicast to
    both CTAs' barriers, each CTA reads its 128 x N block of D from its own tensor memory."""
    from strive_b200 import _cabi
    L = _cabi.lib()
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(2)
    bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
    A = bf(torch.randn(256, 32, generator=g))
    msgs = []
    for n in (64, 128):
        B = bf(torch.randn(n, 32, generator=g))
        Ad, Bd = A.to(dev).contiguous(), B.to(dev).contiguous()          # named: the pointers must outlive the call
        D = torch.full((256, n), float('nan'), device=dev)
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        _cabi.check(L.strive_tc_selftest_pair(_cabi.dptr(Ad), _cabi.dptr(Bd), _cabi.dptr(D), _cabi.dptr(flag), n, _cabi.stream_ptr()))
        torch.cuda.synchronize()
        err = (D.cpu().double() - A.double() @ B.double().t()).abs().max().item()
        msgs.append('N=%d flag %d err %.3e' % (n, int(flag.item()), err))
        assert int(flag.item()) == 0 and err < 1e-4, msgs
    msg = 'tcgen05 CTA-pair selftest (cta_group::2, M = 256): ' + ', '.join(msgs)
    print(msg)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, 'diag_gpu.txt'), 'a') as f:
        f.write(msg + '\n')


def test_conv3_on_cta_pairs_matches_the_single_cta_kernel():
    """strive_mapenc_set_pair(1): conv3 of the map encoder on clusters of two CTAs (cta_group::2 MMAs, B operands split across the
    pair, cross-CTA mbarrier signalling).  Same arithmetic except that the small hi x lo terms are summed apart from the hi x hi
    products: features agree to fp32 re-association level, the pair kernel is run-to-run bitwise reproducible and bitwise invariant
    to the batch composition, and crop counts that leave clusters without work (n = 1, 3) are handled."""
    import strive_b200
    from strive_b200 import _cabi, synth
    from tests.test_gpu_parity import ctx, diag
    dev, model, env = ctx()
    L = _cabi.lib()
    g = torch.Generator().manual_seed(7)
    N = 333
    xy = torch.rand(N, 2, generator=g) * 600 + 200
    ang = torch.rand(N, generator=g) * 6.2831853
    pose = torch.cat([xy, torch.cos(ang)[:, None], torch.sin(ang)[:, None]], 1).to(dev).contiguous()
    mapix = torch.zeros(N, dtype=torch.int32, device=dev)
    try:
        L.strive_mapenc_set_pair(0)
        f_single = model.encode_map_poses(pose, mapix, env).clone()
        L.strive_mapenc_set_pair(1)
        f_pair = model.encode_map_poses(pose, mapix, env).clone()
        f_pair2 = model.encode_map_poses(pose, mapix, env).clone()
        sub = [model.encode_map_poses(pose[:n].contiguous(), mapix[:n].contiguous(), env).clone() for n in (1, 3, 150)]
        torch.cuda.synchronize()
    finally:
        L.strive_mapenc_set_pair(1)          # the default
    d = (f_pair - f_single).abs().max().item()
    diag('conv3 on CTA pairs vs single-CTA kernel: max |feature diff| %.2e on %d crops (max |feature| %.2f)' % (d, N, f_single.abs().max().item()))
    assert torch.equal(f_pair, f_pair2)
    for s_ in sub:
        assert torch.equal(s_, f_pair[:s_.size(0)])           # batch invariance: a crop's features do not depend on the batch around it
    assert d < 5e-5
