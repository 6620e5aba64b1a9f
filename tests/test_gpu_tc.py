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
