"""CPU: the C-ABI library loads and exports every symbol include/strive_b200.h declares, binding layouts match,
weight packing matches the library's segment table, host-side exact helpers match torch."""
import os
import re

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    import __graft_entry__ as ge
    ge.build()


def test_library_exports_every_declared_symbol():
    _build()
    from strive_b200 import _cabi
    L = _cabi.lib()      # also verifies struct layouts against the library
    hdr = open(os.path.join(ROOT, 'include', 'strive_b200.h')).read()
    names = sorted(set(re.findall(r'\b(strive_[a-z0-9_]+)\s*\(', hdr)))
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), 'missing export %s' % n
    assert sorted(_cabi.EXPORTS) == names


def test_weight_packing_matches_library_layout():
    _build()
    import ctypes as C
    from strive_b200 import _cabi, synth, weights
    sd = synth.make_weights(0)
    blob, sizes = weights.pack_decode_weights(sd, 2)
    want = (C.c_int64 * 256)()
    n = C.c_int(0)
    _cabi.check(_cabi.lib().strive_model_layout(2, want, 256, C.byref(n)))
    assert list(want[:n.value]) == sizes
    # a transposed segment really is the transpose of the reference tensor
    segs = dict(weights.decode_segments(sd, 2))
    assert torch.equal(segs['E3_T'], sd['decoder_net.msg.0.edge_mlp.net.3.weight'].t())
    assert torch.equal(segs['CW1'][(3 * 5 + 2) * 5 + 4], sd['map_conv.3.weight'][:, 3, 2, 4])
    assert torch.equal(segs['FCW'][5 * 4 + 1 * 2 + 1], sd['map_feature.weight'][:, 5 * 4 + 3])


def test_product_path_refuses_cpu_tensors():
    _build()
    import pytest
    from strive_b200 import _cabi
    with pytest.raises(RuntimeError):
        _cabi.dptr(torch.zeros(4))


def test_linspace5_matches_torch_linspace_bitwise():
    from strive_b200.losses import linspace5
    g = torch.Generator().manual_seed(0)
    lo = -(torch.rand(200, generator=g) * 4 + 1)
    hi = -lo + torch.rand(200, generator=g) * 0.1
    mine = linspace5(lo, hi)
    ref = torch.stack([torch.linspace(lo[i].item(), hi[i].item(), 5) for i in range(200)])
    assert torch.equal(mine, ref)


def test_oracle_is_not_imported_by_the_product():
    pkg = os.path.join(ROOT, 'strive_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert 'oracle' not in src.replace('the oracle', '').replace('CPU oracle', ''), fn


def test_tape_lease_recycles_buffers_and_bounds_the_free_list():
    """The drop-in decode leases its multi-GB rollout tape from a free list keyed by (device, size): two alternating graphs
    need two buffers, however many iterations run; at most two idle buffers are kept per size."""
    import gc
    import torch
    from strive_b200.traffic_model import _TapeLease
    _TapeLease.release_all()
    made0 = _TapeLease.made
    dev = torch.device('cpu')
    prev = None
    ptrs = set()
    for _ in range(10):
        cur = _TapeLease(4096, dev)          # the new forward runs before the previous graph is dropped
        ptrs.add(cur.buf.data_ptr())
        prev = cur
    assert _TapeLease.made - made0 == 2 and len(ptrs) == 2
    held = [_TapeLease(4096, dev) for _ in range(5)]
    del held, prev, cur
    gc.collect()
    assert len(_TapeLease._free[('cpu', 4096)]) == 2       # the rest went back to the allocator
    other = _TapeLease(8192, dev)
    assert other.buf.numel() == 8192
    _TapeLease.release_all()


def test_conv3_cta_pair_weight_pack_is_the_split_of_the_single_cta_pack():
    """pack_tc_weights: the last blob segment holds conv3 for the CTA-pair kernel -- per CTA rank, K chunk and tap ONE 64-row block
    [W_hi rows 32r..32r+31 ; W_lo rows 32r..32r+31] in the same K-major layout -- i.e. exactly the rows of the single-CTA pack
    ([khalf][prec][64 rows][8 k]) regrouped; sizes agree with the library's blob layout."""
    from strive_b200 import _cabi, synth, weights
    blob = weights.pack_tc_weights(synth.make_weights(0))
    sizes = [7 * 2 * 48 * 16 + 64, 51200, 204800, 2 * (4 * 9 * 2 * 1024), 9 * 2 * 128 * 64 * 2, 18 * 2 * 128 * 64 * 2, 8 * 2 * 64 * 64 * 2,
             2 * 2 * 25 * 2048]
    assert blob.numel() == sum(sizes) == _cabi.lib().strive_model_tc_bytes()
    off = [0]
    for v in sizes:
        off.append(off[-1] + v)
    std = blob[off[2]:off[3]].view(torch.int16).reshape(2, 25, 2, 2, 64, 8)          # (c2, tap, khalf, prec, row, k)
    pair = blob[off[7]:off[8]].view(torch.int16).reshape(2, 2, 25, 2, 64, 8)         # (rank, c2, tap, khalf, row, k)
    for rank in (0, 1):
        assert torch.equal(pair[rank][:, :, :, :32], std[:, :, :, 0, 32 * rank:32 * rank + 32])      # W_hi rows of this rank
        assert torch.equal(pair[rank][:, :, :, 32:], std[:, :, :, 1, 32 * rank:32 * rank + 32])      # W_lo rows of this rank
