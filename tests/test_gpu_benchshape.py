"""GPU parity at the shapes of the MEASURED configurations (round-1 verdict: every GPU-vs-oracle comparison ran on scenes of
<= 8 agents): 32-agent scenes and the 128-agent collision block of BASELINE configs[1], 64-agent scenes (configs[4]), ragged
17/33/40 (configs[2]).  These run the multi-tile path of the edge kernels (ntiles = ceil((n-1)/16) >= 2: running max / arg-max
carried across 16-edge tiles) and full 128-row collision blocks, against the CPU oracle AND against fixtures written by the
unmodified reference (tests/golden/bench_shape.npz, oracle/gen_golden_r2.py).

Tolerances are the ones of the small fixtures (tests/test_gpu_parity.py): 2e-5 on every first-step tensor, teacher-forced
trajectory 2e-5*(1+FT/2), teacher-forced dL/dz 2e-4*max(1,|g|max); integer work (arg-max routing, collision counts) exact
except where the oracle's own margin is at rounding level.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import strive_oracle as O
from strive_b200 import synth
from tests.common import world, golden, EXTENT, REFINE_W
from tests.test_gpu_parity import ctx, diag, to_graph, _tape, _lowlevel

pytestmark = pytest.mark.gpu

SHAPES = {'n32': (81, [32]), 'n64': (82, [64]), 'ragged': (83, [17, 33, 40]), 'n2x32': (84, [32, 32])}


def scene(tag, FT):
    seed, sizes = SHAPES[tag]
    return synth.make_scenes(seed, sizes, map_extent_m=EXTENT, M=2, FT=FT, collide_frac=1.0, offroad_frac=1.0)


def gpu_forward(sc, FT):
    from strive_b200 import _cabi
    dev, model, env = ctx()
    graph = to_graph(sc, dev)
    sb = model.scene_batch(graph, sc['map_idx'].to(dev))
    L = _cabi.lib()
    NA = sb.NA
    nb = L.strive_decode_tape_bytes(NA, FT)
    tape = torch.empty(nb, dtype=torch.uint8, device=dev)
    traj = torch.empty((NA, FT, 4), dtype=torch.float32, device=dev)
    z = sc['z'].to(dev).contiguous()
    mf, pf = sc['map_feat'].to(dev).contiguous(), sc['past_feat'].to(dev).contiguous()
    _cabi.check(L.strive_decode_fwd(model.device_model().handle, C.byref(sb.cstruct), C.byref(env.cstruct), _cabi.dptr(z), _cabi.dptr(mf),
                                    _cabi.dptr(pf), None, FT, _cabi.dptr(traj), _cabi.dptr(tape), nb, _cabi.stream_ptr()))
    torch.cuda.synchronize()
    return traj.cpu(), tape, NA


def read_arg(tape, t, NA, FT):
    from strive_b200 import _cabi
    out = torch.empty((NA, 64), dtype=torch.uint8, device=tape.device)
    _cabi.check(_cabi.lib().strive_decode_tape_read(_cabi.dptr(tape), NA, FT, b'arg', t, _cabi.dptr(out), _cabi.stream_ptr()))
    return out.cpu()


def oracle_decode(sc, FT, taps=None, z=None, override=None):
    raster, dx, sd = world()
    return O.decode(sd, sc['z'] if z is None else z, sc['map_feat'], sc['past_feat'], sc['past'][:, -1, :], sc['lw'], sc['sem'], sc['ptr'],
                    sc['edge_index'], sc['map_idx'], raster, dx, FT, taps=taps, map_feat_override=override)


@pytest.mark.parametrize('tag,FT', [('n32', 5), ('n64', 4), ('ragged', 4)])
def test_per_step_tape_tensors_and_argmax_routing(tag, FT):
    """Every tape tensor of every step + the arg-max routing table against the oracle's intermediates.  Step 0 has identical
    inputs (strict 2e-5); later steps are compared with the oracle TEACHER-FORCED on the GPU's map features (the one
    discontinuous input), so they stay strict as well."""
    sc = scene(tag, FT)
    traj, tape, NA = gpu_forward(sc, FT)
    feats = [_tape(tape, 'map_feat', t, NA, FT, 64) for t in range(1, FT)]
    taps = {}
    with torch.no_grad():
        ref = oracle_decode(sc, FT, taps=taps, override=feats)
    ptr = sc['ptr']
    scene_of = sc['batch']
    base = ptr[scene_of]                                        # first agent of each agent's scene
    worst = 0.0
    flips = 0
    for t in range(FT):
        st = taps['steps'][t]
        row = []
        for name, width, refv in (('past_feat', 64, st['past_feat']), ('x', 64, st['x']), ('aggr', 64, st['aggr']), ('pos', 4, st['pos_in']),
                                  ('loc', 4, st['loc']), ('mem', 192, st['mem'].permute(1, 0, 2).reshape(NA, 192))):
            got = _tape(tape, name, t, NA, FT, width)
            d = (got - refv).abs().max().item()
            row.append('%s=%.1e' % (name, d))
            worst = max(worst, d)
        dtraj = (traj[:, t] - ref[:, t]).abs().max().item()
        worst = max(worst, dtraj)
        # arg-max routing: local source index per (agent, channel); 255 = no in-edges
        arg = read_arg(tape, t, NA, FT).long()
        ref_arg = st['arg'] - base.view(-1, 1)
        neq = arg != ref_arg
        n_neq = int(neq.sum())
        if n_neq:
            # a differing winner is legitimate only where the oracle's own margin is at fp32 rounding level
            assert float(st['arg_margin'][neq].max()) < 1e-5, 'arg-max routing differs where the margin is %.3e' % float(st['arg_margin'][neq].max())
            # ... and the GPU's pick must be the oracle's runner-up value-wise
            gi = (arg + base.view(-1, 1))[neq]
            ii, cc = torch.nonzero(neq, as_tuple=True)
            assert float((st['msg_dense'][ii, gi, cc] - st['aggr'][ii, cc]).abs().max()) < 1e-5
        flips += n_neq
        assert int(arg.max()) < int((ptr[1:] - ptr[:-1]).max())
        diag('bench-shape %s step %d: traj=%.1e %s arg-max mismatches %d of %d' % (tag, t, dtraj, ' '.join(row), n_neq, arg.numel()))
    assert worst < 2e-5 * (1.0 + 0.5 * FT), worst
    assert flips <= 3


def test_teacher_forced_rollout_and_adjoint_n32_ft20():
    """BASELINE configs[1] scene shape over the full horizon: 32 agents x 20 steps, forward and BPTT.  The max aggregation makes
    dL/dz a function of 41 k arg-max routings; where the oracle's own winner leads by less than fp32 rounding the GPU may pick the
    runner-up (checked: only there, and only a value-wise tie), and the two then follow different -- equally valid -- subgradients."""
    FT = 20
    sc = scene('n32', FT)
    traj, feats, bwd = _lowlevel(sc, FT)
    z = sc['z'].clone().requires_grad_(True)
    taps = {}
    ref = oracle_decode(sc, FT, taps=taps, z=z, override=feats)
    e_t = (traj - ref.detach()).abs().amax(dim=(0, 2))
    base = sc['ptr'][sc['batch']]
    flips = 0
    for t in range(FT):
        st = taps['steps'][t]
        arg = read_arg(bwd.tape, t, bwd.NA, FT).long()
        neq = arg != (st['arg'] - base.view(-1, 1))
        if int(neq.sum()):
            assert float(st['arg_margin'][neq].max()) < 1e-5, 'arg-max routing differs where the margin is %.3e' % float(st['arg_margin'][neq].max())
            flips += int(neq.sum())
    seed = torch.randn(traj.shape, generator=torch.Generator().manual_seed(7))
    ref.backward(seed)
    got = bwd(seed)
    scale = z.grad.abs().max().item()
    e_g = (got - z.grad).abs().max().item()
    diag('bench-shape teacher-forced n32 FT=20: traj err per step %s | grad err %.3e (max %.3e) | arg-max ties resolved differently: %d of %d' % (
        ' '.join('%.1e' % v for v in e_t.tolist()), e_g, scale, flips, FT * bwd.NA * 64))
    assert e_t.max().item() < 2e-5 * (1.0 + 0.5 * FT)
    assert flips <= 3
    assert e_g < (2e-4 if flips == 0 else 5e-3) * max(1.0, scale)


@pytest.mark.parametrize('tag,FT', [('n64', 6), ('ragged', 6)])
def test_teacher_forced_rollout_and_adjoint_big_scenes(tag, FT):
    sc = scene(tag, FT)
    traj, feats, bwd = _lowlevel(sc, FT)
    z = sc['z'].clone().requires_grad_(True)
    ref = oracle_decode(sc, FT, z=z, override=feats)
    e_t = (traj - ref.detach()).abs().max().item()
    seed = torch.randn(traj.shape, generator=torch.Generator().manual_seed(8))
    ref.backward(seed)
    got = bwd(seed)
    scale = z.grad.abs().max().item()
    e_g = (got - z.grad).abs().max().item()
    diag('bench-shape teacher-forced %s FT=%d: traj err %.1e | grad err %.3e (max %.3e)' % (tag, FT, e_t, e_g, scale))
    assert e_t < 2e-5 * (1.0 + 0.5 * FT) and e_g < 2e-4 * max(1.0, scale)


@pytest.mark.parametrize('tag', ['n32', 'n64', 'ragged'])
def test_edge_kernels_tensor_core_paths_vs_simt(tag):
    """A/B of the edge implementations (strive_edge_set_impl): 3 = tcgen05 forward + backward (TMEM accumulators, fp16 hi/lo split, the
    transposed weights read through MN-major descriptors), 2 = tcgen05 forward + mma.sync backward, 1 = mma.sync TF32 both ways, 0 = the
    fp32 SIMT kernels; forward (aggr, arg-max, first-step trajectory) and
    backward (dL/dz) at multi-tile scene sizes."""
    from strive_b200 import _cabi
    FT = 3
    sc = scene(tag, FT)
    seed = torch.randn(sc['z'].size(0), FT, 4, generator=torch.Generator().manual_seed(9))
    res = {}
    for impl in (0, 1, 2, 3):
        _cabi.set_edge_impl(impl)
        try:
            traj, tape, NA = gpu_forward(sc, FT)
            aggr = _tape(tape, 'aggr', 0, NA, FT, 64)
            arg = read_arg(tape, 0, NA, FT)
            _, _, bwd = _lowlevel(sc, FT)
            res[impl] = (traj, aggr, arg, bwd(seed))
        finally:
            _cabi.set_edge_impl(3)
    for impl, name in ((1, 'mma.sync'), (2, 'tcgen05 fwd + mma.sync bwd'), (3, 'tcgen05 fwd + bwd')):
        d_aggr = (res[0][1] - res[impl][1]).abs().max().item()
        d_t0 = (res[0][0][:, 0] - res[impl][0][:, 0]).abs().max().item()
        n_arg = int((res[0][2] != res[impl][2]).sum())
        gs = res[0][3].abs().max().item()
        d_g = (res[0][3] - res[impl][3]).abs().max().item()
        diag('edge A/B %s [%s vs simt]: |aggr| %.2e, first-step traj %.2e, arg-max mismatches %d of %d, |d_z| %.3e (max %.3e)' % (
            tag, name, d_aggr, d_t0, n_arg, res[0][2].numel(), d_g, gs))
        assert d_aggr < 2e-5 and d_t0 < 2e-6 and n_arg <= 2
        assert d_g < 5e-2 * max(1.0, gs)      # 3-step BPTT after ~1e-6 forward differences (crop pixel flips included)


def test_avoid_loss_128_agent_block_vs_reference_fixture():
    """AvoidCollLoss built WITHOUT ptr on 4 x 32 agents = one 128-agent collision block (a loss group of configs[1]) on the
    reference's own trajectory: loss, per-term means, counts, dL/dtraj, dL/dz against the unmodified reference's outputs."""
    from strive_b200.losses import AvoidCollLoss
    dev, model, env = ctx()
    g = golden('bench_shape')
    FT = int(g['FT'])
    sc = synth.make_scenes(int(g['g128_seed']), [int(v) for v in g['g128_sizes']], map_extent_m=EXTENT, M=2, FT=FT, collide_frac=1.0, offroad_frac=1.0)
    lw_un = O.unnorm_att(sc['lw'])
    mapixes = sc['map_idx'][sc['batch']]
    mod = AvoidCollLoss(REFINE_W, lw_un.to(dev), mapixes.to(dev), env, (sc['z'] + 0.1).to(dev), veh_coll_buffer=0.2)
    futd = O.unnorm_state(torch.from_numpy(g['g128_traj'])).to(dev).requires_grad_(True)
    zd = sc['z'].clone().to(dev).requires_grad_(True)
    out = mod(futd, zd, (sc['prior_mu'].to(dev), sc['prior_var'].to(dev)))
    out['loss'].backward()
    t = mod.last_terms[0].cpu()
    # the direct latent terms of the fixture's dL/dz (the BPTT part is checked by the teacher-forced tests)
    z = sc['z'].clone().requires_grad_(True)
    fut = O.unnorm_state(torch.from_numpy(g['g128_traj'])).requires_grad_(True)
    raster, dx, _ = world()
    ref = O.avoid_coll_loss(fut, z, (sc['prior_mu'], sc['prior_var']), sc['z'] + 0.1, REFINE_W, lw_un, mapixes, None, raster, dx, veh_coll_buffer=0.2)
    ref['loss'].backward()
    gf = g['g128_d_fut_un']
    e_f = np.abs(futd.grad.cpu().numpy() - gf)
    e_z = (zd.grad.cpu() - z.grad).abs().max().item()
    means = [float(t[1]), float(t[3]), float(t[5]), float(t[6])]
    diag('128-agent block: loss gpu %.5f reference %.5f | counts gpu %d/%d reference %s | means gpu %s reference %s | |d_traj| err max %.3e '
         '(n > 1e-4: %d of %d; max %.3e) |d_z direct| err %.3e' % (float(out['loss']), float(g['g128_loss']), int(t[2]), int(t[4]), g['g128_counts'],
                                                                    np.array2string(np.array(means), precision=5), np.array2string(g['g128_means'], precision=5),
                                                                    e_f.max(), int((e_f > 1e-4).sum()), e_f.size, np.abs(gf).max(), e_z))
    assert [int(t[2]), int(t[4])] == [int(v) for v in g['g128_counts']]
    assert abs(float(out['loss']) - float(g['g128_loss'])) < 1e-4 * abs(float(g['g128_loss']))
    assert np.abs(np.array(means) / g['g128_means'] - 1.0).max() < 1e-4
    # dL/dtraj: the vehicle term agrees to 1e-5; an env-term entry is w/count/pen_d * unit(centre - point), and `point` is a
    # float32 mean of ~600 world coordinates (good to ~1e-4 m): entries whose |centre - point| is a few cm move by ~1e-3 with
    # the summation order (see tests/test_r2_cpu.py::_loop_z_check) -- bounded here, counted, and few
    assert e_f.max() < 5e-3 * max(1.0, np.abs(gf).max()) and int((e_f > 1e-4).sum()) <= 0.02 * e_f.size
    assert e_z < 1e-5 * max(1.0, z.grad.abs().max().item())


@pytest.mark.parametrize('tag', ['g128', 'n64', 'ragged'])
def test_decode_vs_reference_fixture(tag):
    """decode_embedding at the measured shapes against the unmodified reference's trajectories (and dL/dz for a seeded
    d_traj): first step to rounding, later steps inside the envelope of the reference-precision oracle."""
    dev, model, env = ctx()
    raster, dx, sd = world()
    g = golden('bench_shape')
    FT = int(g['FT'])
    sc = synth.make_scenes(int(g[tag + '_seed']), [int(v) for v in g[tag + '_sizes']], map_extent_m=EXTENT, M=2, FT=FT, collide_frac=1.0, offroad_frac=1.0)
    graph = to_graph(sc, dev)
    embed = {'map_feat': sc['map_feat'].to(dev), 'past_feat': sc['past_feat'].to(dev)}
    z = sc['z'].clone().to(dev).requires_grad_(True)
    traj = model.decode_embedding(z, embed, graph, sc['map_idx'].to(dev), env, nfuture=FT)['future_pred']
    ref = g[tag + '_traj']
    e = np.abs(traj.detach().cpu().numpy() - ref).max(axis=(0, 2))
    with torch.no_grad():
        o32 = oracle_decode(sc, FT).numpy()
        sd64 = {k: v.double() for k, v in sd.items()}
        c = lambda t: t.double() if t.is_floating_point() else t
        o64 = O.decode(sd64, c(sc['z']), c(sc['map_feat']), c(sc['past_feat']), c(sc['past'][:, -1, :]), c(sc['lw']), c(sc['sem']), sc['ptr'],
                       sc['edge_index'], sc['map_idx'], raster, dx, FT).numpy()
    e_o = np.abs(o32 - ref).max(axis=(0, 2))
    e_64 = np.abs(o32 - o64).max(axis=(0, 2))                  # what fp32 rounding alone does to THIS rollout (reference precision)
    msg = '%s vs reference: per-step |gpu-ref| %s | |oracle32-ref| %s | |oracle32-oracle64| %s' % (
        tag, ' '.join('%.1e' % v for v in e), ' '.join('%.1e' % v for v in e_o), ' '.join('%.1e' % v for v in e_64))
    if tag != 'g128':
        seed = torch.randn(ref.shape, generator=torch.Generator().manual_seed(int(g[tag + '_seed']) + 100))
        traj.backward(seed.to(dev))
        gz, rz = z.grad.cpu().numpy().reshape(-1), g[tag + '_d_z'].reshape(-1)
        cos = float(np.dot(gz, rz) / (np.linalg.norm(gz) * np.linalg.norm(rz)))
        rel = np.abs(gz - rz).max() / np.abs(rz).max()
        msg += ' | dL/dz vs reference: cos %.6f, max err / max %.2e' % (cos, rel)
        assert cos > 0.9995 and rel < 0.05
    diag(msg)
    assert e[0] < 2e-6
    # The CPU restatement happens to reproduce the reference's poses to the last bit here, so |oracle32-ref| is no yardstick
    # for what rounding does.  With ~8 M nearest-pixel samples per re-encode at these sizes some crop pixels always sit within
    # rounding of a tie: any two fp32 evaluations part ways at the first re-encode (map_feat moves by 1e-5..1e-3) and the
    # rollout amplifies it -- the envelope is the fp32-vs-fp64 gap of the oracle itself, as for the small fixtures.
    assert bool((e <= 10.0 * np.maximum.accumulate(e_64) + 1e-5).all())


def test_edge_tile_table_covers_every_slot_size_and_the_mma_fallback():
    """Scenes of 1, 2, 9, 10, 25, 41, 65 and 129 agents put 8 / 8 / 8 / 16 / 24 / 40 / 64 / 128 rows per target into the 128-edge
    tiles of the tcgen05 edge kernels (several tiles per scene, partly filled last tiles, a tile that is one target); a 130-agent
    scene exceeds a tile and must fall back to the mma.sync kernels.  Forward (aggr, arg-max) and backward (dL/dz) against the fp32
    SIMT kernels."""
    from strive_b200 import _cabi
    FT = 2
    for sizes in ([1, 2, 9, 10, 25, 41, 65, 129], [130, 3]):
        sc = synth.make_scenes(91, sizes, map_extent_m=EXTENT, M=2, FT=FT, collide_frac=1.0, offroad_frac=1.0)
        seed = torch.randn(sc['z'].size(0), FT, 4, generator=torch.Generator().manual_seed(10))
        res = {}
        for impl in (0, 3):
            _cabi.set_edge_impl(impl)
            try:
                traj, tape, NA = gpu_forward(sc, FT)
                res[impl] = (traj, _tape(tape, 'aggr', 0, NA, FT, 64), read_arg(tape, 0, NA, FT), _lowlevel(sc, FT)[2](seed))
            finally:
                _cabi.set_edge_impl(3)
        d_aggr = (res[0][1] - res[3][1]).abs().max().item()
        n_arg = int((res[0][2] != res[3][2]).sum())
        gs = res[0][3].abs().max().item()
        d_g = (res[0][3] - res[3][3]).abs().max().item()
        diag('edge tile table sizes %s: |aggr tc-simt| %.2e, arg-max mismatches %d of %d, |d_z| %.3e (max %.3e)' % (sizes, d_aggr, n_arg, res[0][2].numel(), d_g, gs))
        assert d_aggr < 2e-5 and n_arg <= 2
        assert d_g < 5e-2 * max(1.0, gs)
        if sizes[0] == 1:
            assert bool((res[3][1][0] == 0).all()) and bool((res[3][2][0] == 255).all())      # no in-edges: message 0, arg-max 255 (interaction_net.py:187-188)
