"""Full-size GPU tests (BASELINE.json configs[1], [2], [4] per-GPU shares): the CPU oracle cannot finish these sizes in
seconds, so they check size-independent properties of the CUDA path instead:
  * the rollout of a scene does not depend on which other scenes share the batch, nor on their order: BITWISE over the whole
    horizon (every kernel of the forward pass accumulates in an order that depends on the item's place inside its own crop /
    scene only; conv3's K-chunk order follows the global tile-pair parity);
  * loss-normalisation groups are independent: a sub-batch of whole groups reproduces the full batch's per-group loss terms
    exactly and its gradient rows to the rounding noise of the float atomics in the backward pass;
  * two evaluations of the same batch give bitwise identical trajectories (no races in the forward kernels);
  * everything stays finite over the full horizon; ragged scenes (4..40 agents) run through the adv / solution loops.
"""
import numpy as np
import pytest
import torch

from tests.common import REFINE_W, ADV_W, SOL_W
from tests.test_gpu_parity import to_graph, diag

pytestmark = pytest.mark.gpu

_w = {}


def big_world():
    """4096 x 4096 synthetic raster (the bench world) + model, built once."""
    if not _w:
        import strive_b200
        from strive_b200 import synth
        dev = torch.device('cuda:0')
        raster, dx = synth.make_raster(seed=1, M=1, H=4096, W=4096)
        sd = synth.make_weights(0)
        _w['dev'] = dev
        _w['model'] = strive_b200.make_model(nfuture=20, state_dict=sd, device=dev)
        _w['env'] = strive_b200.MapEnv(raster, dx, device=dev)
    return _w['dev'], _w['model'], _w['env']


def make_loop(sc, gptr, FT, dev, model, env):
    from strive_b200.optim import RefineLoop
    g = to_graph(sc, dev)
    embed = {'map_feat': sc['map_feat'].to(dev), 'past_feat': sc['past_feat'].to(dev),
             'prior_out': (sc['prior_mu'].to(dev), sc['prior_var'].to(dev))}
    return RefineLoop(model, g, sc['map_idx'].to(dev), env, embed, sc['z'].to(dev), REFINE_W, 0.05, FT, veh_coll_buffer=0.2,
                      group_scene_ptr=gptr)


def subset_check(name, sc, gptr, groups, FT, full_loop, dev, model, env, cos_min=0.4):
    """Runs `groups` (any order) alone and compares with the rows / groups of the full batch."""
    from strive_b200 import shard
    sub, lgptr, idx = shard.shard_scenes(sc, gptr, groups)
    loop = make_loop(sub, lgptr, FT, dev, model, env)
    loop.step()
    torch.cuda.synchronize()
    tr_sub, tr_full = loop.traj.cpu(), full_loop['traj'][idx]
    d_t = (tr_sub - tr_full).abs().amax(dim=(0, 2))                      # per rollout step
    t_sub = loop.terms.cpu()
    t_full = full_loop['terms'][torch.tensor(groups)]
    e_loss = ((t_sub[:, 0] - t_full[:, 0]).abs() / (t_full[:, 0].abs() + 1.0)).max().item()
    g_sub, g_full = loop.grad().cpu().flatten().double(), full_loop['grad'][idx].flatten().double()
    cos = float((g_sub * g_full).sum() / (g_sub.norm() * g_full.norm() + 1e-30))
    diag('%s: groups %s alone vs inside the full batch: traj diff per step %s | loss rel err %.2e | grad cosine %.5f' % (
        name, groups, ' '.join('%.1e' % v for v in d_t.tolist()), e_loss, cos))
    assert d_t.max().item() == 0.0                    # sharded == unsharded to the last bit, over the whole horizon
    # the forward pass is identical, so the loss terms are (the per-group sums run over the same values in the same kernels) and
    # the gradients differ by the order of the float atomics in the backward pass only
    assert e_loss < 1e-6 and cos > max(cos_min, 0.999)


def full_step(loop):
    loop._forward()
    torch.cuda.synchronize()
    traj0 = loop.traj.cpu().clone()
    loop._forward()                      # same inputs again: must be bitwise identical
    loop._loss()
    loop._backward()
    torch.cuda.synchronize()
    out = {'traj': loop.traj.cpu().clone(), 'terms': loop.terms.cpu().clone(), 'grad': loop.grad().cpu().clone()}
    assert torch.equal(traj0, out['traj']), 'forward rollout is not deterministic'
    for k, v in out.items():
        assert bool(torch.isfinite(v).all()), '%s has non-finite entries' % k
    return out


def test_c2_refine_full_size_group_independence():
    """BASELINE configs[1]: 64 scenes x 32 agents x 20 steps, groups of 4 scenes."""
    from strive_b200 import synth
    dev, model, env = big_world()
    FT, S, n = 20, 64, 32
    sc = synth.make_scenes(1000, [n] * S, map_extent_m=(200.0, 800.0), M=1, FT=FT, collide_frac=0.25, offroad_frac=0.25)
    gptr = list(range(0, S + 1, 4))
    loop = make_loop(sc, gptr, FT, dev, model, env)
    full = full_step(loop)
    active = (full['terms'][:, 2] > 0).sum().item(), (full['terms'][:, 4] > 0).sum().item()
    diag('c2 full size: NA %d, loss %.4f, groups with vehicle collisions %d / env collisions %d of %d, |grad| max %.3e' % (
        loop.NA, float(full['terms'][:, 0].sum()), active[0], active[1], len(gptr) - 1, full['grad'].abs().max().item()))
    assert active[0] > 0 and active[1] > 0      # both collision terms are exercised at this size
    subset_check('c2', sc, gptr, [11, 2, 7], FT, full, dev, model, env)
    # a few Adam iterations reduce the loss
    l0 = float(full['terms'][:, 0].sum())
    loop.z.copy_(sc['z'].to(dev))
    for _ in range(4):
        loop.step()
    torch.cuda.synchronize()
    l4 = float(loop.terms[:, 0].sum())
    diag('c2 full size: loss %.3f -> %.3f after 4 Adam iterations' % (l0, l4))
    assert np.isfinite(l4) and l4 < l0


def test_c5_stress_share_full_size():
    """BASELINE configs[4] per-GPU share: 128 scenes x 64 agents x 40 steps (8192 agents, 4 map-encoder chunks per step)."""
    from strive_b200 import synth
    dev, model, env = big_world()
    FT, S, n = 40, 128, 64
    sc = synth.make_scenes(2000, [n] * S, map_extent_m=(200.0, 800.0), M=1, FT=FT, collide_frac=0.25, offroad_frac=0.25)
    gptr = list(range(0, S + 1, 4))
    model.FT = FT
    try:
        loop = make_loop(sc, gptr, FT, dev, model, env)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        full = full_step(loop)
        e0.record()
        loop.step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        diag('c5 share: NA %d FT %d, tape %.0f MB, %.1f ms per iteration = %.0f agent*timestep*iter/s, loss %.3f' % (
            loop.NA, FT, loop.tape_bytes / 1e6, ms, loop.NA * FT / (ms / 1e3), float(full['terms'][:, 0].sum())))
        subset_check('c5', sc, gptr, [30, 1], FT, full, dev, model, env, cos_min=0.2)
    finally:
        model.FT = 20


def test_c3_ragged_adv_and_solution_loops():
    """BASELINE configs[2]: ragged scenes (4..40 agents, ~512 agents in total), FT 12 adversarial + FT 16 solution loops in
    planner-replay mode, a few iterations each."""
    from strive_b200 import synth
    from strive_b200.optim import run_adv_gen_optim, run_find_solution_optim
    dev, model, env = big_world()
    rng = np.random.RandomState(5)
    sizes = []
    while sum(sizes) < 512:
        sizes.append(int(rng.randint(4, 41)))
    FT, FTs = 12, 16
    sc = synth.make_scenes(3000, sizes, map_extent_m=(200.0, 800.0), M=1, FT=FTs, collide_frac=0.5, offroad_frac=0.25)
    NA = int(sc['ptr'][-1])
    ego = torch.zeros(NA, dtype=torch.bool)
    ego[sc['ptr'][:-1]] = True
    graph = to_graph(sc, dev)
    pf = sc['ext_future'][:, :FT].contiguous()
    fg = torch.zeros(NA, FT, 6)
    fg[ego, :, :4] = pf
    graph.future_gt = fg.to(dev)
    embed = {'map_feat': sc['map_feat'].to(dev), 'past_feat': sc['past_feat'].to(dev)}
    tgt_prior = (sc['prior_mu'][ego].to(dev), sc['prior_var'][ego].to(dev))
    oth_prior = (sc['prior_mu'][~ego].to(dev), sc['prior_var'][~ego].to(dev))
    logs = []
    model.FT = FT
    try:
        z, traj, out, min_agt, min_t = run_adv_gen_optim(sc['z'].to(dev), 0.05, ADV_W, model, graph, env, sc['map_idx'].to(dev), 6, embed,
                                                          'ego', tgt_prior, oth_prior, 1, -0.5, future_len=FT, veh_coll_buffer=0.1,
                                                          log=lambda it, d: logs.append(d))
        a0, a5 = logs[0]['adv_loss'], logs[-1]['adv_loss']
        assert tuple(z.shape) == (NA, 32) and tuple(traj.shape) == (NA, 1, FT, 4)
        assert bool(torch.isfinite(z).all()) and bool(torch.isfinite(traj).all())
        assert len(min_agt) == len(sizes) and all(int(sc['ptr'][s]) < int(min_agt[s]) < int(sc['ptr'][s + 1]) for s in range(len(sizes)))
        assert np.isfinite(a5) and a5 < a0
        slogs = []
        wfull = {'sol_' + k: v for k, v in SOL_W.items()}
        z2, sol_traj, _ = run_find_solution_optim(z, traj, FTs, 0.05, wfull, model, graph, env, sc['map_idx'].to(dev), 4, embed,
                                                  tgt_prior, oth_prior, log=lambda it, d: slogs.append(d))
        assert tuple(z2.shape) == (NA, 1, 32) and tuple(sol_traj.shape) == (NA, 1, FT, 4)
        assert bool(torch.isfinite(z2).all()) and bool(torch.isfinite(sol_traj).all())
        # non-target agents are pinned to the adversarial result (sol_optim.py:120-121)
        assert torch.allclose(sol_traj[~ego.to(dev)], traj[~ego.to(dev)], atol=1e-5)
        s0 = slogs[0]['tgt_loss'] + slogs[0]['other_loss']
        s3 = slogs[-1]['tgt_loss'] + slogs[-1]['other_loss']
        diag('c3 ragged: %d scenes (%d..%d agents), NA %d | adv loss %.3f -> %.3f in 6 iters | sol loss %.4f -> %.4f in 4 iters' % (
            len(sizes), min(sizes), max(sizes), NA, a0, a5, s0, s3))
        assert np.isfinite(s3)
    finally:
        model.FT = 20
