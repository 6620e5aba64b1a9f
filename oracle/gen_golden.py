"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/src, imported behind oracle/ref_shims.py) on seeded synthetic scenes.

Run in the build container only (the reference does not exist on the GPU box):
    python oracle/gen_golden.py            # writes tests/golden/*.npz and prints oracle-vs-reference diffs

Inputs are NOT stored: they are regenerated from seeds by strive_b200.synth (checksums are stored and
verified by the tests).  Outputs stored are those of the reference code paths cited per case.
"""
import os
import sys
import io
import contextlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims                     # noqa: E402
from oracle import strive_oracle as O             # noqa: E402
from strive_b200 import synth                     # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')

REFINE_W = {'coll_veh': 100.0, 'coll_env': 100.0, 'motion_prior': 1.0, 'init_z': 0.01}   # configs/refine_traffic_optim.cfg:26-29
ADV_W = {'coll_veh': 20.0, 'coll_veh_plan': 20.0, 'coll_env': 20.0, 'init_z': 0.5, 'init_z_atk': 0.05,
         'motion_prior': 1.0, 'motion_prior_atk': 0.005, 'motion_prior_ext': 0.0001, 'match_ext': 10.0,
         'adv_crash': 2.0}                                                                  # configs/adv_gen_rule_based.cfg:34-43
SOL_W = {'motion_prior': 0.005, 'coll_veh': 10.0, 'coll_env': 10.0, 'motion_prior_ext': 0.001,
         'match_ext': 10.0, 'init_z': 0.0}                                                  # configs/adv_gen_rule_based.cfg:45-50

# small world used by the fixtures: two maps of 1280x1280 px
RASTER_KW = dict(seed=3, M=2, H=1280, W=1280)
EXTENT = (90.0, 230.0)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def ref_graph(scene):
    from torch_geometric.data import Data
    NA = scene['past'].size(0)
    return Data(past=scene['past'].clone(), past_gt=scene['past'].clone(),
                past_vis=torch.ones(NA, scene['past'].size(1)),
                lw=scene['lw'].clone(), sem=scene['sem'].clone(), edge_index=scene['edge_index'],
                ptr=scene['ptr'], batch=scene['batch'], x=None, pos=None)


def build(seed_w=0):
    ref_shims.install()
    raster, dx = synth.make_raster(**RASTER_KW)
    env = ref_shims.make_map_env(raster, dx)
    model = quiet(ref_shims.make_ref_model, nfuture=20)
    sd = synth.make_weights(seed_w)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    model.train()    # as the drivers do (refine_traffic_optim.py:487)
    return raster, dx, env, model, sd


def maxdiff(a, b):
    return float((a.detach().double() - b.detach().double()).abs().max())


def case_encode_map(raster, dx, env, model, sd, out):
    """TrafficModel.encode_map (models/traffic_model.py:416-451) on hand-picked poses incl. out-of-map."""
    poses_un = torch.tensor([
        [150.0, 150.0, 1.0, 0.0],
        [123.4, 171.9, 0.6, 0.8],
        [200.3, 96.7, -0.28, 0.96],
        [15.0, 20.0, 0.0, -1.0],        # crop partly outside the map -> pixel (0,0) rule
        [310.0, 250.0, -1.0, 0.0],      # mostly outside on the far side
        [160.25, 149.875, 0.70710678, 0.70710678],
    ])
    mapix = torch.tensor([0, 1, 0, 1, 0, 1])
    pos_n = O.norm_state(poses_un)
    from torch_geometric.data import Data
    sg = Data(pos=pos_n.clone(), batch=torch.arange(6), lw=torch.zeros(6, 2))
    with torch.no_grad():
        ref_feat = model.encode_map(sg, mapix, env)
        crop = env.get_map_crop(Data(pos=poses_un, batch=torch.arange(6)), mapix)
        my_crop = O.map_crop(raster, dx, O.unnorm_state(pos_n), mapix)
        my_feat = O.encode_map(sd, raster, dx, pos_n, mapix)
    # NB: encode_map unnormalises the normalised pose; compare crops on that exact pose
    crop2 = env.get_map_crop(Data(pos=O.unnorm_state(pos_n), batch=torch.arange(6)), mapix)
    print('encode_map: crop mismatches %d, feat maxdiff %.3e' % (int((crop2 != my_crop).sum()), maxdiff(ref_feat, my_feat)))
    out['encode_map'] = dict(pos_n=pos_n.numpy(), mapix=mapix.numpy(), map_feat=ref_feat.numpy(),
                             crop_sum=crop2.long().sum(dim=(2, 3)).numpy(),
                             crop_rowsum=crop2.long().sum(dim=3).numpy().astype(np.int32))


def decode_ref(model, env, scene, z, FT, ext=None):
    sg = ref_graph(scene)
    embed = {'map_feat': scene['map_feat'], 'past_feat': scene['past_feat']}
    return model.decode_embedding(z, embed, sg, scene['map_idx'], env, ext_future=ext, nfuture=FT)['future_pred']


def decode_mine(sd, raster, dx, scene, z, FT, ext=None):
    return O.decode(sd, z, scene['map_feat'], scene['past_feat'], scene['past'][:, -1, :], scene['lw'], scene['sem'],
                    scene['ptr'], scene['edge_index'], scene['map_idx'], raster, dx, FT, ext_future=ext)


def case_decode(name, raster, dx, env, model, sd, out, seed, sizes, FT, with_ext):
    """TrafficModel.decode_embedding (models/traffic_model.py:405-414 -> 589-704)."""
    scene = synth.make_scenes(seed, sizes, map_extent_m=EXTENT, M=2, FT=FT, collide_frac=1.0, offroad_frac=1.0)
    ext = scene['ext_future'] if with_ext else None
    with torch.no_grad():
        ref = decode_ref(model, env, scene, scene['z'], FT, ext)
        mine = decode_mine(sd, raster, dx, scene, scene['z'], FT, ext)
    print('%s: traj maxdiff %.3e' % (name, maxdiff(ref, mine)))
    out[name] = dict(seed=seed, sizes=np.array(sizes), FT=FT, with_ext=int(with_ext), traj=ref.numpy(),
                     z_sum=synth.checksum(scene['z']), past_sum=synth.checksum(scene['past']))


def case_refine(raster, dx, env, model, sd, out, seed=11, sizes=(3, 1, 5), FT=6, iters=5, lr=0.05):
    """refine_traffic_optim.py:163-218: Adam([z]) + AvoidCollLoss(veh_coll_buffer=0.2), reference modules."""
    from losses.adv_gen_nusc import AvoidCollLoss
    scene = synth.make_scenes(seed, list(sizes), map_extent_m=EXTENT, M=2, FT=FT, collide_frac=1.0, offroad_frac=1.0)
    z = scene['z'].clone().requires_grad_(True)
    opt = torch.optim.Adam([z], lr=lr)
    lw_un = model.get_att_normalizer().unnormalize(scene['lw'])
    loss_fn = quiet(AvoidCollLoss, REFINE_W, lw_un, scene['map_idx'][scene['batch']], env, z.clone().detach(),
                    veh_coll_buffer=0.2)
    prior = (scene['prior_mu'], scene['prior_var'])
    rec = {'loss': [], 'grad': [], 'z': [], 'terms': []}
    traj0 = None
    for it in range(iters):
        opt.zero_grad()
        fut = decode_ref(model, env, scene, z, FT)
        ld = loss_fn(model.get_normalizer().unnormalize(fut), z, prior)
        ld['loss'].backward()
        if it == 0:
            traj0 = fut.detach().clone()
        rec['loss'].append(float(ld['loss']))
        rec['terms'].append([float(ld[k].mean()) for k in ('coll_veh_loss', 'coll_env_loss', 'motion_prior_loss', 'init_loss')]
                            + [float(ld['coll_veh_loss'].numel()), float(ld['coll_env_loss'].numel())])
        rec['grad'].append(z.grad.detach().clone())
        opt.step()
        rec['z'].append(z.detach().clone())
    # my oracle, same loop
    mine = []
    zf = O.refine_loop(sd, scene, raster, dx, REFINE_W, iters, lr, FT, veh_coll_buffer=0.2, record=mine)
    print('refine: loss0 ref %.6f mine %.6f | grad0 maxdiff %.3e (|g|max %.3e) | z_final maxdiff %.3e' % (
        rec['loss'][0], mine[0]['loss'], maxdiff(rec['grad'][0], mine[0]['grad']), float(rec['grad'][0].abs().max()),
        maxdiff(rec['z'][-1], zf)))
    print('        terms0', rec['terms'][0])
    out['refine'] = dict(seed=seed, sizes=np.array(sizes), FT=FT, iters=iters, lr=lr, traj0=traj0.numpy(),
                         loss=np.array(rec['loss']), terms=np.array(rec['terms']),
                         grad=torch.stack(rec['grad']).numpy(), z=torch.stack(rec['z']).numpy())


def case_losses(raster, dx, env, model, sd, out, seed=21, sizes=(4, 6), FT=8):
    """AdvGenLoss / TgtMatchingLoss / AvoidCollLoss(single_veh_idx=0) forward + grads on a fixed trajectory
    (losses/adv_gen_nusc.py:14-341), as used by adv_gen_optim.py:74-154 and sol_optim.py:47-97."""
    from losses.adv_gen_nusc import AdvGenLoss, TgtMatchingLoss, AvoidCollLoss
    scene = synth.make_scenes(seed, list(sizes), map_extent_m=EXTENT, M=2, FT=FT, collide_frac=1.0, offroad_frac=1.0)
    ptr = scene['ptr']
    NA = int(ptr[-1])
    B = len(sizes)
    ego_mask = torch.zeros(NA, dtype=torch.bool)
    ego_mask[ptr[:-1]] = True
    with torch.no_grad():
        fut_n = decode_ref(model, env, scene, scene['z'], FT)
    lw_un = model.get_att_normalizer().unnormalize(scene['lw'])
    mapixes = scene['map_idx'][scene['batch']]
    res = {}
    # --- adversarial loss
    fut = O.unnorm_state(fut_n).clone().requires_grad_(True)
    tgt = O.unnorm_state(scene['ext_future'][:, :FT]).clone()
    z_o = scene['z'][~ego_mask].clone().requires_grad_(True)
    prior_o = (scene['prior_mu'][~ego_mask], scene['prior_var'][~ego_mask])
    init_o = (scene['z'][~ego_mask] + 0.05).clone()
    adv = quiet(AdvGenLoss, ADV_W, lw_un, mapixes, env, init_o, ptr, veh_coll_buffer=0.1,
                crash_loss_min_time=2, crash_loss_min_infront=-0.5)
    ld = quiet(adv, fut, tgt, z_o, prior_o, return_mins=True)
    ld['loss'].backward()
    res['adv'] = dict(loss=float(ld['loss']), d_fut=fut.grad.clone(), d_z=z_o.grad.clone(),
                      crash=ld['adv_crash_loss'].detach().clone(), min_agt=ld['min_agt'], min_t=ld['min_t'],
                      means=[float(ld[k].mean()) for k in ('init_loss', 'motion_prior_loss', 'coll_veh_loss',
                                                          'coll_veh_plan_loss', 'coll_env_loss', 'adv_crash_loss')])
    fut2 = O.unnorm_state(fut_n).clone().requires_grad_(True)
    z2 = scene['z'][~ego_mask].clone().requires_grad_(True)
    md = O.adv_gen_loss(fut2, tgt, z2, prior_o, init_o, ADV_W, lw_un, mapixes, ptr, raster, dx,
                        veh_coll_buffer=0.1, crash_min_t=2, crash_min_infront=-0.5)
    md['loss'].backward()
    print('adv loss ref %.6f mine %.6f | d_fut maxdiff %.3e | d_z maxdiff %.3e | mins %s %s vs %s %s' % (
        res['adv']['loss'], float(md['loss']), maxdiff(fut.grad, fut2.grad), maxdiff(z_o.grad, z2.grad),
        list(ld['min_agt']), list(ld['min_t']), md['min_agt'], md['min_t']))
    # --- target matching (with the :46 bug)
    futm = O.unnorm_state(fut_n)[ego_mask].clone().requires_grad_(True)
    tm = TgtMatchingLoss(ADV_W)
    zt = scene['z'][ego_mask].clone().requires_grad_(True)
    ldm = tm(futm, tgt, zt, (scene['prior_mu'][ego_mask], scene['prior_var'][ego_mask]))
    ldm['loss'].backward()
    futm2 = O.unnorm_state(fut_n)[ego_mask].clone().requires_grad_(True)
    mdm = O.tgt_matching_loss(futm2, tgt, ADV_W)
    mdm['loss'].backward()
    print('match loss ref %.6f mine %.6f | d_fut maxdiff %.3e | z.grad is None: %s' % (
        float(ldm['loss']), float(mdm['loss']), maxdiff(futm.grad, futm2.grad), zt.grad is None))
    res['match'] = dict(loss=float(ldm['loss']), d_fut=futm.grad.clone())
    # --- solution-phase avoid loss on the ego only
    futs = O.unnorm_state(fut_n).clone().requires_grad_(True)
    zs = scene['prior_mu'][ego_mask].clone().requires_grad_(True)
    init_s = zs.detach().clone()
    av = quiet(AvoidCollLoss, SOL_W, lw_un, mapixes, env, init_s, veh_coll_buffer=0.5, single_veh_idx=0, ptr=ptr)
    lds = av(futs, zs, (scene['prior_mu'][ego_mask], scene['prior_var'][ego_mask]))
    lds['loss'].backward()
    futs2 = O.unnorm_state(fut_n).clone().requires_grad_(True)
    zs2 = scene['prior_mu'][ego_mask].clone().requires_grad_(True)
    mds = O.avoid_coll_loss(futs2, zs2, (scene['prior_mu'][ego_mask], scene['prior_var'][ego_mask]), init_s, SOL_W,
                            lw_un, mapixes, ptr, raster, dx, veh_coll_buffer=0.5, single_veh_idx=0)
    mds['loss'].backward()
    print('sol avoid loss ref %.6f mine %.6f | d_fut maxdiff %.3e | d_z maxdiff %.3e' % (
        float(lds['loss']), float(mds['loss']), maxdiff(futs.grad, futs2.grad), maxdiff(zs.grad, zs2.grad)))
    res['sol'] = dict(loss=float(lds['loss']), d_fut=futs.grad.clone(), d_z=zs.grad.clone())
    out['losses'] = dict(seed=seed, sizes=np.array(sizes), FT=FT, fut_n=fut_n.numpy(),
                         adv_loss=res['adv']['loss'], adv_d_fut=res['adv']['d_fut'].numpy(), adv_d_z=res['adv']['d_z'].numpy(),
                         adv_crash=res['adv']['crash'].numpy(), adv_min_agt=np.array(res['adv']['min_agt']),
                         adv_min_t=np.array(res['adv']['min_t']), adv_means=np.array(res['adv']['means']),
                         match_loss=res['match']['loss'], match_d_fut=res['match']['d_fut'].numpy(),
                         sol_loss=res['sol']['loss'], sol_d_fut=res['sol']['d_fut'].numpy(), sol_d_z=res['sol']['d_z'].numpy())


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    raster, dx, env, model, sd = build()
    out = {}
    case_encode_map(raster, dx, env, model, sd, out)
    case_decode('decode_small', raster, dx, env, model, sd, out, seed=5, sizes=[3, 1, 5], FT=6, with_ext=False)
    case_decode('decode_ext', raster, dx, env, model, sd, out, seed=5, sizes=[3, 1, 5], FT=6, with_ext=True)
    case_decode('decode_c1', raster, dx, env, model, sd, out, seed=9, sizes=[8], FT=20, with_ext=False)
    case_refine(raster, dx, env, model, sd, out)
    case_losses(raster, dx, env, model, sd, out)
    os.makedirs(GOLD, exist_ok=True)
    meta = dict(weights_sum=synth.checksum(torch.cat([v.reshape(-1) for v in sd.values()])),
                raster_sum=float(raster.double().sum()), dx=dx.numpy())
    np.savez_compressed(os.path.join(GOLD, 'meta.npz'), **meta)
    for name, d in out.items():
        np.savez_compressed(os.path.join(GOLD, name + '.npz'), **d)
        print('wrote', name, sum(np.asarray(v).nbytes for v in d.values()), 'bytes')


if __name__ == '__main__':
    main()
