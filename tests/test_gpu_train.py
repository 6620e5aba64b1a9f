"""GPU: one train_traffic step against the UNMODIFIED reference's loss and parameter gradients
(tests/golden/train_step.npz, oracle/gen_golden_r2.py case_train: TrafficModel.forward(future_sample=True) -> TrafficModelLoss ->
backward with fixed rsample noise)."""
import numpy as np
import pytest
import torch

from strive_b200 import synth
from tests.common import golden, EXTENT
from tests.test_gpu_parity import ctx, diag, to_graph
from tests.test_r2_cpu import full_weights

pytestmark = pytest.mark.gpu
TRAIN_W = {'recon': 1.0, 'kl': 0.004, 'coll_veh_prior': 0.05, 'coll_env_prior': 0.1}


def _case(dev):
    g = golden('train_step')
    FT = int(g['FT'])
    seed = int(g['seed'])
    sc = synth.make_scenes(seed, [int(v) for v in g['sizes']], map_extent_m=EXTENT, M=2, FT=FT, collide_frac=1.0, offroad_frac=1.0)
    fut, fvis, pvis = synth.make_future(seed + 1, sc, FT)
    gen = torch.Generator().manual_seed(seed + 2)
    NA = sc['z'].size(0)
    e_post, e_prior = torch.randn(NA, 32, generator=gen), torch.randn(NA, 32, generator=gen)
    gr = to_graph(sc, dev)
    gr.past_vis, gr.future, gr.future_gt, gr.future_vis = pvis.to(dev), fut.to(dev), fut.to(dev), fvis.to(dev)
    return g, sc, gr, e_post.to(dev), e_prior.to(dev), FT


def _trainer(FT, autocast):
    import strive_b200
    from strive_b200.train import TrafficModelTrainer
    dev, _, env = ctx()
    _, _, sd = full_weights(FT)
    model = strive_b200.make_model(nfuture=FT, state_dict=sd, device=dev)
    return model, TrafficModelTrainer(model, env, TRAIN_W, lr=1e-5, autocast_bf16=autocast)


def test_train_step_loss_and_all_parameter_grads_vs_reference_fp32():
    dev, _, env = ctx()
    g, sc, gr, e_post, e_prior, FT = _case(dev)
    model, tr = _trainer(FT, autocast=False)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False               # cuDNN convolutions default to TF32 on this GPU: not the fp32 reference arithmetic
    try:
        ld = tr.backward(gr, sc['map_idx'].to(dev), eps_post=e_post, eps_prior=e_prior)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    got = {'loss': float(ld['loss']), 'recon': float(ld['recon_loss'].mean()), 'kl': float(ld['kl_loss'].mean()),
           'coll_veh': float(ld['coll_veh_prior']), 'coll_env': float(ld['coll_env_prior'].mean())}
    rel = {k: abs(v - float(g[k])) / max(1e-6, abs(float(g[k]))) for k, v in got.items()}
    names = [str(n) for n in g['names']]
    params = dict(model.named_parameters())
    assert sorted(names) == sorted(params.keys())                          # same parameter tree as the reference checkpoint
    gen = torch.Generator().manual_seed(99)
    e_norm, e_proj = [], []
    tot_ref = float(np.sqrt((g['grad_norm'] ** 2).sum()))
    for i, k in enumerate(names):
        gk = params[k].grad.detach().double().cpu()
        r = torch.randn(gk.numel(), generator=gen, dtype=torch.float64)
        e_norm.append(abs(float(gk.norm()) - g['grad_norm'][i]))
        e_proj.append(abs(float((gk.reshape(-1) * r).sum()) - g['grad_proj'][i]) / max(1e-12, g['grad_norm'][i] * float(r.norm()) / np.sqrt(gk.numel())))
    worst = int(np.argmax(np.array(e_norm) / (g['grad_norm'] + 1e-3 * tot_ref)))
    d_b = np.abs(model.map_feature.bias.grad.cpu().numpy() - g['grad_map_feature_bias']).max() / np.abs(g['grad_map_feature_bias']).max()
    d_g = np.abs(model.decoder_memory.bias_hh_l2.grad.cpu().numpy() - g['grad_gru_bias']).max() / np.abs(g['grad_gru_bias']).max()
    diag('train step fp32 vs reference: loss terms rel err %s | %d parameter tensors: max |norm err| / total |grad| = %.2e (worst %s), '
         'random-projection err (in units of the tensor rms) median %.2e max %.2e | map_feature.bias grad rel %.2e, GRU bias_hh_l2 grad rel %.2e' % (
             ' '.join('%s=%.1e' % kv for kv in rel.items()), len(names), max(e_norm) / tot_ref, names[worst], float(np.median(e_proj)), max(e_proj), d_b, d_g))
    assert max(rel.values()) < 1e-3
    assert max(e_norm) / tot_ref < 2e-3 and float(np.median(e_proj)) < 2e-2
    # single small tensors at the end of a 5-step BPTT through nearest-pixel crops (two fp32 evaluations part ways at the first
    # re-encode, tests/test_gpu_benchshape.py): looser than the aggregate measures above
    assert d_b < 5e-2 and d_g < 1.5e-1
    # first rollout step of both decodes is free of amplification
    fp = tr.forward(gr, sc['map_idx'].to(dev), future_sample=True, eps_post=e_post, eps_prior=e_prior)
    assert np.abs(fp['future_pred'][:, 0].detach().cpu().numpy() - g['future_pred'][:, 0]).max() < 5e-6
    assert np.abs(fp['future_samp'][:, 0].detach().cpu().numpy() - g['future_samp'][:, 0]).max() < 5e-6


def test_train_step_bf16_autocast_tracks_fp32_and_adam_moves_weights():
    dev, _, env = ctx()
    g, sc, gr, e_post, e_prior, FT = _case(dev)
    model, tr = _trainer(FT, autocast=True)
    ld = tr.backward(gr, sc['map_idx'].to(dev), eps_post=e_post, eps_prior=e_prior)
    rel = abs(float(ld['loss']) - float(g['loss'])) / float(g['loss'])
    names = [str(n) for n in g['names']]
    params = dict(model.named_parameters())
    gn = np.array([float(params[k].grad.double().norm()) for k in names])
    cos = float((gn * g['grad_norm']).sum() / (np.linalg.norm(gn) * np.linalg.norm(g['grad_norm'])))
    w0 = tr.bucket.flat_p.clone()
    tr.opt.step()
    moved = float((tr.bucket.flat_p - w0).abs().max())
    diag('train step bf16 autocast: loss rel err vs fp32 reference %.2e, per-tensor gradient-norm profile cosine %.5f, Adam moved weights by %.2e' % (rel, cos, moved))
    assert rel < 3e-2 and cos > 0.99 and 0.0 < moved <= 1.1e-5
