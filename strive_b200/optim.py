"""Latent-optimisation loops (reference src/refine_traffic_optim.py:146-226, src/utils/init_optim.py,
adv_gen_optim.py, sol_optim.py) in two forms:

  * `refine_traffic_optim(...)`  -- the reference function's signature and semantics, built from the drop-in
    TrafficModel.decode_embedding + AvoidCollLoss + torch.optim.Adam (public API path; what a STRIVE user calls).
  * `RefineLoop`                 -- the same iteration as ONE device-resident pipeline with no host synchronisation:
    strive_decode_fwd -> strive_loss_fwd_bwd -> strive_decode_bwd -> strive_adam_step, all buffers preallocated,
    optionally replayed as a CUDA graph.  Scenes may be split into loss-normalisation groups (= the batches the
    reference driver would have formed, refine_traffic_optim.py:291-310); each group is optimised exactly as the
    reference would optimise that batch on its own.
"""
import ctypes as C

import torch

from . import _cabi
from .losses import LossPlan, run_loss, AvoidCollLoss
from .runtime import SceneBatch


def detach_embed_info(embed):
    """reference src/utils/scenario_gen.py:19-28"""
    out = {}
    for k, v in embed.items():
        if isinstance(v, torch.Tensor):
            out[k] = v.detach()
        elif isinstance(v, tuple):
            out[k] = (v[0].detach(), v[1].detach())
    return out


def refine_traffic_optim(scene_graph, map_idx, map_env, model, loss_weights, num_iters, samp_future_len,
                         save_future_len, optim_use_adam, lr, log=None):
    """Same contract as reference refine_traffic_optim.py:146-226 (Adam branch).  `log(iter, dict)` is optional and, when
    given, is the only thing that synchronises with the host (the reference prints + .item()s every iteration)."""
    if not optim_use_adam:
        raise RuntimeError('strive_b200: only the Adam branch of refine_traffic_optim is supported (the reference default)')
    with torch.no_grad():
        sample_pred = model.sample_batched(scene_graph, map_idx, map_env, 1, include_mean=False)
        embed_info_attached = model.embed(scene_graph, map_idx, map_env)
    embed_info = detach_embed_info(embed_info_attached)
    init_future_pred = sample_pred['future_pred'][:, 0]
    cur_z = sample_pred['z_samp'][:, 0].clone().detach()
    cur_z.requires_grad = True
    opt = torch.optim.Adam([cur_z], lr=lr)
    avoid_loss = AvoidCollLoss(loss_weights, model.get_att_normalizer().unnormalize(scene_graph.lw), map_idx[scene_graph.batch],
                               map_env, cur_z.clone().detach(), veh_coll_buffer=0.2)
    for it in range(num_iters):
        opt.zero_grad()
        dec = model.decode_embedding(cur_z, embed_info, scene_graph, map_idx, map_env, nfuture=samp_future_len)
        fut = model.get_normalizer().unnormalize(dec['future_pred'])
        ld = avoid_loss(fut, cur_z, embed_info['prior_out'])
        if log is not None:
            log(it, {k: float(torch.mean(v)) for k, v in ld.items()})
        ld['loss'].backward()
        opt.step()
    with torch.no_grad():
        out = model.decode_embedding(cur_z, embed_info, scene_graph, map_idx, map_env, nfuture=save_future_len)
    return init_future_pred, cur_z, out['future_pred'].unsqueeze(1).clone().detach(), embed_info


class RefineLoop(object):
    """Device-resident refine iteration (decode -> AvoidCollLoss -> d/dz -> Adam)."""

    def __init__(self, model, scene_graph, map_idx, map_env, embed_info, z_init, loss_weights, lr, FT,
                 veh_coll_buffer=0.2, group_scene_ptr=None, betas=(0.9, 0.999), eps=1e-8):
        self.L = _cabi.lib()
        self.model = model
        self.dm = model.device_model()
        self.env = map_env
        dev = z_init.device
        self.scene = model.scene_batch(scene_graph, map_idx)
        NA = self.scene.NA
        self.NA, self.FT, self.lr, self.betas, self.eps = NA, int(FT), float(lr), betas, float(eps)
        self.z = z_init.detach().clone().contiguous().float()
        self.init_z = self.z.clone()
        self.map_feat = embed_info['map_feat'].detach().contiguous().float()
        self.past_feat = embed_info['past_feat'].detach().contiguous().float()
        self.prior_mu = embed_info['prior_out'][0].detach().contiguous().float()
        self.prior_var = embed_info['prior_out'][1].detach().contiguous().float()
        lw_un = model.get_att_normalizer().unnormalize(self.scene.lw)
        # refine builds AvoidCollLoss without ptr: one collision block per reference batch (= group)
        self.plan = LossPlan(_cabi.LOSS_AVOID, loss_weights, lw_un, self.scene.agent_map, map_env, self.scene.ptr_host, dev,
                             group_scene_ptr=group_scene_ptr, coll_by_scene=False, veh_coll_buffer=veh_coll_buffer,
                             traj_unnormalized=False)
        self.traj = torch.empty((NA, self.FT, 4), dtype=torch.float32, device=dev)
        self.d_traj = torch.empty_like(self.traj)
        self.d_z_bptt = torch.empty((NA, 32), dtype=torch.float32, device=dev)
        self.d_z_direct = torch.empty((NA, 32), dtype=torch.float32, device=dev)
        self.terms = torch.zeros((self.plan.G, _cabi.STRIVE_TERMS), dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.z)
        self.exp_avg_sq = torch.zeros_like(self.z)
        self.tape_bytes = self.L.strive_decode_tape_bytes(NA, self.FT)
        self.tape = torch.empty(self.tape_bytes, dtype=torch.uint8, device=dev)
        self.step_count = 0
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.graph = None
        # kernels launched per iteration (counted from the launch sequence in csrc/*.cu)
        self.launches_per_iter = self._count_launches()

    def _count_launches(self):
        FT, NA = self.FT, self.NA
        chunks = (NA + 2047) // 2048
        fwd = 1 + FT * 3 + (FT - 1) * (1 + chunks * 8)      # init_tape; node/edge/post per step; gru + (crop_pack, conv1..6, fc) per chunk
        bwd = FT * 3 + (FT - 1)
        loss = 7
        return fwd + bwd + loss + 1

    def _forward(self):
        _cabi.check(self.L.strive_decode_fwd(self.dm.handle, C.byref(self.scene.cstruct), C.byref(self.env.cstruct),
                                             _cabi.dptr(self.z), _cabi.dptr(self.map_feat), _cabi.dptr(self.past_feat), None,
                                             self.FT, _cabi.dptr(self.traj), _cabi.dptr(self.tape), self.tape_bytes,
                                             _cabi.stream_ptr()))

    def _loss(self):
        run_loss(self.plan, self.scene.cstruct, self.traj, self.z, self.prior_mu, self.prior_var, self.init_z,
                 d_traj=self.d_traj, d_z=self.d_z_direct, terms=self.terms)

    def _backward(self):
        _cabi.check(self.L.strive_decode_bwd(self.dm.handle, C.byref(self.scene.cstruct), self.FT, None, _cabi.dptr(self.d_traj),
                                             _cabi.dptr(self.d_z_bptt), _cabi.dptr(self.tape), self.tape_bytes,
                                             _cabi.stream_ptr()))

    def _adam(self):
        self.step_count += 1
        _cabi.check(self.L.strive_adam_step(_cabi.dptr(self.z), _cabi.dptr(self.d_z_bptt), _cabi.dptr(self.d_z_direct),
                                            _cabi.dptr(self.exp_avg), _cabi.dptr(self.exp_avg_sq), self.z.numel(),
                                            self.step_count, self.lr, self.betas[0], self.betas[1], self.eps,
                                            _cabi.stream_ptr()))

    def step(self):
        self._forward()
        self._loss()
        self._backward()
        self._adam()

    def run(self, iters):
        for _ in range(iters):
            self.step()
        return self.z

    def grad(self):
        """dL/dz of the last evaluated iteration (before Adam consumed it)."""
        return self.d_z_bptt + self.d_z_direct


# ----------------------------------------------------------------------------------------------------------
# init / adversarial / solution loops (reference src/utils/init_optim.py, adv_gen_optim.py, sol_optim.py)
# ----------------------------------------------------------------------------------------------------------
def collate_tgt_other_z(scene_graph, tgt_z, other_z):
    """reference adv_gen_optim.py:19-36, without the per-scene Python loop / O(B^2) concatenations: ego rows are
    scene_graph.ptr[:-1], everything else keeps graph order."""
    ptr = scene_graph.ptr
    NA = int(other_z.size(0) + tgt_z.size(0))
    ego = torch.zeros(NA, dtype=torch.bool, device=other_z.device)
    ego[ptr[:-1].long()] = True
    out = torch.empty((NA,) + tuple(other_z.shape[1:]), dtype=other_z.dtype, device=other_z.device)
    out[ego] = tgt_z
    out[~ego] = other_z
    return out


def _ego_mask(scene_graph, NA, device):
    m = torch.zeros(NA, dtype=torch.bool, device=device)
    m[scene_graph.ptr[:-1].long()] = True
    return m


def run_init_optim(cur_z, init_traj, traj_vis, lr, loss_weights, model, scene_graph, map_env, map_idx, num_iters, embed_info,
                   prior_distrib, log=None):
    """reference init_optim.py:11-68: fit z so the decoded future matches the observed one (TgtMatchingLoss with the
    init_* weights), Adam(lr)."""
    from .losses import TgtMatchingLoss
    init_traj = model.get_normalizer().unnormalize(init_traj)[traj_vis == 1.0]
    cur_z = cur_z.clone().detach()
    cur_z.requires_grad = True
    opt = torch.optim.Adam([cur_z], lr=lr)
    match_loss = TgtMatchingLoss({k[5:]: v for k, v in loss_weights.items() if k[:5] == 'init_'})
    for it in range(num_iters):
        opt.zero_grad()
        dec = model.decode_embedding(cur_z, embed_info, scene_graph, map_idx, map_env)
        fut = model.get_normalizer().unnormalize(dec['future_pred'])[traj_vis == 1.0]
        ld = match_loss(fut, init_traj, cur_z, prior_distrib)
        if log is not None:
            log(it, {k: float(torch.mean(v)) for k, v in ld.items()})
        ld['loss'].backward()
        opt.step()
    with torch.no_grad():
        out = model.decode_embedding(cur_z, embed_info, scene_graph, map_idx, map_env)
    return cur_z, out['future_pred'].clone().detach(), out


def run_adv_gen_optim(cur_z, lr, loss_weights, model, scene_graph, map_env, map_idx, num_iters, embed_info, planner_name,
                      tgt_prior_distrib, other_prior_distrib, feasibility_time, feasibility_infront_min, planner=None,
                      planner_viz_out=None, attack_agt_idx=None, future_len=None, veh_coll_buffer=0.1, log=None, debug=None):
    """reference adv_gen_optim.py:39-211, planner replay mode (planner_name == 'ego').

    The reference decodes twice per iteration with identical forward values (once with other_z detached for the target's
    matching loss, once with tgt_z detached for the adversarial loss, :119-130).  Here: ONE rollout and TWO adjoint sweeps
    over its tape (strive_decode_bwd with two seeds), which yields exactly the same gradients."""
    from .losses import TgtMatchingLoss, AdvGenLoss
    if planner_name != 'ego':
        raise RuntimeError('strive_b200: only planner="ego" (open-loop replay) is supported; the closed-loop rule-based planner '
                           '(adv_gen_optim.py:133-139) is CPU host code outside the scope of this port')
    NA = cur_z.size(0)
    dev = cur_z.device
    ego_mask = _ego_mask(scene_graph, NA, dev)
    ego_inds = scene_graph.ptr[:-1].long()
    if attack_agt_idx is not None:
        attack_agt_idx = torch.as_tensor(attack_agt_idx, device=dev).long() + ego_inds
    if future_len is None:
        future_len = model.FT
    tgt_z = cur_z[ego_mask].clone().detach().requires_grad_(True)
    other_z = cur_z[~ego_mask].clone().detach().requires_grad_(True)
    opt = torch.optim.Adam([tgt_z, other_z], lr=lr)
    nrm = model.get_normalizer()
    tgt_loss = TgtMatchingLoss(loss_weights)
    adv_loss = AdvGenLoss(loss_weights, model.get_att_normalizer().unnormalize(scene_graph.lw), map_idx[scene_graph.batch], map_env,
                          other_z.clone().detach(), scene_graph.ptr, veh_coll_buffer=veh_coll_buffer,
                          crash_loss_min_time=feasibility_time, crash_loss_min_infront=feasibility_infront_min)
    planner_fut = scene_graph.future_gt[ego_mask][:, :, :4]
    assert planner_fut.size(1) == future_len
    planner_un = nrm.unnormalize(planner_fut)
    for it in range(num_iters):
        opt.zero_grad()
        z_all = collate_tgt_other_z(scene_graph, tgt_z.detach(), other_z.detach()).requires_grad_(True)
        fut = model.decode_embedding(z_all, embed_info, scene_graph, map_idx, map_env, ext_future=planner_fut, nfuture=future_len)['future_pred']
        fut_un = nrm.unnormalize(fut)
        ld_t = tgt_loss(fut_un[ego_mask], planner_un, tgt_z, tgt_prior_distrib)
        ld_a = adv_loss(fut_un, planner_un, other_z, other_prior_distrib, attack_agt_idx=attack_agt_idx)
        g_t = torch.autograd.grad(ld_t['loss'], z_all, retain_graph=True)[0]          # adjoint sweep 1: target rows
        g_a, g_o = torch.autograd.grad(ld_a['loss'], [z_all, other_z])               # adjoint sweep 2 + direct latent terms
        tgt_z.grad = g_t[ego_mask]
        other_z.grad = g_a[~ego_mask] + g_o
        if debug is not None and it == 0:
            debug.update(g_tgt=tgt_z.grad.detach().clone(), g_other=other_z.grad.detach().clone())
        if log is not None:
            d = {'tgt_match_' + k: float(torch.mean(v)) for k, v in ld_t.items()}
            d.update({'adv_' + k: float(torch.mean(v)) for k, v in ld_a.items()})
            log(it, d)
        opt.step()
    cur_z = collate_tgt_other_z(scene_graph, tgt_z.detach(), other_z.detach())
    with torch.no_grad():
        final = model.decode_embedding(cur_z, embed_info, scene_graph, map_idx, map_env, nfuture=future_len)
    final_traj = final['future_pred'].unsqueeze(1).clone().detach()
    final_traj[ego_inds, 0] = planner_fut
    ld = adv_loss(nrm.unnormalize(final['future_pred']), nrm.unnormalize(final_traj[ego_inds, 0]), cur_z[~ego_mask].clone().detach(),
                  other_prior_distrib, return_mins=True)
    min_agt = ld['min_agt'] + ego_inds.cpu().numpy()
    return cur_z, final_traj, final, min_agt, ld['min_t']


def run_find_solution_optim(cur_z, final_result_traj, future_len, lr, loss_weights, model, scene_graph, map_env, map_idx,
                            num_iters, embed_info, tgt_prior_distrib, other_prior_distrib, log=None, debug=None):
    """reference sol_optim.py:19-123: the target (node 0 of every scene) avoids collisions (AvoidCollLoss, single_veh_idx=0,
    rollout of `future_len`) while the others keep matching the adversarial result (TgtMatchingLoss over model.FT steps).
    One rollout of max(future_len, FT) steps + two adjoint sweeps replaces the reference's two decodes (:73-77); the
    rollout is causal, so its first FT steps equal the shorter decode."""
    from .losses import AvoidCollLoss, TgtMatchingLoss
    NA = final_result_traj.size(0)
    dev = cur_z.device
    nrm = model.get_normalizer()
    tgt_mask = _ego_mask(scene_graph, NA, dev)
    other_match = nrm.unnormalize(final_result_traj[:, 0][~tgt_mask])            # (NA-B, FT, 4)
    FTm = other_match.size(1)
    tgt_z = tgt_prior_distrib[0].clone().detach().requires_grad_(True)             # (B, D)  sol_optim.py:38-40
    other_z = cur_z[~tgt_mask].reshape(NA - int(tgt_mask.sum()), -1).clone().detach().requires_grad_(True)
    opt = torch.optim.Adam([tgt_z, other_z], lr=lr)
    w = {k[4:]: v for k, v in loss_weights.items() if k[:4] == 'sol_'}
    avoid_loss = AvoidCollLoss(w, model.get_att_normalizer().unnormalize(scene_graph.lw), map_idx[scene_graph.batch], map_env,
                               tgt_z.clone().detach(), veh_coll_buffer=0.5, single_veh_idx=0, ptr=scene_graph.ptr)
    match_loss = TgtMatchingLoss(w)
    FTd = max(int(future_len), int(FTm))
    for it in range(num_iters):
        opt.zero_grad()
        z_all = collate_tgt_other_z(scene_graph, tgt_z.detach(), other_z.detach()).requires_grad_(True)
        fut = model.decode_embedding(z_all, embed_info, scene_graph, map_idx, map_env, nfuture=FTd)['future_pred']
        fut_un = nrm.unnormalize(fut)
        ld_t = avoid_loss(fut_un[:, :future_len].contiguous(), tgt_z, tgt_prior_distrib)
        ld_o = match_loss(fut_un[~tgt_mask][:, :FTm], other_match, other_z, other_prior_distrib)
        g_t, g_td = torch.autograd.grad(ld_t['loss'], [z_all, tgt_z], retain_graph=True)
        g_o = torch.autograd.grad(ld_o['loss'], z_all)[0]
        tgt_z.grad = g_t[tgt_mask] + g_td
        other_z.grad = g_o[~tgt_mask]
        if debug is not None and it == 0:
            debug.update(g_tgt=tgt_z.grad.detach().clone(), g_other=other_z.grad.detach().clone())
        if log is not None:
            d = {'tgt_' + k: float(torch.mean(v)) for k, v in ld_t.items()}
            d.update({'other_' + k: float(torch.mean(v)) for k, v in ld_o.items()})
            log(it, d)
        opt.step()
    cur_z = collate_tgt_other_z(scene_graph, tgt_z.detach(), other_z.detach())
    cur_z = cur_z.unsqueeze(1)                                                     # (NA,1,D) as the reference's 3-D z (sol_optim.py:38-44)
    with torch.no_grad():
        sol = model.decode_embedding(cur_z, embed_info, scene_graph, map_idx, map_env)    # future_pred (NA,1,FT,4), :118
    sol_traj = sol['future_pred'].clone().detach()
    sol_traj[~tgt_mask] = nrm.normalize(nrm.unnormalize(final_result_traj[~tgt_mask]))[:, :, :sol_traj.size(2)]   # :121, others keep the adversarial result
    return cur_z, sol_traj, sol


# ----------------------------------------------------------------------------------------------------------
# multi-GPU: one process per GPU, whole loss-normalisation groups per rank, no collective inside the loop
# ----------------------------------------------------------------------------------------------------------
class _DictGraph(object):
    pass


def refine_sharded(model, scene, map_env, loss_weights, iters, lr, FT, group_scene_ptr, veh_coll_buffer=0.2, dst=0):
    """Refines the latents of a whole batch across the ranks of the default process group.

    scene: dict of CPU tensors (ptr, past, lw, sem, map_idx, z, map_feat, past_feat, prior_mu, prior_var) describing the FULL
    batch, identical on every rank; group_scene_ptr: scene offsets of the loss-normalisation groups (the batches the reference
    driver would have formed).  Every rank runs a device-resident RefineLoop on the groups `shard.partition_groups` gives it;
    rank `dst` returns the refined (NA,32) latents in batch order (CPU), the other ranks return None."""
    import torch.distributed as dist
    from . import shard
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    costs = shard.group_costs(scene['ptr'], group_scene_ptr, FT)
    mine = shard.partition_groups(costs, world)[rank]
    sub, lgptr, agent_index = shard.shard_scenes(scene, group_scene_ptr, mine)
    NA = int(scene['ptr'][-1])
    if agent_index.numel() == 0:
        return shard.gather_rows(torch.zeros((0, scene['z'].size(1)), dtype=scene['z'].dtype), agent_index, NA, dst=dst)
    dev = map_env.device
    g = _DictGraph()
    for k in ('past', 'lw', 'sem', 'ptr', 'batch', 'edge_index'):
        setattr(g, k, sub[k].to(dev))
    embed = {'map_feat': sub['map_feat'].to(dev), 'past_feat': sub['past_feat'].to(dev),
             'prior_out': (sub['prior_mu'].to(dev), sub['prior_var'].to(dev))}
    loop = RefineLoop(model, g, sub['map_idx'].to(dev), map_env, embed, sub['z'].to(dev), loss_weights, lr, FT,
                      veh_coll_buffer=veh_coll_buffer, group_scene_ptr=lgptr)
    z = loop.run(iters)
    return shard.gather_rows(z, agent_index, NA, dst=dst)
