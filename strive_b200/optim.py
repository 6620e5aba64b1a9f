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
            log(it, {k: float(torch.mean(v)) for k, v in ld.items() if not k.startswith('_')})
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
        fwd = 1 + FT * 3 + (FT - 1) * (1 + chunks * 7)
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
