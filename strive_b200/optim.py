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


class _DeviceLoop(object):
    """Buffers and the four C-ABI calls of one device-resident latent iteration:
        strive_decode_fwd -> strive_loss_fwd_bwd -> strive_decode_bwd (one or two adjoint sweeps) -> strive_adam_step_dev.
    Nothing synchronises with the host; after the first (eager) iteration the launch sequence is captured once into a CUDA
    graph and replayed (use_graph=True): all buffers are preallocated and the Adam step count lives in device memory."""

    def __init__(self, model, scene_graph, map_idx, map_env, map_feat, past_feat, z_init, prior_mu, prior_var, lr, FT,
                 betas=(0.9, 0.999), eps=1e-8, ext_future=None, use_graph=True):
        self.L = _cabi.lib()
        self.model = model
        self.dm = model.device_model()
        self.env = map_env
        self.scene = model.scene_batch(scene_graph, map_idx)
        dev = self.scene.device
        self.dev = dev
        NA = self.scene.NA
        f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()
        self.NA, self.FT, self.lr, self.betas, self.eps = NA, int(FT), float(lr), betas, float(eps)
        self.z = f32(z_init).clone()
        if tuple(self.z.shape) != (NA, 32):
            raise RuntimeError('strive_b200: z must be (%d,32), got %s' % (NA, tuple(self.z.shape)))
        self.map_feat, self.past_feat = f32(map_feat), f32(past_feat)
        self.prior_mu, self.prior_var = f32(prior_mu), f32(prior_var)
        self.ext = None
        if ext_future is not None:
            self.ext = f32(ext_future[:, :self.FT, :4])
            if tuple(self.ext.shape) != (self.scene.S, self.FT, 4):
                raise RuntimeError('strive_b200: ext_future must be (%d,%d,4)' % (self.scene.S, self.FT))
        self.traj = torch.empty((NA, self.FT, 4), dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.z)
        self.exp_avg_sq = torch.zeros_like(self.z)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.tape_bytes = self.L.strive_decode_tape_bytes(NA, self.FT)
        self.tape = torch.empty(self.tape_bytes, dtype=torch.uint8, device=dev)
        self.use_graph = bool(use_graph)
        self.graph = None
        self.iterations = 0

    # ---- the four calls
    def _forward(self):
        _cabi.check(self.L.strive_decode_fwd(self.dm.handle, C.byref(self.scene.cstruct), C.byref(self.env.cstruct),
                                             _cabi.dptr(self.z), _cabi.dptr(self.map_feat), _cabi.dptr(self.past_feat),
                                             _cabi.dptr(self.ext), self.FT, _cabi.dptr(self.traj), _cabi.dptr(self.tape),
                                             self.tape_bytes, _cabi.stream_ptr()))

    def _sweep(self, d_traj, d_z):
        _cabi.check(self.L.strive_decode_bwd(self.dm.handle, C.byref(self.scene.cstruct), self.FT, _cabi.dptr(self.ext),
                                             _cabi.dptr(d_traj), _cabi.dptr(d_z), _cabi.dptr(self.tape), self.tape_bytes,
                                             _cabi.stream_ptr()))

    def _sweep_pair(self, d_traj_a, d_z_a, d_traj_b, d_z_b):
        """Both adjoint sweeps of an adv / sol iteration in one call: two parallel branches on the device."""
        _cabi.check(self.L.strive_decode_bwd_pair(self.dm.handle, C.byref(self.scene.cstruct), self.FT, _cabi.dptr(self.ext),
                                                  _cabi.dptr(d_traj_a), _cabi.dptr(d_z_a), _cabi.dptr(d_traj_b), _cabi.dptr(d_z_b),
                                                  _cabi.dptr(self.tape), self.tape_bytes, _cabi.stream_ptr()))

    def _adam_dev(self, g_a, g_b=None, g_direct=None, row_sel=None):
        _cabi.check(self.L.strive_adam_step_dev(_cabi.dptr(self.z), _cabi.dptr(g_a), _cabi.dptr(g_b), _cabi.dptr(g_direct),
                                                _cabi.dptr(row_sel), 32, _cabi.dptr(self.exp_avg), _cabi.dptr(self.exp_avg_sq),
                                                self.z.numel(), _cabi.dptr(self.step_dev), self.lr, self.betas[0],
                                                self.betas[1], self.eps, _cabi.stream_ptr()))

    def _iteration(self):
        raise NotImplementedError

    # ---- driving
    def step(self):
        """One iteration.  The first one runs eagerly (lazy workspaces, function attributes); it is then captured and every
        further call replays the graph."""
        with torch.cuda.device(self.dev):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._iteration()
                if self.use_graph and self.iterations == 0:
                    self._capture()
        self.iterations += 1

    def eager_step(self):
        """One iteration launched kernel by kernel even when a captured graph exists (per-kernel event profiling)."""
        with torch.cuda.device(self.dev):
            self._iteration()
        self.iterations += 1

    def _capture(self):
        torch.cuda.synchronize(self.dev)
        # the capture must not disturb the state the eager iteration left: a capture records launches without running them
        g = torch.cuda.CUDAGraph()
        # an explicit capture stream on THIS loop's device: torch.cuda.graph's default capture stream is one class-wide stream on the
        # device of the first capture of the process, and entering it would switch the current device away from the loop's tensors
        with torch.cuda.graph(g, stream=torch.cuda.Stream(device=self.dev)):
            self._iteration()
        self.graph = g

    def run(self, iters):
        for _ in range(int(iters)):
            self.step()
        return self.z

    def _rollout_launches(self, sweeps):
        FT, NA = self.FT, self.NA
        chunks = (NA + 2047) // 2048
        tiles = 1 if self.scene.max_n <= 129 else 0         # edge_tc_tiles (tcgen05 edge kernels: scenes up to 129 agents)
        fwd = 1 + tiles + FT * 3 + (FT - 1) * (1 + chunks * 8)      # init_tape; node/edge/post per step; gru + (crop_pack, conv1..6, fc) per chunk
        bwd = tiles + FT * 3 + (FT - 1)
        return fwd + sweeps * bwd


class RefineLoop(_DeviceLoop):
    """Device-resident refine iteration (decode -> AvoidCollLoss -> d/dz -> Adam), reference refine_traffic_optim.py:185-218."""

    def __init__(self, model, scene_graph, map_idx, map_env, embed_info, z_init, loss_weights, lr, FT,
                 veh_coll_buffer=0.2, group_scene_ptr=None, betas=(0.9, 0.999), eps=1e-8, use_graph=True):
        super().__init__(model, scene_graph, map_idx, map_env, embed_info['map_feat'], embed_info['past_feat'], z_init,
                         embed_info['prior_out'][0], embed_info['prior_out'][1], lr, FT, betas=betas, eps=eps, use_graph=use_graph)
        dev, NA = self.dev, self.NA
        self.init_z = self.z.clone()
        lw_un = model.get_att_normalizer().unnormalize(self.scene.lw)
        # refine builds AvoidCollLoss without ptr: one collision block per reference batch (= group)
        self.plan = LossPlan(_cabi.LOSS_AVOID, loss_weights, lw_un, self.scene.agent_map, map_env, self.scene.ptr_host, dev,
                             group_scene_ptr=group_scene_ptr, coll_by_scene=False, veh_coll_buffer=veh_coll_buffer,
                             traj_unnormalized=False)
        self.d_traj = torch.empty_like(self.traj)
        self.d_z_bptt = torch.empty((NA, 32), dtype=torch.float32, device=dev)
        self.d_z_direct = torch.empty((NA, 32), dtype=torch.float32, device=dev)
        self.terms = torch.zeros((self.plan.G, _cabi.STRIVE_TERMS), dtype=torch.float32, device=dev)
        # kernels launched per iteration (counted from the launch sequence in csrc/*.cu)
        self.launches_per_iter = self._rollout_launches(1) + 7 + 2

    def _loss(self):
        run_loss(self.plan, self.scene.cstruct, self.traj, self.z, self.prior_mu, self.prior_var, self.init_z,
                 d_traj=self.d_traj, d_z=self.d_z_direct, terms=self.terms)

    def _backward(self):
        self._sweep(self.d_traj, self.d_z_bptt)

    def _adam(self):
        self._adam_dev(self.d_z_bptt, g_direct=self.d_z_direct)

    def _iteration(self):
        self._forward()
        self._loss()
        self._backward()
        self._adam()

    def grad(self):
        """dL/dz of the last evaluated iteration (before Adam consumed it)."""
        return self.d_z_bptt + self.d_z_direct


class InitLoop(_DeviceLoop):
    """Device-resident form of utils/init_optim.py:30-58: z is fitted so the decoded future matches the observed one at the
    visible (agent, step) entries (TgtMatchingLoss with the init_* weights, incl. its :46 quirk)."""

    def __init__(self, model, scene_graph, map_idx, map_env, embed_info, z_init, init_traj, traj_vis, loss_weights, lr, FT=None,
                 prior=None, group_scene_ptr=None, use_graph=True):
        FT = model.FT if FT is None else int(FT)
        NA = z_init.size(0)
        dev = z_init.device
        mu = prior[0] if prior is not None else torch.zeros((NA, 32), device=dev)
        var = prior[1] if prior is not None else torch.ones((NA, 32), device=dev)
        super().__init__(model, scene_graph, map_idx, map_env, embed_info['map_feat'], embed_info['past_feat'], z_init, mu, var, lr, FT,
                         use_graph=use_graph)
        w = {k[5:]: v for k, v in loss_weights.items() if k[:5] == 'init_'}
        lw_un = model.get_att_normalizer().unnormalize(self.scene.lw)
        vis = (traj_vis[:, :FT] == 1.0)
        self.plan = LossPlan(_cabi.LOSS_MATCH, w, lw_un, self.scene.agent_map, map_env, self.scene.ptr_host, self.dev,
                             group_scene_ptr=group_scene_ptr, traj_unnormalized=False, match_mask=vis)
        self.match_tgt = init_traj[:, :FT, :4].detach().to(self.dev, torch.float32).contiguous()       # NORMALISED, as the rollout
        self.d_traj = torch.empty_like(self.traj)
        self.g = torch.empty((self.NA, 32), dtype=torch.float32, device=self.dev)
        self.terms = torch.zeros((self.plan.G, _cabi.STRIVE_TERMS), dtype=torch.float32, device=self.dev)
        self.launches_per_iter = self._rollout_launches(1) + 4 + 2

    def _iteration(self):
        self._forward()
        run_loss(self.plan, self.scene.cstruct, self.traj, None, None, None, None, match_tgt=self.match_tgt,
                 d_traj_match=self.d_traj, terms=self.terms)
        self._sweep(self.d_traj, self.g)
        self._adam_dev(self.g)

    def log_dict(self):
        t = self.terms.sum(dim=0).tolist()
        return {'match_ext_loss': t[10], 'loss': t[11]}


class AdvLoop(_DeviceLoop):
    """Device-resident form of utils/adv_gen_optim.py:106-175 in planner-replay mode: ONE rollout (the planner's future injected
    for the ego rows), TgtMatchingLoss on the ego rows + AdvGenLoss on everything else in one fused loss call, TWO adjoint
    sweeps over the same tape (the reference decodes twice with identical values and different detach masks, :119-130), and
    one Adam step over [tgt_z ; other_z] held as a single (NA,32) buffer in graph order (= collate_tgt_other_z)."""

    def __init__(self, model, scene_graph, map_idx, map_env, embed_info, z_init, planner_fut, loss_weights, lr, FT, prior,
                 veh_coll_buffer=0.1, crash_min_t=0, crash_min_infront=None, attack_mask=None, group_scene_ptr=None, use_graph=True):
        super().__init__(model, scene_graph, map_idx, map_env, embed_info['map_feat'], embed_info['past_feat'], z_init, prior[0], prior[1],
                         lr, FT, ext_future=planner_fut, use_graph=use_graph)
        dev, NA, FT = self.dev, self.NA, self.FT
        ego = self.scene.ego_mask
        self.ego_u8 = ego.to(torch.uint8).contiguous()
        self.init_z = self.z.clone()
        lw_un = model.get_att_normalizer().unnormalize(self.scene.lw)
        mm = torch.zeros((NA, FT), dtype=torch.bool, device=dev)
        mm[ego] = True
        self.plan = LossPlan(_cabi.LOSS_ADV | _cabi.LOSS_MATCH, loss_weights, lw_un, self.scene.agent_map, map_env, self.scene.ptr_host, dev,
                             group_scene_ptr=group_scene_ptr, coll_by_scene=True, veh_coll_buffer=veh_coll_buffer, crash_min_t=crash_min_t,
                             crash_min_infront=crash_min_infront, traj_unnormalized=False, match_mask=mm)
        if attack_mask is not None:
            self.plan.set_attack_mask(attack_mask)
        self.match_tgt = torch.zeros((NA, FT, 4), dtype=torch.float32, device=dev)
        self.match_tgt[ego] = self.ext
        self.d_traj = torch.empty_like(self.traj)
        self.d_traj_match = torch.empty_like(self.traj)
        self.g_tgt = torch.empty((NA, 32), dtype=torch.float32, device=dev)
        self.g_oth = torch.empty((NA, 32), dtype=torch.float32, device=dev)
        self.d_z_direct = torch.empty((NA, 32), dtype=torch.float32, device=dev)
        self.terms = torch.zeros((self.plan.G, _cabi.STRIVE_TERMS), dtype=torch.float32, device=dev)
        self.launches_per_iter = self._rollout_launches(2) + 10 + 2

    def _iteration(self):
        self._forward()
        run_loss(self.plan, self.scene.cstruct, self.traj, self.z, self.prior_mu, self.prior_var, self.init_z, match_tgt=self.match_tgt,
                 adv_tgt=self.ext, d_traj=self.d_traj, d_traj_match=self.d_traj_match, d_z=self.d_z_direct, terms=self.terms)
        # target rows: matching loss | all other rows: adversarial loss
        self._sweep_pair(self.d_traj_match, self.g_tgt, self.d_traj, self.g_oth)
        self._adam_dev(self.g_tgt, g_b=self.g_oth, g_direct=self.d_z_direct, row_sel=self.ego_u8)

    def grads(self):
        ego = self.scene.ego_mask
        return self.g_tgt[ego], (self.g_oth + self.d_z_direct)[~ego]

    def log_dict(self):
        t = self.terms.sum(dim=0).tolist() if self.plan.G == 1 else self.terms.mean(dim=0).tolist()
        return {'tgt_match_match_ext_loss': t[10], 'tgt_match_loss': t[11], 'adv_init_loss': t[6], 'adv_motion_prior_loss': t[5],
                'adv_coll_veh_loss': t[1], 'adv_coll_veh_plan_loss': t[7], 'adv_coll_env_loss': t[3], 'adv_adv_crash_loss': t[9],
                'adv_loss': t[0]}


class AdvClosedLoop(_DeviceLoop):
    """Closed-loop form of the adversarial iteration (planner_name == 'hardcode', adv_gen_optim.py:98-154): no planner future is
    injected into the rollout; every iteration the CPU planner reacts to the current prediction of the other agents
    (`planner.rollout`, :133-139), the target's latent is fitted to the planner's answer (TgtMatchingLoss) and the adversarial
    loss attacks the model's OWN prediction of the target, whose gradient reaches the other latents through the interaction
    net (:143-154, `adv_own_pred`).  One rollout, two sweeps.  The planner round trip overlaps the adversarial loss and its
    adjoint sweep: the predicted futures leave on a side stream right after the rollout, the host runs the planner while the
    device works, and only the matching loss + first sweep wait for the planner's answer.  `planner_ms` / `iter_ms` accumulate
    host wall-clock of the planner call and of the whole iteration (planner time reported separately)."""

    def __init__(self, model, scene_graph, map_idx, map_env, embed_info, z_init, planner, loss_weights, lr, FT, prior,
                 veh_coll_buffer=0.1, crash_min_t=0, crash_min_infront=None, attack_mask=None, group_scene_ptr=None):
        super().__init__(model, scene_graph, map_idx, map_env, embed_info['map_feat'], embed_info['past_feat'], z_init, prior[0], prior[1],
                         lr, FT, use_graph=False)
        import numpy as np
        dev, NA, FT = self.dev, self.NA, self.FT
        ego = self.scene.ego_mask
        self.ego = ego
        self.ego_u8 = ego.to(torch.uint8).contiguous()
        self.init_z = self.z.clone()
        self.planner = planner
        self.nrm = model.get_normalizer()
        lw_un = model.get_att_normalizer().unnormalize(self.scene.lw)
        mm = torch.zeros((NA, FT), dtype=torch.bool, device=dev)
        mm[ego] = True
        common = dict(group_scene_ptr=group_scene_ptr, traj_unnormalized=False)
        self.plan_adv = LossPlan(_cabi.LOSS_ADV, loss_weights, lw_un, self.scene.agent_map, map_env, self.scene.ptr_host, dev, coll_by_scene=True,
                                 veh_coll_buffer=veh_coll_buffer, crash_min_t=crash_min_t, crash_min_infront=crash_min_infront, **common)
        self.plan_adv.cfg.adv_own_pred = 1
        if attack_mask is not None:
            self.plan_adv.set_attack_mask(attack_mask)
        self.plan_match = LossPlan(_cabi.LOSS_MATCH, loss_weights, lw_un, self.scene.agent_map, map_env, self.scene.ptr_host, dev, match_mask=mm, **common)
        self.match_tgt = torch.zeros((NA, FT, 4), dtype=torch.float32, device=dev)
        self.d_traj = torch.empty_like(self.traj)
        self.d_traj_match = torch.empty_like(self.traj)
        self.g_tgt = torch.empty((NA, 32), dtype=torch.float32, device=dev)
        self.g_oth = torch.empty((NA, 32), dtype=torch.float32, device=dev)
        self.d_z_direct = torch.empty((NA, 32), dtype=torch.float32, device=dev)
        self.terms = torch.zeros((self.plan_adv.G, _cabi.STRIVE_TERMS), dtype=torch.float32, device=dev)
        self.terms_m = torch.zeros((self.plan_adv.G, _cabi.STRIVE_TERMS), dtype=torch.float32, device=dev)
        self.other_idx = torch.nonzero(~ego, as_tuple=False).flatten()
        self.ego_idx = torch.nonzero(ego, as_tuple=False).flatten()
        self.h_pred = torch.empty((NA - self.scene.S, FT, 4), dtype=torch.float32).pin_memory()
        self.h_plan = torch.empty((self.scene.S, FT, 4), dtype=torch.float32).pin_memory()
        self.side = torch.cuda.Stream(device=dev)
        self.ev_fwd = torch.cuda.Event()
        self.ev_copy = torch.cuda.Event()
        ptr = self.scene.ptr_host
        self.agt_ptr = (ptr - torch.arange(ptr.numel())).numpy()                 # cur_agt_ptr, :103
        self.plan_t = np.linspace(model.dt, model.dt * FT, FT)                    # :104
        self.planner_ms, self.iter_ms = 0.0, 0.0
        self.launches_per_iter = self._rollout_launches(2) + 9 + 4 + 2

    def plan(self, final=False, viz=None):
        """planner.rollout on the others' current prediction (self.traj) -> normalised (S,FT,4) device tensor (:133-139; the final
        call passes viz= as the reference does, :186-193)."""
        import time
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ev_fwd)
            pred_un = self.nrm.unnormalize(self.traj.index_select(0, self.other_idx))
            self.h_pred.copy_(pred_un, non_blocking=True)
            self.ev_copy.record(self.side)
        self.ev_copy.synchronize()
        t0 = time.perf_counter()
        if final:
            fut = self.planner.rollout(self.h_pred.numpy(), self.plan_t, self.agt_ptr, self.plan_t, viz=viz, control_all=False)
        else:
            fut = self.planner.rollout(self.h_pred.numpy(), self.plan_t, self.agt_ptr, self.plan_t, control_all=False)
        self.planner_ms += 1000.0 * (time.perf_counter() - t0)
        self.h_plan.copy_(torch.as_tensor(fut, dtype=torch.float32)[:, :, :4])
        return self.nrm.normalize(self.h_plan.to(self.dev, non_blocking=True))

    def _iteration(self):
        import time
        t0 = time.perf_counter()
        self._forward()
        self.ev_fwd.record()
        # adversarial loss on the model's own target prediction + its adjoint sweep: queued before the host touches the planner
        run_loss(self.plan_adv, self.scene.cstruct, self.traj, self.z, self.prior_mu, self.prior_var, self.init_z,
                 d_traj=self.d_traj, d_z=self.d_z_direct, terms=self.terms)
        self._sweep(self.d_traj, self.g_oth)
        planner_fut = self.plan()                                                 # host: overlaps the work queued above
        self.match_tgt.index_copy_(0, self.ego_idx, planner_fut)
        run_loss(self.plan_match, self.scene.cstruct, self.traj, None, None, None, None, match_tgt=self.match_tgt,
                 d_traj_match=self.d_traj_match, terms=self.terms_m)
        self._sweep(self.d_traj_match, self.g_tgt)
        self._adam_dev(self.g_tgt, g_b=self.g_oth, g_direct=self.d_z_direct, row_sel=self.ego_u8)
        self.iter_ms += 1000.0 * (time.perf_counter() - t0)

    def grads(self):
        return self.g_tgt[self.ego], (self.g_oth + self.d_z_direct)[~self.ego]

    def log_dict(self):
        t = self.terms.sum(dim=0).tolist() if self.plan_adv.G == 1 else self.terms.mean(dim=0).tolist()
        m = self.terms_m.sum(dim=0).tolist() if self.plan_adv.G == 1 else self.terms_m.mean(dim=0).tolist()
        return {'tgt_match_match_ext_loss': m[10], 'tgt_match_loss': m[11], 'adv_init_loss': t[6], 'adv_motion_prior_loss': t[5],
                'adv_coll_veh_loss': t[1], 'adv_coll_veh_plan_loss': t[7], 'adv_coll_env_loss': t[3], 'adv_adv_crash_loss': t[9],
                'adv_loss': t[0]}


class SolLoop(_DeviceLoop):
    """Device-resident form of utils/sol_optim.py:68-112: the target (node 0 of every scene, latent initialised at its prior
    mean) avoids collisions (AvoidCollLoss, single_veh_idx=0, buffer 0.5) over `future_len` steps while every other agent
    keeps matching the adversarial result over its first FTm steps; one rollout of `future_len` (>= FTm) steps, two sweeps."""

    def __init__(self, model, scene_graph, map_idx, map_env, embed_info, z_init, other_match_n, loss_weights, lr, future_len, prior,
                 group_scene_ptr=None, use_graph=True):
        FTm = int(other_match_n.size(1))
        if int(future_len) < FTm:
            raise RuntimeError('strive_b200: SolLoop needs future_len >= the matched horizon (%d < %d)' % (int(future_len), FTm))
        super().__init__(model, scene_graph, map_idx, map_env, embed_info['map_feat'], embed_info['past_feat'], z_init, prior[0], prior[1],
                         lr, int(future_len), use_graph=use_graph)
        dev, NA, FT = self.dev, self.NA, self.FT
        ego = self.scene.ego_mask
        self.ego_u8 = ego.to(torch.uint8).contiguous()
        self.z[ego] = self.prior_mu[ego]                       # sol_optim.py:38-40
        self.init_z = self.z.clone()
        lw_un = model.get_att_normalizer().unnormalize(self.scene.lw)
        mm = torch.zeros((NA, FT), dtype=torch.bool, device=dev)
        mm[~ego, :FTm] = True
        self.plan = LossPlan(_cabi.LOSS_AVOID | _cabi.LOSS_MATCH, loss_weights, lw_un, self.scene.agent_map, map_env, self.scene.ptr_host, dev,
                             group_scene_ptr=group_scene_ptr, coll_by_scene=True, veh_coll_buffer=0.5, single_veh_idx=0,
                             traj_unnormalized=False, match_mask=mm)
        self.match_tgt = torch.zeros((NA, FT, 4), dtype=torch.float32, device=dev)
        self.match_tgt[~ego, :FTm] = other_match_n.detach().to(dev, torch.float32)
        self.d_traj = torch.empty_like(self.traj)
        self.d_traj_match = torch.empty_like(self.traj)
        self.g_tgt = torch.empty((NA, 32), dtype=torch.float32, device=dev)
        self.g_oth = torch.empty((NA, 32), dtype=torch.float32, device=dev)
        self.d_z_direct = torch.empty((NA, 32), dtype=torch.float32, device=dev)
        self.terms = torch.zeros((self.plan.G, _cabi.STRIVE_TERMS), dtype=torch.float32, device=dev)
        self.launches_per_iter = self._rollout_launches(2) + 8 + 2

    def _iteration(self):
        self._forward()
        run_loss(self.plan, self.scene.cstruct, self.traj, self.z, self.prior_mu, self.prior_var, self.init_z, match_tgt=self.match_tgt,
                 d_traj=self.d_traj, d_traj_match=self.d_traj_match, d_z=self.d_z_direct, terms=self.terms)
        # target rows: avoid-collision loss | other rows: matching loss
        self._sweep_pair(self.d_traj, self.g_tgt, self.d_traj_match, self.g_oth)
        self._adam_dev(self.g_tgt, g_b=self.g_oth, g_direct=self.d_z_direct, row_sel=self.ego_u8)

    def grads(self):
        ego = self.scene.ego_mask
        return (self.g_tgt + self.d_z_direct)[ego], self.g_oth[~ego]

    def log_dict(self):
        t = self.terms.sum(dim=0).tolist() if self.plan.G == 1 else self.terms.mean(dim=0).tolist()
        return {'tgt_coll_veh_loss': t[1], 'tgt_coll_env_loss': t[3], 'tgt_motion_prior_loss': t[5], 'tgt_init_loss': t[6], 'tgt_loss': t[0],
                'other_match_ext_loss': t[10], 'other_loss': t[11]}


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
                   prior_distrib, log=None, fused=True):
    """reference init_optim.py:11-68: fit z so the decoded future matches the observed one (TgtMatchingLoss with the
    init_* weights), Adam(lr).  fused=True (default) runs the device-resident InitLoop (no host synchronisation unless `log`
    is given); fused=False runs the same iteration through the drop-in modules + autograd + torch.optim.Adam."""
    from .losses import TgtMatchingLoss
    if fused:
        loop = InitLoop(model, scene_graph, map_idx, map_env, embed_info, cur_z, init_traj, traj_vis, loss_weights, lr, prior=prior_distrib)
        for it in range(num_iters):
            loop.step()
            if log is not None:
                log(it, loop.log_dict())
        cur_z = loop.z.clone().requires_grad_(True)
    else:
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


def _full_rows(NA, mask, rows_true, rows_false):
    out = torch.empty((NA,) + tuple(rows_true.shape[1:]), dtype=rows_true.dtype, device=rows_true.device)
    out[mask] = rows_true
    out[~mask] = rows_false
    return out


def run_adv_gen_optim(cur_z, lr, loss_weights, model, scene_graph, map_env, map_idx, num_iters, embed_info, planner_name,
                      tgt_prior_distrib, other_prior_distrib, feasibility_time, feasibility_infront_min, planner=None,
                      planner_viz_out=None, attack_agt_idx=None, future_len=None, veh_coll_buffer=0.1, log=None, debug=None, fused=True):
    """reference adv_gen_optim.py:39-211, planner replay mode (planner_name == 'ego').

    The reference decodes twice per iteration with identical forward values (once with other_z detached for the target's
    matching loss, once with tgt_z detached for the adversarial loss, :119-130).  Here: ONE rollout and TWO adjoint sweeps
    over its tape (strive_decode_bwd with two seeds), which yields exactly the same gradients.  fused=True (default) runs the
    device-resident AdvLoop; fused=False the same iteration through the drop-in modules + autograd + torch.optim.Adam."""
    from .losses import TgtMatchingLoss, AdvGenLoss
    if planner_name not in ('ego', 'hardcode'):
        raise RuntimeError('strive_b200: planner_name must be "ego" (open-loop replay) or "hardcode" (closed loop, adv_gen_optim.py:98-104)')
    if planner_name == 'hardcode':
        return _run_adv_gen_closed_loop(cur_z, lr, loss_weights, model, scene_graph, map_env, map_idx, num_iters, embed_info, tgt_prior_distrib,
                                        other_prior_distrib, feasibility_time, feasibility_infront_min, planner, planner_viz_out, attack_agt_idx,
                                        future_len, veh_coll_buffer, log, debug, fused)
    NA = cur_z.size(0)
    dev = cur_z.device
    ego_mask = _ego_mask(scene_graph, NA, dev)
    ego_inds = scene_graph.ptr[:-1].long()
    if attack_agt_idx is not None:
        attack_agt_idx = torch.as_tensor(attack_agt_idx, device=dev).long() + ego_inds
    if future_len is None:
        future_len = model.FT
    nrm = model.get_normalizer()
    planner_fut = scene_graph.future_gt[ego_mask][:, :, :4]
    assert planner_fut.size(1) == future_len
    planner_un = nrm.unnormalize(planner_fut)
    adv_loss = AdvGenLoss(loss_weights, model.get_att_normalizer().unnormalize(scene_graph.lw), map_idx[scene_graph.batch], map_env,
                          cur_z[~ego_mask].clone().detach(), scene_graph.ptr, veh_coll_buffer=veh_coll_buffer,
                          crash_loss_min_time=feasibility_time, crash_loss_min_infront=feasibility_infront_min)
    if fused:
        atk = None
        if attack_agt_idx is not None:
            atk = torch.zeros(NA, dtype=torch.int32, device=dev)
            atk[attack_agt_idx] = 1
        prior = (_full_rows(NA, ego_mask, tgt_prior_distrib[0], other_prior_distrib[0]),
                 _full_rows(NA, ego_mask, tgt_prior_distrib[1], other_prior_distrib[1]))
        loop = AdvLoop(model, scene_graph, map_idx, map_env, embed_info, cur_z, planner_fut, loss_weights, lr, future_len, prior,
                       veh_coll_buffer=veh_coll_buffer, crash_min_t=feasibility_time, crash_min_infront=feasibility_infront_min, attack_mask=atk)
        for it in range(num_iters):
            loop.step()
            if debug is not None and it == 0:
                gt, go = loop.grads()
                debug.update(g_tgt=gt.clone(), g_other=go.clone())
            if log is not None:
                log(it, loop.log_dict())
        cur_z = loop.z.clone()
    else:
        tgt_z = cur_z[ego_mask].clone().detach().requires_grad_(True)
        other_z = cur_z[~ego_mask].clone().detach().requires_grad_(True)
        opt = torch.optim.Adam([tgt_z, other_z], lr=lr)
        tgt_loss = TgtMatchingLoss(loss_weights)
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


def _run_adv_gen_closed_loop(cur_z, lr, loss_weights, model, scene_graph, map_env, map_idx, num_iters, embed_info, tgt_prior_distrib,
                             other_prior_distrib, feasibility_time, feasibility_infront_min, planner, planner_viz_out, attack_agt_idx,
                             future_len, veh_coll_buffer, log, debug, fused):
    """planner_name == 'hardcode' (adv_gen_optim.py:85-104, 133-154, 186-207): `planner` is any object with the reference planner's
    reset(init_state, veh_att, batch, B, map_idx) / rollout(agent_obs, agent_t, agent_ptr, planner_t, control_all=False[, viz=]) calls
    (src/planners/hardcode_goalcond_nusc.py:109,178); it runs on the host as in the reference."""
    from .losses import TgtMatchingLoss, AdvGenLoss
    if planner is None:
        raise RuntimeError('strive_b200: planner_name="hardcode" needs a planner object')
    NA = cur_z.size(0)
    dev = cur_z.device
    B = int(map_idx.size(0))
    ego_mask = _ego_mask(scene_graph, NA, dev)
    ego_inds = scene_graph.ptr[:-1].long()
    if attack_agt_idx is not None:
        attack_agt_idx = torch.as_tensor(attack_agt_idx, device=dev).long() + ego_inds
    if future_len is None:
        future_len = model.FT
    nrm = model.get_normalizer()
    lw_un = model.get_att_normalizer().unnormalize(scene_graph.lw)
    planner.reset(nrm.unnormalize(scene_graph.past_gt[:, -1, :]), lw_un, scene_graph.batch, B, map_idx)        # :85-89
    adv_loss = AdvGenLoss(loss_weights, lw_un, map_idx[scene_graph.batch], map_env, cur_z[~ego_mask].clone().detach(), scene_graph.ptr,
                          veh_coll_buffer=veh_coll_buffer, crash_loss_min_time=feasibility_time, crash_loss_min_infront=feasibility_infront_min)
    atk = None
    if attack_agt_idx is not None:
        atk = torch.zeros(NA, dtype=torch.int32, device=dev)
        atk[attack_agt_idx] = 1
    prior = (_full_rows(NA, ego_mask, tgt_prior_distrib[0], other_prior_distrib[0]),
             _full_rows(NA, ego_mask, tgt_prior_distrib[1], other_prior_distrib[1]))
    loop = AdvClosedLoop(model, scene_graph, map_idx, map_env, embed_info, cur_z, planner, loss_weights, lr, future_len, prior,
                         veh_coll_buffer=veh_coll_buffer, crash_min_t=feasibility_time, crash_min_infront=feasibility_infront_min, attack_mask=atk)
    if fused:
        for it in range(num_iters):
            loop.step()
            if debug is not None and it == 0:
                gt, go = loop.grads()
                debug.update(g_tgt=gt.clone(), g_other=go.clone())
            if log is not None:
                log(it, loop.log_dict())
        cur_z = loop.z.clone()
    else:
        # the same iteration through the drop-in modules + autograd (AdvGenLoss with a differentiable tgt_traj)
        tgt_z = cur_z[ego_mask].clone().detach().requires_grad_(True)
        other_z = cur_z[~ego_mask].clone().detach().requires_grad_(True)
        opt = torch.optim.Adam([tgt_z, other_z], lr=lr)
        tgt_loss = TgtMatchingLoss(loss_weights)
        for it in range(num_iters):
            opt.zero_grad()
            z_all = collate_tgt_other_z(scene_graph, tgt_z.detach(), other_z.detach()).requires_grad_(True)
            fut = model.decode_embedding(z_all, embed_info, scene_graph, map_idx, map_env, nfuture=future_len)['future_pred']
            loop.traj.copy_(fut.detach())
            loop.ev_fwd.record()
            planner_fut = loop.plan()
            fut_un = nrm.unnormalize(fut)
            ld_t = tgt_loss(fut_un[ego_mask], nrm.unnormalize(planner_fut), tgt_z, tgt_prior_distrib)
            ld_a = adv_loss(fut_un, fut_un[ego_mask], other_z, other_prior_distrib, attack_agt_idx=attack_agt_idx)
            g_t = torch.autograd.grad(ld_t['loss'], z_all, retain_graph=True)[0]
            g_a, g_o = torch.autograd.grad(ld_a['loss'], [z_all, other_z])
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
    loop.traj.copy_(final['future_pred'])
    loop.ev_fwd.record()
    planner_fut = loop.plan(final=True, viz=planner_viz_out)                         # :186-193, the planner's true reaction
    final_traj = final['future_pred'].unsqueeze(1).clone().detach()
    final_traj[ego_inds, 0] = planner_fut
    ld = adv_loss(nrm.unnormalize(final['future_pred']), nrm.unnormalize(final_traj[ego_inds, 0]), cur_z[~ego_mask].clone().detach(),
                  other_prior_distrib, return_mins=True)
    min_agt = ld['min_agt'] + ego_inds.cpu().numpy()
    if debug is not None:
        debug.update(planner_ms=loop.planner_ms, iter_ms=loop.iter_ms)
    return cur_z, final_traj, final, min_agt, ld['min_t']


def run_find_solution_optim(cur_z, final_result_traj, future_len, lr, loss_weights, model, scene_graph, map_env, map_idx,
                            num_iters, embed_info, tgt_prior_distrib, other_prior_distrib, log=None, debug=None, fused=True):
    """reference sol_optim.py:19-123: the target (node 0 of every scene) avoids collisions (AvoidCollLoss, single_veh_idx=0,
    rollout of `future_len`) while the others keep matching the adversarial result (TgtMatchingLoss over model.FT steps).
    One rollout of max(future_len, FT) steps + two adjoint sweeps replaces the reference's two decodes (:73-77); the
    rollout is causal, so its first FT steps equal the shorter decode.  fused=True (default) runs the device-resident SolLoop
    (needs future_len >= FT, as in configs/adv_gen_rule_based.cfg: 16 >= 12); fused=False the drop-in modules + autograd."""
    from .losses import AvoidCollLoss, TgtMatchingLoss
    NA = final_result_traj.size(0)
    dev = cur_z.device
    nrm = model.get_normalizer()
    tgt_mask = _ego_mask(scene_graph, NA, dev)
    other_match = nrm.unnormalize(final_result_traj[:, 0][~tgt_mask])            # (NA-B, FT, 4)
    FTm = other_match.size(1)
    w = {k[4:]: v for k, v in loss_weights.items() if k[:4] == 'sol_'}
    cur_z2 = cur_z.reshape(NA, -1)
    if fused and int(future_len) >= int(FTm):
        prior = (_full_rows(NA, tgt_mask, tgt_prior_distrib[0], other_prior_distrib[0]),
                 _full_rows(NA, tgt_mask, tgt_prior_distrib[1], other_prior_distrib[1]))
        loop = SolLoop(model, scene_graph, map_idx, map_env, embed_info, cur_z2, final_result_traj[:, 0][~tgt_mask], w, lr, future_len, prior)
        for it in range(num_iters):
            loop.step()
            if debug is not None and it == 0:
                gt, go = loop.grads()
                debug.update(g_tgt=gt.clone(), g_other=go.clone())
            if log is not None:
                log(it, loop.log_dict())
        cur_z = loop.z.clone()
    else:
        tgt_z = tgt_prior_distrib[0].clone().detach().requires_grad_(True)             # (B, D)  sol_optim.py:38-40
        other_z = cur_z2[~tgt_mask].clone().detach().requires_grad_(True)
        opt = torch.optim.Adam([tgt_z, other_z], lr=lr)
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


class ShardedJob(object):
    """One global batch of independent reference batches ("loss-normalisation groups") optimised across the ranks of the
    default process group: `shard.partition_groups` gives every rank whole groups (cost-balanced), the rank runs a
    device-resident loop on them, and `run()` ends with the gather of the optimised rows on rank `dst`.  No collective inside
    the loop: every group gets exactly what the reference computes for that batch on its own, whatever the world size.

    scene: dict of CPU tensors describing the FULL batch, identical on every rank (ptr, past, lw, sem, map_idx, z, map_feat,
    past_feat, prior_mu, prior_var [, ext_future for kind='adv']).  kind: 'refine' (RefineLoop) or 'adv' (AdvLoop)."""

    def __init__(self, kind, model, scene, map_env, loss_weights, lr, FT, group_scene_ptr, dst=0, use_graph=True, **loop_kw):
        import torch.distributed as dist
        from . import shard
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.world = self.dist.get_world_size() if self.dist is not None else 1
        self.rank = self.dist.get_rank() if self.dist is not None else 0
        self.dst = dst
        self.NA = int(scene['ptr'][-1])
        self.FT = int(FT)
        self.zdim = int(scene['z'].size(1))
        costs = shard.group_costs(scene['ptr'], group_scene_ptr, FT)
        self.assign = shard.partition_groups(costs, self.world)
        self.loads = [sum(costs[g] for g in a) for a in self.assign]
        mine = self.assign[self.rank]
        sub, lgptr, self.agent_index = shard.shard_scenes(scene, group_scene_ptr, mine)
        # every rank can compute every other rank's rows: the gather needs no size exchange
        ptr = scene['ptr']
        self.rows_per_rank = [sum(int(ptr[int(group_scene_ptr[g + 1])]) - int(ptr[int(group_scene_ptr[g])]) for g in a) for a in self.assign]
        self.all_index = None
        if self.rank == dst:
            self.all_index = [shard.group_rows(ptr, group_scene_ptr, a) for a in self.assign]
        self.loop = None
        self.units_local = int(self.agent_index.numel()) * self.FT
        if self.agent_index.numel() == 0:
            return
        dev = map_env.device
        g = _DictGraph()
        for k in ('past', 'lw', 'sem', 'ptr', 'batch', 'edge_index'):
            setattr(g, k, sub[k].to(dev))
        embed = {'map_feat': sub['map_feat'].to(dev), 'past_feat': sub['past_feat'].to(dev),
                 'prior_out': (sub['prior_mu'].to(dev), sub['prior_var'].to(dev))}
        midx = sub['map_idx'].to(dev)
        if kind == 'refine':
            self.loop = RefineLoop(model, g, midx, map_env, embed, sub['z'].to(dev), loss_weights, lr, FT, group_scene_ptr=lgptr,
                                   use_graph=use_graph, **loop_kw)
        elif kind == 'adv':
            self.loop = AdvLoop(model, g, midx, map_env, embed, sub['z'].to(dev), sub['ext_future'].to(dev), loss_weights, lr, FT,
                                embed['prior_out'], group_scene_ptr=lgptr, use_graph=use_graph, **loop_kw)
        else:
            raise RuntimeError('strive_b200: unknown sharded loop kind %r' % kind)

    @property
    def imbalance(self):
        """max / mean of the per-rank cost estimate (1.0 = perfectly balanced)."""
        mean = sum(self.loads) / float(len(self.loads))
        return max(self.loads) / mean if mean > 0 else 1.0

    def run(self, iters):
        """`iters` iterations on this rank's groups, then the gather; rank `dst` returns the (NA, zdim) latents in batch order
        (a CPU tensor), the other ranks None."""
        from . import shard
        if self.loop is not None:
            z = self.loop.run(iters)
        else:
            z = torch.zeros((0, self.zdim), dtype=torch.float32)
        return shard.gather_rows(z, self.agent_index, self.NA, dst=self.dst, all_index=self.all_index, rows_per_rank=self.rows_per_rank)


def refine_sharded(model, scene, map_env, loss_weights, iters, lr, FT, group_scene_ptr, veh_coll_buffer=0.2, dst=0):
    """Refines the latents of a whole batch across the ranks of the default process group; see ShardedJob."""
    return ShardedJob('refine', model, scene, map_env, loss_weights, lr, FT, group_scene_ptr, dst=dst, veh_coll_buffer=veh_coll_buffer).run(iters)


def adv_sharded(model, scene, map_env, loss_weights, iters, lr, FT, group_scene_ptr, dst=0, **adv_kw):
    """The adversarial loop (planner replay) of a whole batch of reference batches across ranks; scene['ext_future'] (S,FT,4) is
    the planner's future per scene."""
    return ShardedJob('adv', model, scene, map_env, loss_weights, lr, FT, group_scene_ptr, dst=dst, **adv_kw).run(iters)
