"""Drop-in loss modules of STRIVE's latent optimisation (reference src/losses/adv_gen_nusc.py) on the fused CUDA
loss kernels (csrc/loss.cu through strive_loss_fwd_bwd).

Same constructors / forward signatures / returned dict keys as the reference:
    AvoidCollLoss(loss_weights, veh_att, mapixes, map_env, init_z, veh_coll_buffer=0.0, single_veh_idx=None, ptr=None)
        (future_pred_UNNORMALISED, z, prior_out) -> {loss, coll_veh_loss, coll_env_loss, motion_prior_loss, init_loss}   :264-341
    AdvGenLoss(loss_weights, veh_att, mapixes, map_env, init_z, ptr, veh_coll_buffer, crash_loss_min_time, crash_loss_min_infront)
        (future_pred, tgt_traj, z, prior_out, return_mins=False, attack_agt_idx=None) -> {loss, ..., adv_crash_loss}       :53-262
    TgtMatchingLoss(loss_weights)(future_pred, tgt_traj, z, prior_out) -> {loss, match_ext_loss, ...}                      :14-51
Per-term entries hold the term's batch MEAN as a 1-element tensor (the drivers only ever take torch.mean(v).item(),
refine_traffic_optim.py:203-207); `loss` is differentiable w.r.t. future_pred and z exactly as in the reference.
"""
import ctypes as C

import torch
from torch import nn

from . import _cabi
from .runtime import MapEnv

ENV_PAD = 128


def linspace5(cmin, cmax):
    """Vectorised torch.linspace(cmin[i], cmax[i], 5) in float32 (ATen: start+step*i for i<steps/2, else end-step*(steps-1-i)),
    = VehCollLoss centroids, adv_gen_nusc.py:432-435."""
    step = (cmax - cmin) / 4.0
    return torch.stack([cmin, cmin + step, cmax - step * 2.0, cmax - step, cmax], dim=1).contiguous()


class LossPlan(object):
    """Device-side constants of one loss configuration (what the reference loss constructors precompute)."""

    def __init__(self, kind, w, lw_un, agent_map, map_env, ptr_host, device, group_scene_ptr=None, coll_by_scene=True,
                 veh_coll_buffer=0.0, single_veh_idx=None, crash_min_t=0, crash_min_infront=None, traj_unnormalized=False,
                 z_row_mask=None, match_mask=None):
        if not isinstance(map_env, MapEnv):
            raise RuntimeError('strive_b200: map_env must be a strive_b200.MapEnv')
        self.kind, self.device, self.env = kind, device, map_env
        dev = device
        NA = int(lw_un.size(0))
        self.NA = NA
        ptr_host = torch.as_tensor(ptr_host, dtype=torch.int64).cpu()
        S = ptr_host.numel() - 1
        self.S = S
        sizes = ptr_host[1:] - ptr_host[:-1]
        if group_scene_ptr is None:
            group_scene_ptr = [0, S]
        gsp = torch.as_tensor(group_scene_ptr, dtype=torch.int64)
        G = gsp.numel() - 1
        self.G = G
        gap = ptr_host[gsp]                                              # agent ranges of the groups
        i32 = lambda t: t.to(device=dev, dtype=torch.int32).contiguous()
        self.group_agent_ptr = i32(gap)
        gsizes = gap[1:] - gap[:-1]
        self.group_of = i32(torch.repeat_interleave(torch.arange(G), gsizes))
        self.agent_map = i32(agent_map)
        if coll_by_scene:
            cptr = ptr_host
        else:
            cptr = gap
        self.cblock_ptr = i32(cptr)
        self.cblock_of = i32(torch.repeat_interleave(torch.arange(cptr.numel() - 1), cptr[1:] - cptr[:-1]))
        self.lw_un = lw_un.detach().to(dev, torch.float32).contiguous()
        rad = self.lw_un[:, 1] / 2.0
        self.circ_cx = linspace5(-(self.lw_un[:, 0] / 2.0) + rad, (self.lw_un[:, 0] / 2.0) - rad)
        ego = torch.zeros(NA, dtype=torch.bool)
        ego[ptr_host[:-1]] = True
        # rows entering the latent means and the env term
        if z_row_mask is None:
            if kind & _cabi.LOSS_ADV:
                z_row_mask = ~ego
            elif single_veh_idx is not None:
                z_row_mask = torch.zeros(NA, dtype=torch.bool)
                z_row_mask[ptr_host[:-1] + single_veh_idx] = True
            else:
                z_row_mask = torch.ones(NA, dtype=torch.bool)
        z_row_mask = z_row_mask.cpu()
        self.z_mask = z_row_mask.to(dev, torch.uint8).contiguous()
        grp_cpu = torch.repeat_interleave(torch.arange(G), gsizes)
        self.group_zrows = i32(torch.bincount(grp_cpu[z_row_mask], minlength=G))
        # get_coll_point grid per group (nuscenes_utils.py:351-354): batch-mean lw of the rows the env loss sees
        if kind & _cabi.LOSS_ADV:
            env_rows = ~ego
        elif single_veh_idx is not None:
            env_rows = z_row_mask
        else:
            env_rows = torch.ones(NA, dtype=torch.bool)
        mdx = torch.mean(map_env.nusc_dx.cpu()) * 0.5
        lw_cpu = self.lw_un.cpu()
        Ls, Ws = [], []
        lin_l = torch.zeros((G, ENV_PAD), dtype=torch.float32)
        lin_w = torch.zeros((G, ENV_PAD), dtype=torch.float32)
        for g in range(G):
            rows = env_rows & (grp_cpu == g)
            if int(rows.sum()) == 0:
                Ls.append(1); Ws.append(1)
                continue
            mlw = torch.mean(lw_cpu[rows], dim=0)
            Lg = int(torch.round(mlw[0] / mdx).int().item())
            Wg = int(torch.round(mlw[1] / mdx).int().item())
            if Lg > ENV_PAD or Wg > ENV_PAD or Lg < 1 or Wg < 1:
                raise RuntimeError('strive_b200: env collision grid %dx%d exceeds the supported %d' % (Lg, Wg, ENV_PAD))
            Ls.append(Lg); Ws.append(Wg)
            lin_l[g, :Lg] = torch.linspace(-1.0, 1.0, Lg)
            lin_w[g, :Wg] = torch.linspace(-1.0, 1.0, Wg)
        self.env_L, self.env_W = i32(torch.tensor(Ls)), i32(torch.tensor(Ws))
        self.env_lin_l, self.env_lin_w = lin_l.to(dev), lin_w.to(dev)
        self.match_mask = None
        self.group_match_rows = None
        self.attack_mask = None
        self.adv_min = torch.zeros((S, 2), dtype=torch.int32, device=dev) if (kind & _cabi.LOSS_ADV) else None
        g = lambda k: float(w.get(k, 0.0))
        c = _cabi.StriveLossCfg()
        c.kind, c.traj_unnormalized, c.num_groups = kind, int(bool(traj_unnormalized)), G
        c.group_agent_ptr, c.group_of, c.agent_map = _cabi.dptr(self.group_agent_ptr), _cabi.dptr(self.group_of), _cabi.dptr(self.agent_map)
        c.group_zrows = _cabi.dptr(self.group_zrows)
        c.cblock_ptr, c.cblock_of = _cabi.dptr(self.cblock_ptr), _cabi.dptr(self.cblock_of)
        c.w_coll_veh, c.w_coll_env, c.w_motion_prior, c.w_init_z = g('coll_veh'), g('coll_env'), g('motion_prior'), g('init_z')
        c.w_coll_veh_plan, c.w_init_z_atk, c.w_motion_prior_atk = g('coll_veh_plan'), g('init_z_atk'), g('motion_prior_atk')
        c.w_adv_crash, c.w_match_ext, c.w_motion_prior_ext = g('adv_crash'), g('match_ext'), g('motion_prior_ext')
        c.veh_coll_buffer = float(veh_coll_buffer)
        c.single_veh_idx = -1 if single_veh_idx is None else int(single_veh_idx)
        c.crash_min_t = int(crash_min_t)
        c.use_infront = int(crash_min_infront is not None)
        c.crash_min_infront = float(crash_min_infront) if crash_min_infront is not None else 0.0
        c.adv_min_out = _cabi.dptr(self.adv_min)
        c.env_L, c.env_W = _cabi.dptr(self.env_L), _cabi.dptr(self.env_W)
        c.env_lin_l, c.env_lin_w = _cabi.dptr(self.env_lin_l), _cabi.dptr(self.env_lin_w)
        c.circ_cx, c.lw_un = _cabi.dptr(self.circ_cx), _cabi.dptr(self.lw_un)
        self.cfg = c
        self.grp_cpu = grp_cpu
        if match_mask is not None:
            self.set_match_mask(match_mask)
        self._ws = None
        self._ws_ft = -1

    def set_match_mask(self, mask):
        """mask (NA,FT) bool: rows of the TgtMatchingLoss mean."""
        m = mask.to(self.device)
        self.match_mask = m.to(torch.uint8).contiguous()
        rows = torch.zeros(self.G, dtype=torch.int64)
        rows.index_add_(0, self.grp_cpu, m.sum(dim=1).cpu())
        self.group_match_rows = rows.clamp_min(1).to(self.device, torch.int32).contiguous()
        self.cfg.group_match_rows = _cabi.dptr(self.group_match_rows)

    def set_attack_mask(self, mask):
        self.attack_mask = None if mask is None else mask.to(self.device, torch.int32).contiguous()
        self.cfg.attack_mask = _cabi.dptr(self.attack_mask)

    def workspace(self, FT):
        if self._ws is None or self._ws_ft != FT:
            nb = _cabi.lib().strive_loss_workspace_bytes(self.NA, FT, self.G)
            self._ws = torch.empty(nb, dtype=torch.uint8, device=self.device)
            self._ws_ft = FT
        return self._ws


def run_loss(plan, scene_struct, traj, z_full, prior_mu, prior_var, init_z, match_tgt=None, adv_tgt=None,
             d_traj=None, d_traj_match=None, d_z=None, terms=None):
    """One strive_loss_fwd_bwd call.  All tensors CUDA float32 contiguous; returns (d_traj, d_traj_match, d_z, terms)."""
    L = _cabi.lib()
    NA, FT = traj.size(0), traj.size(1)
    dev = traj.device
    main = plan.kind & (_cabi.LOSS_AVOID | _cabi.LOSS_ADV)
    if main and d_traj is None:
        d_traj = torch.empty_like(traj)
    if main and d_z is None:
        d_z = torch.empty((NA, 32), dtype=torch.float32, device=dev)
    if (plan.kind & _cabi.LOSS_MATCH) and d_traj_match is None:
        d_traj_match = torch.empty_like(traj)
    if terms is None:
        terms = torch.empty((plan.G, _cabi.STRIVE_TERMS), dtype=torch.float32, device=dev)
    ws = plan.workspace(FT)
    with torch.cuda.device(traj.device):
        _cabi.check(L.strive_loss_fwd_bwd(C.byref(plan.cfg), C.byref(scene_struct), C.byref(plan.env.cstruct), FT,
                                          _cabi.dptr(traj), _cabi.dptr(z_full), _cabi.dptr(prior_mu), _cabi.dptr(prior_var),
                                          _cabi.dptr(init_z), _cabi.dptr(plan.z_mask), _cabi.dptr(match_tgt),
                                          _cabi.dptr(plan.match_mask), _cabi.dptr(adv_tgt), _cabi.dptr(d_traj),
                                          _cabi.dptr(d_traj_match), _cabi.dptr(d_z), _cabi.dptr(terms), _cabi.dptr(ws),
                                          ws.numel(), _cabi.stream_ptr()))
    return d_traj, d_traj_match, d_z, terms


def _mini_scene(ptr_host, agent_map, device, NC=2):
    """StriveScene carrying only what the loss kernels read (ptr, scene_of)."""
    ptr_host = torch.as_tensor(ptr_host, dtype=torch.int64).cpu()
    S = ptr_host.numel() - 1
    sizes = ptr_host[1:] - ptr_host[:-1]
    keep = dict(ptr=ptr_host.to(device, torch.int32).contiguous(),
                scene_of=torch.repeat_interleave(torch.arange(S), sizes).to(device, torch.int32).contiguous())
    st = _cabi.StriveScene(int(ptr_host[-1]), S, int(sizes.max()), NC, _cabi.dptr(keep['ptr']), _cabi.dptr(keep['scene_of']),
                           None, None, None, None)
    return st, keep


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, future, z_rows, mod, prior_mu, prior_var, init_z, match_tgt, adv_tgt, want_match):
        plan = mod.plan
        fut = future.detach().contiguous().float()
        d_traj, d_traj_m, d_z, terms = run_loss(plan, mod.scene_struct, fut, mod._full(z_rows), mod._full(prior_mu),
                                                mod._full(prior_var), mod._full(init_z), match_tgt=match_tgt, adv_tgt=adv_tgt)
        ctx.mod = mod
        ctx.z_is_none = z_rows is None
        ctx.save_for_backward(d_traj if d_traj is not None else d_traj_m, d_z if d_z is not None else torch.zeros(1, device=fut.device))
        ctx.match_only = d_traj is None
        col = 11 if ctx.match_only else 0
        loss = terms[:, col].sum()
        # the per-term table is a diagnostic (the drivers only print its means): without this it would require grad, and a
        # caller copying it into a persistent buffer (`log_buf.copy_(terms)`) would chain every iteration's graph -- and its
        # multi-GB rollout tape -- onto that buffer forever
        ctx.mark_non_differentiable(terms)
        return loss, terms

    @staticmethod
    def backward(ctx, g_loss, g_terms):
        d_traj, d_z = ctx.saved_tensors
        gf = d_traj * g_loss
        gz = None
        if not ctx.match_only and not ctx.z_is_none:
            gz = ctx.mod._rows(d_z) * g_loss
        return gf, gz, None, None, None, None, None, None, None


class _Base(nn.Module):
    def _full(self, rows):
        """(K,32) rows of the masked agents -> (NA,32) in graph order."""
        if rows is None:
            return None
        rows = rows.detach()
        if rows.dim() == 3:
            rows = rows[:, 0, :]
        rows = rows.float()
        if rows.size(0) == self.NA:
            return rows.contiguous()
        full = torch.zeros((self.NA, 32), dtype=torch.float32, device=rows.device)
        full[self.row_idx] = rows
        return full

    def _rows(self, full):
        if self.row_idx.numel() == self.NA:
            return full
        return full[self.row_idx]


class AvoidCollLoss(_Base):
    def __init__(self, loss_weights, veh_att, mapixes, map_env, init_z, veh_coll_buffer=0.0, single_veh_idx=None, ptr=None,
                 group_scene_ptr=None, ptr_for_groups=None):
        """Reference signature (adv_gen_nusc.py:268-271) plus one extension: `group_scene_ptr` (+ `ptr_for_groups`, the scene
        ptr) evaluates several independent reference batches ("loss-normalisation groups") in one call; every mean and,
        when ptr is None, every collision block is then per group.  Default None = exactly one reference batch."""
        super().__init__()
        dev = veh_att.device
        self.NA = int(veh_att.size(0))
        self.loss_weights = loss_weights
        self.init_z = init_z
        if single_veh_idx is not None and ptr is None:
            raise RuntimeError('single_veh_idx requires ptr (adv_gen_nusc.py:294-295)')
        if group_scene_ptr is not None:
            src = ptr if ptr is not None else ptr_for_groups
            if src is None:
                raise RuntimeError('group_scene_ptr needs the scene ptr (ptr or ptr_for_groups)')
            ptr_host = src.detach().cpu()
        else:
            ptr_host = torch.tensor([0, self.NA]) if ptr is None else ptr.detach().cpu()
        # ptr=None: the whole batch is ONE collision block (VehCollLoss :443-451), as refine_traffic_optim.py:176-181 builds it
        self.plan = LossPlan(_cabi.LOSS_AVOID, loss_weights, veh_att, mapixes, map_env, ptr_host, dev, coll_by_scene=ptr is not None,
                             veh_coll_buffer=veh_coll_buffer, single_veh_idx=single_veh_idx, traj_unnormalized=True,
                             group_scene_ptr=group_scene_ptr)
        self.scene_struct, self._keep = _mini_scene(ptr_host, mapixes, dev)
        self.row_idx = torch.nonzero(self.plan.z_mask.bool(), as_tuple=False).flatten()

    def forward(self, future_pred, z, prior_out):
        zz = z[:, 0, :] if z.dim() == 3 else z
        # 3-D latents (B,1,D) as sol_optim.py:38-44 passes them: the reference's init term sums over dim=1 -- the size-1 SAMPLE
        # axis -- and then averages over B*D elements (adv_gen_nusc.py:333-335), i.e. 1/D of the 2-D value.  Replicated.
        init_scale = 1.0 / float(z.size(-1)) if z.dim() == 3 else 1.0
        self.plan.cfg.w_init_z = float(self.loss_weights.get('init_z', 0.0)) * init_scale
        init_z = self.init_z[:, 0, :] if self.init_z is not None and self.init_z.dim() == 3 else self.init_z
        loss, terms = _LossFn.apply(future_pred, zz, self, prior_out[0], prior_out[1], init_z, None, None, False)
        out = {}
        w = self.loss_weights
        if w['coll_veh'] > 0.0:
            out['coll_veh_loss'] = terms[:, 1]
        if w['coll_env'] > 0.0:
            out['coll_env_loss'] = terms[:, 3]
        if w['motion_prior'] > 0.0:
            out['motion_prior_loss'] = terms[:, 5]
        if w.get('init_z', 0.0) > 0.0:
            out['init_loss'] = terms[:, 6] * init_scale
        out['loss'] = loss
        self.last_terms = terms          # (G, STRIVE_TERMS) raw per-group table of the last call (diagnostics; not a reference key)
        return out


class TgtMatchingLoss(_Base):
    def __init__(self, loss_weights):
        super().__init__()
        self.loss_weights = loss_weights

    def forward(self, future_pred, tgt_traj, z, prior_out):
        """Pure elementwise term: kept in PyTorch ops on the device (4 tiny kernels), including the reference's :46 quirk
        (the prior NLL is computed but the matching mean is added a second time)."""
        w = self.loss_weights
        out = {}
        loss = 0.0
        tl = torch.sum((future_pred - tgt_traj) ** 2, dim=-1)
        if w['match_ext'] > 0.0:
            loss = loss + w['match_ext'] * tl.mean()
            out['match_ext_loss'] = tl
        if w['motion_prior_ext'] > 0.0:
            loss = loss + w['motion_prior_ext'] * tl.mean()
            mu, var = prior_out
            zz = z
            if zz.dim() == 3:
                mu, var = mu.unsqueeze(1), var.unsqueeze(1)
            out['motion_prior_ext_loss'] = (torch.log(torch.sqrt(var)) + 0.9189385332046727 + (zz - mu) ** 2 / (2 * var)).sum(-1)
        out['loss'] = loss
        return out


class AdvGenLoss(_Base):
    def __init__(self, loss_weights, veh_att, mapixes, map_env, init_z, ptr, veh_coll_buffer=0.0, crash_loss_min_time=0,
                 crash_loss_min_infront=None):
        super().__init__()
        dev = veh_att.device
        self.NA = int(veh_att.size(0))
        self.loss_weights = loss_weights
        self.init_z = init_z
        self.ptr = ptr
        ptr_host = ptr.detach().cpu()
        if crash_loss_min_infront is not None:
            assert -1 <= crash_loss_min_infront <= 1
        self.plan = LossPlan(_cabi.LOSS_ADV, loss_weights, veh_att, mapixes, map_env, ptr_host, dev, coll_by_scene=True,
                             veh_coll_buffer=veh_coll_buffer, crash_min_t=crash_loss_min_time,
                             crash_min_infront=crash_loss_min_infront, traj_unnormalized=True)
        self.scene_struct, self._keep = _mini_scene(ptr_host, mapixes, dev)
        self.row_idx = torch.nonzero(self.plan.z_mask.bool(), as_tuple=False).flatten()

    def forward(self, future_pred, tgt_traj, z, prior_out, return_mins=False, attack_agt_idx=None):
        # closed-loop planner mode (adv_gen_optim.py:98-103,143): the attacked trajectory is the model's own prediction of the
        # target, tgt_traj = future_pred[ptr[:-1]] WITH its graph; the kernel then reads those rows of future_pred itself and
        # folds d(loss)/d(tgt_traj) into the same rows of d(loss)/d(future_pred)
        own = bool(tgt_traj.requires_grad)
        if own and not getattr(self, '_own_checked', False):
            if not torch.equal(tgt_traj.detach()[:, :, :4], future_pred.detach()[self.ptr[:-1].long()][:, :, :4]):
                raise RuntimeError('strive_b200: a differentiable tgt_traj must be future_pred[scene_graph.ptr[:-1]] (closed-loop mode)')
            self._own_checked = True
        self.plan.cfg.adv_own_pred = int(own)
        if attack_agt_idx is not None:
            m = torch.zeros(self.NA, dtype=torch.int32, device=future_pred.device)
            m[attack_agt_idx] = 1
            self.plan.set_attack_mask(m)
        else:
            self.plan.set_attack_mask(None)
        tgt = tgt_traj.detach()[:, :, :4].contiguous().float()
        loss, terms = _LossFn.apply(future_pred, z, self, prior_out[0], prior_out[1], self.init_z, None, tgt, False)
        w = self.loss_weights
        out = {'init_loss': terms[:, 6], 'motion_prior_loss': terms[:, 5], 'coll_veh_loss': terms[:, 1],
               'coll_veh_plan_loss': terms[:, 7], 'coll_env_loss': terms[:, 3], 'adv_crash_loss': terms[:, 9], 'loss': loss}
        self.last_terms = terms
        if return_mins:
            mins = self.plan.adv_min.cpu().numpy()
            out['min_agt'] = mins[:, 0].astype(int)
            out['min_t'] = mins[:, 1].astype(int)
        return out
