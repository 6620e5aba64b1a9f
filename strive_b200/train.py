"""train_traffic step of the graph-VAE (BASELINE configs[3]) with data-parallel gradient all-reduce.

Reference: src/train_traffic.py:64-171 (run_one_epoch: forward -> TrafficModelLoss -> backward -> Adam.step),
src/models/traffic_model.py:178-225 (TrafficModel.forward: encode_map / encode_past / encode_future / prior / posterior /
rsample / decoder, twice with future_sample) and src/losses/traffic_model.py:20-118, 166-295 (TrafficModelLoss and the
training variants of VehCollLoss / EnvCollLoss).

Scope of this row, stated plainly: training needs gradients with respect to the WEIGHTS of the map CNN, the interaction nets
and the GRU.  The hand-written sm_100a kernels of this package implement the latent-optimisation path (d/dz only; the CNN has
no backward there at all), so the training step differentiates a PyTorch restatement of the model -- the same parameter tree
(`strive_b200.TrafficModel`, reference checkpoint keys), crops from the package's CUDA crop kernel (bit exact), bf16 autocast
for the dense layers (BASELINE configs[3]), fp32 master weights and Adam.  What is native here is the data-parallel plumbing:
one process per GPU, all parameters and gradients live in ONE flat fp32 buffer each, and the gradient exchange is a single NCCL
all-reduce of that 4.4 MB bucket over NVLink launched on a side stream as soon as the backward pass has produced the last
gradient (there is nothing to overlap it with earlier: every module is used at every rollout step, so all gradients complete
together at the end of BPTT).  The reference has no distributed code at all.
"""
import math

import torch
import torch.nn.functional as F

A_STATS = (0.409074, 1.045530)          # datasets/utils.py:121-127
DDH_STATS = (0.000046, 0.075032)
DT, MAXHDOT, MAXS = 0.5, 2.0 * math.pi, 50.0


def clique_edges(ptr, device):
    """(src, dst) of the full directed clique per scene (nuscenes_dataset.py:678-687), vectorised."""
    sizes = (ptr[1:] - ptr[:-1]).tolist()
    src, dst = [], []
    for s, n in enumerate(sizes):
        if n < 2:
            continue
        a = int(ptr[s])
        ii = torch.arange(a, a + n, device=device)
        I, J = torch.meshgrid(ii, ii, indexing='ij')
        m = I != J
        dst.append(I[m])
        src.append(J[m])
    if not src:
        z = torch.zeros(0, dtype=torch.long, device=device)
        return z, z
    return torch.cat(src), torch.cat(dst)


def t2f(frame, poses):
    """utils/transforms.py:78-139 (non-inverse), (N,4),(N,4)->(N,4)"""
    c, s = frame[:, 2], frame[:, 3]
    dx, dy = poses[:, 0] - frame[:, 0], poses[:, 1] - frame[:, 1]
    return torch.stack([c * dx + s * dy, -s * dx + c * dy, poses[:, 2] * c + poses[:, 3] * s, poses[:, 3] * c - poses[:, 2] * s], 1)


def interaction_net(net, feat, pos, sem, src, dst):
    """models/interaction_net.py:52-77 with precomputed clique edges (single-sample branch)."""
    x = net.mlp_in(feat)
    N = x.size(0)
    if src.numel() > 0:
        rel = t2f(pos[dst], pos[src])
        rel = torch.where(torch.isnan(rel), torch.zeros_like(rel), rel)
        msg = net.msg[0].edge_mlp(torch.cat([x[dst], x[src], sem[dst], sem[src], rel.to(x.dtype)], -1))
        aggr = torch.zeros((N, msg.size(1)), dtype=msg.dtype, device=x.device)
        aggr = aggr.scatter_reduce(0, dst.view(-1, 1).expand(-1, msg.size(1)), msg, 'amax', include_self=False)
        x = x.to(msg.dtype)
    else:
        aggr = torch.zeros_like(x)
    return net.mlp_out(net.msg[0].update_mlp(torch.cat([x, aggr, sem.to(x.dtype)], -1)))


def bicycle_step(state_un, a, ddh, veh_len):
    """traffic_model.py:714-733 + models/common.py:47-67, one step, fp32"""
    x, y, hx, hy, s, hdot = state_un.unbind(1)
    h = torch.atan2(hy, hx)
    newhdot = (hdot + ddh * DT).clamp(-MAXHDOT, MAXHDOT)
    newh = h + DT * s.abs() / veh_len * newhdot
    news = (s + a * DT).clamp(0.0, MAXS)
    return torch.stack([x + news * newh.cos() * DT, y + news * newh.sin() * DT, newh.cos(), newh.sin(), news, newhdot], 1)


class FlatBucket(object):
    """All parameters in ONE flat fp32 buffer and all gradients in another (parameters / .grad become views): the gradient exchange
    of data-parallel training is then a single all-reduce of one contiguous bucket (1 093 202 floats = 4.4 MB for the reference
    model) -- NCCL over NVLink when the default process group is NCCL, issued on a side stream behind an event recorded after
    the backward pass.  Works on CPU tensors with gloo as well (tests)."""

    def __init__(self, params):
        self.params = list(params)
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat_p = torch.empty(self.numel, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        o = 0
        for p in self.params:
            n = p.numel()
            self.flat_p[o:o + n].copy_(p.detach().reshape(-1))
            p.data = self.flat_p[o:o + n].view_as(p)
            p.grad = self.flat_g[o:o + n].view_as(p)
            o += n
        import torch.distributed as dist
        self.dist = dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.cuda = dev.type == 'cuda'
        if self.cuda:
            self.comm = torch.cuda.Stream(device=dev)
            self.ev_bwd = torch.cuda.Event()
            self.ev_c0, self.ev_c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.timed = False

    def zero_grad(self):
        self.flat_g.zero_()

    def check_views(self):
        """autograd must have accumulated IN PLACE into the bucket (it does when .grad is already defined)."""
        lo, hi = self.flat_g.data_ptr(), self.flat_g.data_ptr() + 4 * self.numel
        for p in self.params:
            if p.grad is None or not (lo <= p.grad.data_ptr() < hi):
                raise RuntimeError('strive_b200: a parameter gradient left the flat all-reduce bucket')

    def all_reduce_mean(self):
        self.check_views()
        if self.dist is None:
            return
        world = self.dist.get_world_size()
        if self.cuda:
            self.ev_bwd.record()
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(self.ev_bwd)
                self.ev_c0.record(self.comm)
                self.dist.all_reduce(self.flat_g, op=self.dist.ReduceOp.SUM)
                self.flat_g.div_(world)
                self.ev_c1.record(self.comm)
            torch.cuda.current_stream().wait_event(self.ev_c1)
            self.timed = True
        else:
            self.dist.all_reduce(self.flat_g, op=self.dist.ReduceOp.SUM)
            self.flat_g.div_(world)

    def last_all_reduce_ms(self):
        """device time of the last bucket all-reduce (CUDA only; synchronises)."""
        if not (self.cuda and self.timed):
            return 0.0
        self.ev_c1.synchronize()
        return self.ev_c0.elapsed_time(self.ev_c1)


class TrafficModelTrainer(object):
    """forward / loss / step of train_traffic on one rank.  `model` is a strive_b200.TrafficModel on a CUDA device."""

    def __init__(self, model, map_env, loss_weights, lr=1e-5, autocast_bf16=True, betas=(0.9, 0.999), eps=1e-8):
        self.model, self.env = model, map_env
        self.w = dict(loss_weights)
        self.autocast = bool(autocast_bf16)
        self.nrm, self.att = model.get_normalizer(), model.get_att_normalizer()
        self.dev = next(model.parameters()).device
        self.bucket = FlatBucket([p for p in model.parameters()])
        self.params = self.bucket.params
        self.flat_g = self.bucket.flat_g
        self.opt = torch.optim.Adam(self.params, lr=lr, betas=betas, eps=eps)

    # ---- model pieces -----------------------------------------------------------------------------------
    def encode_map(self, pos_n, mapixes):
        """traffic_model.py:416-451: crop (package CUDA kernel, integer gather: no gradient) -> map_conv -> map_feature."""
        pose_un = self.nrm.unnormalize(pos_n.detach()).contiguous().float()
        crop = self.env.crop_poses(pose_un, mapixes).float()
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=self.autocast):
            f = self.model.map_conv(crop)
            return self.model.map_feature(f.reshape(f.size(0), -1)).float()

    def decoder(self, g, map_feat, past_feat, z, mapixes, src, dst, FT):
        """traffic_model.py:589-704 (output_bicycle, single sample), differentiable in every weight."""
        m = self.model
        prev = g.past[:, -1, :]
        pos = prev[:, :4]
        veh_len = self.att.unnormalize(g.lw)[:, 0]
        mem = past_feat.unsqueeze(0).expand(3, -1, -1).contiguous()
        pf, mf = past_feat, map_feat
        traj = []
        for t in range(FT):
            feat = torch.cat([pf, mf, g.sem, z, g.lw], -1)
            with torch.autocast('cuda', dtype=torch.bfloat16, enabled=self.autocast):
                out = interaction_net(m.decoder_net, feat, pos, g.sem, src, dst).float()
            a = out[:, 0] * A_STATS[1] + A_STATS[0]
            ddh = out[:, 1] * DDH_STATS[1] + DDH_STATS[0]
            cur = self.nrm.normalize(bicycle_step(self.nrm.unnormalize(prev), a, ddh, veh_len))
            loc = t2f(prev[:, :4], cur[:, :4])
            traj.append(cur[:, :4])
            prev = cur
            if t < FT - 1:
                o, mem = m.decoder_memory(loc.unsqueeze(1), mem)
                pf = o[:, 0]
                mf = self.encode_map(cur[:, :4], mapixes)
                pos = cur[:, :4]
        return torch.stack(traj, 1)

    def forward(self, g, map_idx, future_sample=False, eps_post=None, eps_prior=None):
        """TrafficModel.forward, traffic_model.py:178-225.  eps_* (NA,32) fix the rsample noise (tests)."""
        m = self.model
        FT = m.FT
        mapixes = map_idx[g.batch].to(torch.int32)
        src, dst = clique_edges(g.ptr, self.dev)
        map_feat = self.encode_map(g.past[:, -1, :4], mapixes)
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=self.autocast):
            past_feat = m.encode_past(g).float()
            future_feat = m.encode_future(g).float()
            pos0 = g.past[:, -1, :4]
            pr = interaction_net(m.prior_net, torch.cat([past_feat, map_feat, g.sem], -1), pos0, g.sem, src, dst).float()
            po = interaction_net(m.posterior_net, torch.cat([past_feat, future_feat, map_feat, g.sem], -1), pos0, g.sem, src, dst).float()
        pm, pv = pr[:, :32], torch.exp(pr[:, 32:])
        qm, qv = po[:, :32], torch.exp(po[:, 32:])
        e = torch.randn_like(qm) if eps_post is None else eps_post
        z = qm + e * torch.sqrt(qv)
        out = {'prior_out': (pm, pv), 'posterior_out': (qm, qv), 'future_pred': self.decoder(g, map_feat, past_feat, z, mapixes, src, dst, FT)}
        if future_sample:
            e = torch.randn_like(pm) if eps_prior is None else eps_prior
            out['future_samp'] = self.decoder(g, map_feat, past_feat, pm + e * torch.sqrt(pv), mapixes, src, dst, FT)
        return out

    # ---- TrafficModelLoss, losses/traffic_model.py:34-118 ---------------------------------------------------
    def veh_coll_prior(self, g, traj_un):
        """training VehCollLoss (:166-240): penalties of ALL valid pairs (0 where not colliding), summed / number of pairs."""
        lw = self.att.unnormalize(g.lw)
        rad = lw[:, 1] / 2.0
        cmin, cmax = -(lw[:, 0] / 2.0) + rad, (lw[:, 0] / 2.0) - rad
        step = (cmax - cmin) / 4.0
        cx = torch.stack([cmin, cmin + step, cmax - step * 2.0, cmax - step, cmax], 1)        # torch.linspace(.,.,5)
        ptr = g.ptr.tolist()
        tot = traj_un.new_zeros(())
        pairs = 0
        for s in range(len(ptr) - 1):
            a, b = ptr[s], ptr[s + 1]
            n = b - a
            pairs += n * n - n
            if n < 2:
                continue
            tr = traj_un[a:b]
            wx = tr[:, :, 2:3] * cx[a:b].unsqueeze(1) + tr[:, :, 0:1]
            wy = tr[:, :, 3:4] * cx[a:b].unsqueeze(1) + tr[:, :, 1:2]
            cent = torch.stack([wx, wy], -1).transpose(0, 1).reshape(tr.size(1), n * 5, 2)
            d = torch.cdist(cent, cent, compute_mode='donot_use_mm_for_euclid_dist').view(-1, n, 5, n, 5).permute(0, 1, 3, 2, 4).reshape(-1, n, n, 25)
            mind = d.min(-1)[0]
            pd = rad[a:b].view(n, 1) + rad[a:b].view(1, n)
            mask = (mind <= pd) & (~torch.eye(n, dtype=torch.bool, device=self.dev)).view(1, n, n)
            tot = tot + torch.where(mask, 1.0 - mind / pd, torch.zeros_like(mind)).sum()
        return tot / max(pairs, 1)

    def env_coll_prior(self, g, traj_un_ego, map_idx):
        """training EnvCollLoss on the ego rows (:97-104, 242-295): (B,T) penalties, 0 where no collision point exists."""
        B, T, _ = traj_un_ego.shape
        ego = g.ptr[:-1].long()
        lw = self.att.unnormalize(g.lw[ego])
        flat = traj_un_ego.reshape(B * T, 4)
        att = lw.view(B, 1, 2).expand(B, T, 2).reshape(B * T, 2)
        mix = map_idx.view(B, 1).expand(B, T).reshape(B * T).long()
        dxm = self.env.nusc_dx
        mdx = torch.mean(dxm) * 0.5
        mlw = torch.mean(att, 0)
        L, W = int(torch.round(mlw[0] / mdx).int()), int(torch.round(mlw[1] / mdx).int())            # nuscenes_utils.py:351-354
        car = flat.detach()
        lwise = torch.linspace(-1.0, 1.0, L, device=self.dev).view(1, L, 1) * att[:, 0].view(-1, 1, 1) / 2
        wwise = torch.linspace(-1.0, 1.0, W, device=self.dev).view(1, 1, W) * att[:, 1].view(-1, 1, 1) / 2
        hc, hs = car[:, 2].view(-1, 1, 1), car[:, 3].view(-1, 1, 1)
        xyw = torch.stack([(lwise * hc - wwise * hs) + car[:, 0].view(-1, 1, 1), (lwise * hs + wwise * hc) + car[:, 1].view(-1, 1, 1)], -1)
        pix = torch.round(xyw / dxm[mix].view(-1, 1, 1, 2)).long()
        drv = self.env.nusc_raster[:, 0]
        outside = (pix[..., 1] < 0) | (pix[..., 1] >= drv.shape[1]) | (pix[..., 0] < 0) | (pix[..., 0] >= drv.shape[2])
        pix = torch.where(outside.unsqueeze(-1), torch.zeros_like(pix), pix)
        nd = drv[mix.view(-1, 1, 1).expand(-1, L, W), pix[..., 1], pix[..., 0]] == 0
        num = nd.sum((1, 2))
        pt = (xyw * nd.unsqueeze(-1)).sum((1, 2)) / num.view(-1, 1)
        valid = (num > 0) & (num < L * W)
        pen_d = torch.sqrt(att[:, 0] ** 2 / 4.0 + att[:, 1] ** 2 / 4.0)
        dist = torch.norm(flat[:, :2] - torch.where(valid.view(-1, 1), pt, flat[:, :2].detach() + 1.0), dim=1)
        return torch.where(valid, 1.0 - dist / pen_d, torch.zeros_like(dist)).view(B, T)

    def loss(self, g, pred, map_idx):
        vis = g.future_vis == 1.0
        gt = g.future_gt[vis][:, :4]
        pf = pred['future_pred'][vis]
        recon = (0.5 * math.log(2 * math.pi) + (pf - gt) ** 2 / 2.0).sum(-1)                           # -log_normal(pred, gt, 1), losses/common.py:26-41
        pm, pv = pred['prior_out']
        qm, qv = pred['posterior_out']
        kl = (0.5 * (torch.log(pv) - torch.log(qv) + qv / pv + (qm - pm).pow(2) / pv - 1)).sum(-1)        # kl_normal, :8-24
        loss = self.w['recon'] * recon.mean() + self.w['kl'] * kl.mean()
        out = {'recon_loss': recon, 'kl_loss': kl}
        if self.w.get('coll_veh_prior', 0.0) > 0.0 and 'future_samp' in pred:
            cv = self.veh_coll_prior(g, self.nrm.unnormalize(pred['future_samp']))
            loss = loss + self.w['coll_veh_prior'] * cv
            out['coll_veh_prior'] = cv.view(1)
        if self.w.get('coll_env_prior', 0.0) > 0.0 and 'future_samp' in pred:
            ce = self.env_coll_prior(g, self.nrm.unnormalize(pred['future_samp'][g.ptr[:-1].long()]), map_idx)
            loss = loss + self.w['coll_env_prior'] * ce.mean()
            out['coll_env_prior'] = ce.view(-1)
        out['loss'] = loss.view(1)
        return out

    # ---- one optimisation step, train_traffic.py:101-114 ------------------------------------------------------
    def backward(self, g, map_idx, eps_post=None, eps_prior=None):
        """zero_grad -> forward -> loss -> backward -> all-reduce of the flat gradient bucket (mean over ranks)."""
        self.bucket.zero_grad()
        do_sample = self.w.get('coll_veh_prior', 0.0) > 0.0 or self.w.get('coll_env_prior', 0.0) > 0.0
        pred = self.forward(g, map_idx, future_sample=do_sample, eps_post=eps_post, eps_prior=eps_prior)
        ld = self.loss(g, pred, map_idx)
        ld['loss'][0].backward()
        self.bucket.all_reduce_mean()
        return ld

    def step(self, g, map_idx):
        ld = self.backward(g, map_idx)
        self.opt.step()
        return ld
