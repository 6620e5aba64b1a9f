"""Drop-in for the decode side of STRIVE's `models.traffic_model.TrafficModel`.

Keeps the reference call surface the latent-optimisation drivers use:
    decode_embedding(z, embed_out, scene_graph, map_idx, map_env, ext_future=None, nfuture=None) -> {'future_pred'}
    encode_map / embed / sample_batched / prior / encode_past, set_/get_normalizer, set_/get_att_normalizer,
    set_bicycle_params, attributes FT, dt, normalizer, state_dict keys identical to the reference checkpoint.
(reference src/models/traffic_model.py:23-176, 319-451, 545-587).

The per-iteration path -- decoder rollout (GNN + GRU + bicycle + per-step map re-encode) and its adjoint with
respect to z -- runs in hand-written sm_100a CUDA behind the C-ABI (include/strive_b200.h).  The once-per-batch
producers (`embed`: past encoder + prior net) stay plain PyTorch, as north_star prescribes; they call the CUDA map
encoder for `map_feat`.  There is no CPU fallback: tensors must be CUDA tensors.
"""
import ctypes as C

import torch
from torch import nn

from . import _cabi
from .runtime import DeviceModel, SceneBatch, MapEnv

STATE_MEAN = (0.0, 0.0, 0.0, 0.0, 1.802009, -0.000037)     # reference src/datasets/utils.py:131-140
STATE_STD = (15.0, 15.0, 1.0, 1.0, 3.507907, 0.055684)
ATT_MEAN = (4.844294, 2.021752)
ATT_STD = (1.084860, 0.299647)
NUSC_BIKE_PARAMS = {'maxs': 50.0, 'maxhdot': 6.283185307179586, 'dt': 0.5,
                    'a_stats': (0.409074, 1.045530), 'ddh_stats': (0.000046, 0.075032)}


class MeanStdNormalizer(object):
    """Same semantics as reference src/datasets/utils.py:44-113 (normalise the first D' <= D components)."""

    def __init__(self, mean_vals, std_vals):
        self.mean_vals = torch.as_tensor(mean_vals).to(torch.float)
        self.std_vals = torch.as_tensor(std_vals).to(torch.float)
        self.D = self.mean_vals.size(0)

    def _ms(self, x):
        d = x.size(-1)
        shape = [1] * (x.dim() - 1) + [d]
        return self.mean_vals[:d].reshape(shape).to(x.device), self.std_vals[:d].reshape(shape).to(x.device)

    def normalize(self, x):
        m, s = self._ms(x)
        return (x - m) / s

    def unnormalize(self, x):
        m, s = self._ms(x)
        return (x * s) + m

    def normalize_single(self, x, idx):
        return (x - self.mean_vals[idx].to(x.device)) / self.std_vals[idx].to(x.device)

    def unnormalize_single(self, x, idx):
        return (x * self.std_vals[idx].to(x.device)) + self.mean_vals[idx].to(x.device)


class MLP(nn.Module):
    """Linear, then [LayerNorm, ReLU, Linear] per extra layer; parameter names `net.<i>` as the reference MLP
    (src/models/common.py:8-44) so checkpoints load."""

    def __init__(self, layers):
        super().__init__()
        mods = [nn.Linear(layers[0], layers[1])]
        for i in range(1, len(layers) - 1):
            mods += [nn.LayerNorm(layers[i]), nn.ReLU(), nn.Linear(layers[i], layers[i + 1])]
        self.net = nn.ModuleList(mods)

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class _Conv(nn.Module):
    def __init__(self, node, sem, edge, out, hidden=128):
        super().__init__()
        self.edge_mlp = MLP([2 * (node + sem) + edge, hidden, hidden, out])
        self.update_mlp = MLP([node + out + sem, hidden, out])


class SceneInteractionNet(nn.Module):
    """Host-side (PyTorch) interaction net for the once-per-batch prior; same parameter tree as the reference
    (src/models/interaction_net.py:16-77) -- the decoder instance is executed by the CUDA kernels instead."""

    def __init__(self, in_node, in_sem, in_edge, msg_node, out_channels):
        super().__init__()
        self.mlp_in = MLP([in_node, 128, 128, msg_node])
        self.msg = nn.ModuleList([_Conv(msg_node, in_sem, in_edge, msg_node)])
        self.mlp_out = MLP([msg_node, 128, 128, out_channels])

    def forward(self, feat, pos, sem, ptr):
        x = self.mlp_in(feat)
        N = x.size(0)
        sizes = (ptr[1:] - ptr[:-1]).tolist()
        src, dst = [], []
        for s, n in enumerate(sizes):
            if n < 2:
                continue
            a = int(ptr[s])
            ii = torch.arange(a, a + n, device=x.device)
            I, J = torch.meshgrid(ii, ii, indexing='ij')
            m = I != J
            dst.append(I[m])
            src.append(J[m])
        if src:
            src, dst = torch.cat(src), torch.cat(dst)
            f, p = pos[dst], pos[src]
            c, s_ = f[:, 2], f[:, 3]
            dx, dy = p[:, 0] - f[:, 0], p[:, 1] - f[:, 1]
            rel = torch.stack([c * dx + s_ * dy, -s_ * dx + c * dy, p[:, 2] * c + p[:, 3] * s_, p[:, 3] * c - p[:, 2] * s_], 1)
            rel = torch.where(torch.isnan(rel), torch.zeros_like(rel), rel)
            msg = self.msg[0].edge_mlp(torch.cat([x[dst], x[src], sem[dst], sem[src], rel], -1))
            aggr = torch.zeros((N, msg.size(1)), dtype=x.dtype, device=x.device)
            aggr = aggr.scatter_reduce(0, dst.view(-1, 1).expand(-1, msg.size(1)), msg, 'amax', include_self=False)
        else:
            aggr = torch.zeros((N, 64), dtype=x.dtype, device=x.device)
        x = self.msg[0].update_mlp(torch.cat([x, aggr, sem], -1))
        return self.mlp_out(x)


class _TapeLease(object):
    """A rollout tape (several GB at BASELINE configs[1]) borrowed from a per-device free list and returned to it when the
    autograd node that owns it dies.  Cycling buffers of this size through torch's caching allocator lets it split the freed
    block for small requests, after which the next tape no longer fits and costs a cudaMalloc (measured: 50-180 ms stalls
    in 1 of 3 iterations of the drop-in API path)."""
    _free = {}
    made = 0          # buffers ever allocated (diagnostic)

    def __init__(self, nbytes, device):
        key = (str(device), int(nbytes))
        lst = _TapeLease._free.setdefault(key, [])
        self.key = key
        if lst:
            self.buf = lst.pop()
        else:
            _TapeLease.made += 1
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)

    def __del__(self):
        try:
            lst = _TapeLease._free.get(self.key)
            if lst is not None and len(lst) < 2 and self.buf is not None:
                lst.append(self.buf)           # keep at most two per size; the rest goes back to the allocator
            self.buf = None
        except Exception:
            pass

    @staticmethod
    def release_all():
        _TapeLease._free.clear()


class _DecodeFn(torch.autograd.Function):
    """autograd boundary of the CUDA rollout: forward = strive_decode_fwd, backward = strive_decode_bwd (d/dz only;
    embed tensors are detached by every caller, refine_traffic_optim.py:160, adv_scenario_gen.py:271)."""

    @staticmethod
    def forward(ctx, z, model, scene, env, map_feat, past_feat, ext_future, FT):
        L = _cabi.lib()
        NA = scene.NA
        z = z.detach().contiguous().float()
        traj = torch.empty((NA, FT, 4), dtype=torch.float32, device=z.device)
        nbytes = L.strive_decode_tape_bytes(NA, FT)
        lease = _TapeLease(nbytes, z.device)
        tape = lease.buf
        ext = None if ext_future is None else ext_future.detach().contiguous().float()
        with torch.cuda.device(z.device):
            _cabi.check(L.strive_decode_fwd(model.handle, C.byref(scene.cstruct), C.byref(env.cstruct), _cabi.dptr(z),
                                            _cabi.dptr(map_feat), _cabi.dptr(past_feat), _cabi.dptr(ext), FT,
                                            _cabi.dptr(traj), _cabi.dptr(tape), nbytes, _cabi.stream_ptr()))
        ctx.model, ctx.scene, ctx.tape, ctx.nbytes, ctx.ext, ctx.FT = model, scene, tape, nbytes, ext, FT
        ctx.lease = lease        # returned to the free list when this node is freed
        return traj

    @staticmethod
    def backward(ctx, d_traj):
        L = _cabi.lib()
        d_traj = d_traj.contiguous().float()
        d_z = torch.empty((ctx.scene.NA, 32), dtype=torch.float32, device=d_traj.device)
        with torch.cuda.device(d_traj.device):
            _cabi.check(L.strive_decode_bwd(ctx.model.handle, C.byref(ctx.scene.cstruct), ctx.FT, _cabi.dptr(ctx.ext),
                                            _cabi.dptr(d_traj), _cabi.dptr(d_z), _cabi.dptr(ctx.tape), ctx.nbytes,
                                            _cabi.stream_ptr()))
        return d_z, None, None, None, None, None, None, None


class TrafficModel(nn.Module):
    def __init__(self, npast, nfuture, map_obs_size_pix, nclasses, map_feat_size=64, past_feat_size=64,
                 future_feat_size=64, latent_size=32, output_bicycle=True, traj_encoder='mlp', conv_channel_in=4,
                 conv_kernel_list=(7, 5, 5, 3, 3, 3), conv_stride_list=(2, 2, 2, 2, 2, 2),
                 conv_filter_list=(16, 32, 64, 64, 128, 128)):
        super().__init__()
        if (map_obs_size_pix != 256 or conv_channel_in != 4 or tuple(conv_kernel_list) != (7, 5, 5, 3, 3, 3)
                or tuple(conv_stride_list) != (2, 2, 2, 2, 2, 2) or tuple(conv_filter_list) != (16, 32, 64, 64, 128, 128)
                or map_feat_size != 64 or past_feat_size != 64 or latent_size != 32 or not output_bicycle
                or traj_encoder != 'mlp'):
            raise RuntimeError('strive_b200: kernels are built for the reference default architecture '
                               '(configs/*.cfg: 256 px crop, 4 layers, conv 7-5-5-3-3-3, 64-d features, z=32, bicycle, mlp)')
        self.normalizer = self.att_normalizer = None
        self.PT, self.FT, self.dt, self.NC = npast, nfuture, 0.5, nclasses
        self.output_bicycle = True
        self.bicycle_params = None
        self.state_size, self.att_feat_size, self.z_size = 6, 2, latent_size
        self.map_obs_size_pix = map_obs_size_pix
        chans = [conv_channel_in] + list(conv_filter_list)
        layers = []
        for i in range(6):
            layers += [nn.Conv2d(chans[i], chans[i + 1], conv_kernel_list[i], stride=2), nn.GroupNorm(1, chans[i + 1]), nn.ReLU()]
        self.map_conv = nn.Sequential(*layers)           # parameters only; executed by csrc/mapenc.cu
        self.map_feat_in_size = 128 * 2 * 2
        self.map_feature = nn.Linear(self.map_feat_in_size, map_feat_size)
        self.past_feat_size = past_feat_size
        self.past_in_size = nclasses + npast * (6 + 2 + 1)
        self.past_encoder = MLP([self.past_in_size, 128, 128, 128, past_feat_size])
        self.future_in_size = nclasses + nfuture * (6 + 2 + 1)
        self.future_encoder = MLP([self.future_in_size, 128, 128, 128, future_feat_size])
        self.prior_net = SceneInteractionNet(past_feat_size + map_feat_size + nclasses, nclasses, 4, 2 * past_feat_size, 2 * latent_size)
        self.posterior_net = SceneInteractionNet(future_feat_size + past_feat_size + map_feat_size + nclasses, nclasses, 4,
                                                 2 * past_feat_size, 2 * latent_size)
        self.decoder_net = SceneInteractionNet(latent_size + past_feat_size + map_feat_size + nclasses + 2, nclasses, 4, 64, 2)
        self.num_memory_layers = 3
        self.decoder_memory = nn.GRU(4, past_feat_size, 3, batch_first=True)
        self._dev_model = None
        self._dev_key = None

    # ---- reference setters / getters (traffic_model.py:160-176) ----
    def set_normalizer(self, normalizer):
        self._check_stats(normalizer, STATE_MEAN, STATE_STD, 'state')
        self.normalizer = normalizer

    def get_normalizer(self):
        return self.normalizer

    def set_att_normalizer(self, normalizer):
        self._check_stats(normalizer, ATT_MEAN, ATT_STD, 'attribute')
        self.att_normalizer = normalizer

    def get_att_normalizer(self):
        return self.att_normalizer

    def set_bicycle_params(self, bicycle_params):
        for k, v in NUSC_BIKE_PARAMS.items():
            got = bicycle_params[k]
            same = all(abs(float(a) - float(b)) < 1e-9 for a, b in zip(got, v)) if isinstance(v, tuple) else abs(float(got) - float(v)) < 1e-6
            if not same:
                raise RuntimeError('strive_b200: kernels are compiled for NUSC_BIKE_PARAMS (datasets/utils.py:121-127); %s differs' % k)
        self.bicycle_params = bicycle_params

    @staticmethod
    def _check_stats(nrm, mean, std, what):
        m = [float(v) for v in nrm.mean_vals]
        s = [float(v) for v in nrm.std_vals]
        if any(abs(a - b) > 1e-6 for a, b in zip(m, mean)) or any(abs(a - b) > 1e-6 for a, b in zip(s, std)):
            raise RuntimeError('strive_b200: kernels are compiled for the car/truck %s statistics of '
                               'datasets/utils.py:131-140; got mean=%s std=%s' % (what, m, s))

    # ---- device plumbing ----
    def device_model(self):
        p = next(self.parameters())
        if not p.is_cuda:
            raise RuntimeError('strive_b200: model must live on a CUDA device (no CPU fallback)')
        key = (p.device, sum(int(q._version) for q in self.parameters()))
        if self._dev_model is None or self._dev_key != key:
            self._dev_model = DeviceModel(self.state_dict(), self.NC, p.device)
            self._dev_key = key
        return self._dev_model

    @staticmethod
    def _sources(scene_graph, map_idx):
        return (scene_graph.past, scene_graph.ptr, scene_graph.lw, scene_graph.sem, map_idx)

    @staticmethod
    def _cache_get(scene_graph, slot, srcs, extra=None):
        """Cached derived object stored ON the scene graph (so it dies with it; no id()/data_ptr() reuse across batches).  An
        entry is valid only if every source tensor is the very same object, unmodified in place since (tensor._version), and
        of the same shape; the entry holds the sources strongly, so their identity cannot be recycled while it lives."""
        ent = getattr(scene_graph, slot, None)
        if ent is None or ent[0] != extra or len(ent[1]) != len(srcs):
            return None
        for (t, ver, shp), cur in zip(ent[1], srcs):
            if t is not cur or cur._version != ver or tuple(cur.shape) != shp:
                return None
        return ent[2]

    @staticmethod
    def _cache_put(scene_graph, slot, srcs, obj, extra=None):
        try:
            setattr(scene_graph, slot, (extra, [(t, t._version, tuple(t.shape)) for t in srcs], obj))
        except Exception:          # graph type without settable attributes: rebuild per call
            pass

    def scene_batch(self, scene_graph, map_idx):
        srcs = self._sources(scene_graph, map_idx)
        dev = next(self.parameters()).device
        sb = self._cache_get(scene_graph, '_strive_scene_batch', srcs, extra=str(dev))
        if sb is None:
            sb = SceneBatch(scene_graph, map_idx, device=dev)
            self._cache_put(scene_graph, '_strive_scene_batch', srcs, sb, extra=str(dev))
        return sb

    @staticmethod
    def _env(map_env):
        if not isinstance(map_env, MapEnv):
            raise RuntimeError('strive_b200: map_env must be a strive_b200.MapEnv (wrap nusc_raster / nusc_dx with it)')
        return map_env

    # ---- hot path ----
    def decode_embedding(self, z, embed_out, scene_graph, map_idx, map_env, ext_future=None, nfuture=None):
        """reference traffic_model.py:405-414.  z (NA,32) or (NA,1,32)."""
        FT = self.FT if nfuture is None else int(nfuture)
        scene = self.scene_batch(scene_graph, map_idx)
        three_d = z.dim() == 3
        if three_d:
            if z.size(1) != 1:
                raise RuntimeError('strive_b200: differentiable decode supports NS=1 (sol_optim.py:38-44); use sample_batched for NS>1')
            z2 = z[:, 0, :]
        else:
            z2 = z
        mf = embed_out['map_feat'].detach().contiguous().float()
        pf = embed_out['past_feat'].detach().contiguous().float()
        # the kernels index raw pointers: shapes the reference would reject with an IndexError are rejected here
        if FT < 1:
            raise RuntimeError('strive_b200: nfuture must be >= 1')
        if z2.dim() != 2 or z2.size(0) != scene.NA or z2.size(1) != self.z_size:
            raise RuntimeError('strive_b200: z must be (%d,%d) or (%d,1,%d), got %s' % (scene.NA, self.z_size, scene.NA, self.z_size, tuple(z.shape)))
        for name, t in (('map_feat', mf), ('past_feat', pf)):
            if tuple(t.shape) != (scene.NA, 64):
                raise RuntimeError('strive_b200: embed_out[%r] must be (%d,64), got %s' % (name, scene.NA, tuple(t.shape)))
        if ext_future is not None and (ext_future.dim() != 3 or ext_future.size(0) != scene.S or ext_future.size(1) < FT
                                       or ext_future.size(2) < 4):
            raise RuntimeError('strive_b200: ext_future must be (%d, >=%d, 4), got %s' % (scene.S, FT, tuple(ext_future.shape)))
        for name, t in (('z', z2), ('map_feat', mf), ('past_feat', pf), ('ext_future', ext_future)):
            if t is not None and t.device != scene.device:
                raise RuntimeError('strive_b200: %s is on %s, the scene batch on %s' % (name, t.device, scene.device))
        if ext_future is not None:
            ext_future = ext_future[:, :FT, :4]
        with torch.cuda.device(scene.device):
            traj = _DecodeFn.apply(z2, self.device_model(), scene, self._env(map_env), mf, pf, ext_future, FT)
        if three_d:
            traj = traj.unsqueeze(1)
        return {'future_pred': traj}

    def encode_map(self, scene_graph, map_idx, map_env):
        """reference traffic_model.py:416-451; scene_graph.pos (NA,4) NORMALISED."""
        env = self._env(map_env)
        pose_un = self.normalizer.unnormalize(scene_graph.pos.detach()).contiguous().float()
        mapixes = map_idx[scene_graph.batch].to(torch.int32).contiguous()
        return self.encode_map_poses(pose_un, mapixes, env)

    def encode_map_poses(self, pose_un, mapixes, env):
        L = _cabi.lib()
        n = pose_un.size(0)
        out = torch.empty((n, 64), dtype=torch.float32, device=pose_un.device)
        nb = L.strive_mapenc_workspace_bytes(n)
        ws = torch.empty(nb, dtype=torch.uint8, device=pose_un.device)
        mapixes = mapixes.to(device=pose_un.device, dtype=torch.int32).contiguous()
        with torch.cuda.device(pose_un.device):
            _cabi.check(L.strive_mapenc_fwd(self.device_model().handle, C.byref(env.cstruct), _cabi.dptr(pose_un), _cabi.dptr(mapixes), n,
                                            _cabi.dptr(out), _cabi.dptr(ws), nb, _cabi.stream_ptr()))
        return out

    # ---- once-per-batch producers (PyTorch host code, reference :372-403, 453-486, 545-565) ----
    def encode_past(self, scene_graph):
        NA, PT, _ = scene_graph.past.size()
        f = scene_graph.past[:, -1, :4]
        p = scene_graph.past[:, :, :4]
        c, s = f[:, 2:3], f[:, 3:4]
        dx, dy = p[:, :, 0] - f[:, 0:1], p[:, :, 1] - f[:, 1:2]
        local = torch.stack([c * dx + s * dy, -s * dx + c * dy, p[:, :, 2] * c + p[:, :, 3] * s, p[:, :, 3] * c - p[:, :, 2] * s], 2)
        local = torch.cat([local, scene_graph.past[:, :, 4:]], 2)
        local = torch.where((scene_graph.past_vis == 0.0).unsqueeze(-1), torch.zeros_like(local), local)
        local = torch.cat([local, scene_graph.past_vis.unsqueeze(-1)], -1)
        enc_in = torch.cat([local, scene_graph.lw.unsqueeze(1).expand(NA, PT, 2)], -1)
        enc_in = torch.cat([enc_in.reshape(NA, -1), scene_graph.sem], 1)
        return self.past_encoder(enc_in)

    def prior(self, scene_graph, map_feat, past_feat):
        feat = torch.cat([past_feat, map_feat, scene_graph.sem], -1)
        out = self.prior_net(feat, scene_graph.past[:, -1, :4], scene_graph.sem, scene_graph.ptr)
        return out[:, :self.z_size], torch.exp(out[:, self.z_size:])

    def encode_future(self, scene_graph):
        """reference :488-522 (mlp trajectory encoder): future in the frame of the last past step, unobserved frames zeroed."""
        NA, FT, _ = scene_graph.future.size()
        f = scene_graph.past[:, -1, :4]
        p = scene_graph.future[:, :, :4]
        c, s = f[:, 2:3], f[:, 3:4]
        dx, dy = p[:, :, 0] - f[:, 0:1], p[:, :, 1] - f[:, 1:2]
        local = torch.stack([c * dx + s * dy, -s * dx + c * dy, p[:, :, 2] * c + p[:, :, 3] * s, p[:, :, 3] * c - p[:, :, 2] * s], 2)
        local = torch.cat([local, scene_graph.future[:, :, 4:]], 2)
        local = torch.where((scene_graph.future_vis == 0.0).unsqueeze(-1), torch.zeros_like(local), local)
        local = torch.cat([local, scene_graph.future_vis.unsqueeze(-1)], -1)
        enc_in = torch.cat([local, scene_graph.lw.unsqueeze(1).expand(NA, FT, 2)], -1)
        enc_in = torch.cat([enc_in.reshape(NA, -1), scene_graph.sem], 1)
        return self.future_encoder(enc_in)

    def encoder(self, scene_graph, map_feat, past_feat, future_feat):
        """posterior q(z | past, future, map), reference :524-543."""
        feat = torch.cat([past_feat, future_feat, map_feat, scene_graph.sem], -1)
        out = self.posterior_net(feat, scene_graph.past[:, -1, :4], scene_graph.sem, scene_graph.ptr)
        return out[:, :self.z_size], torch.exp(out[:, self.z_size:])

    @staticmethod
    def _has(scene_graph, name):
        try:
            return name in scene_graph                       # torch_geometric Data / Batch
        except TypeError:
            return getattr(scene_graph, name, None) is not None

    def embed(self, scene_graph, map_idx, map_env):
        """reference :372-403; 'posterior_out' is added whenever the graph carries a future (adv_scenario_gen.py:284 reads it)."""
        scene_graph.pos = scene_graph.past[:, -1, :4]
        map_feat = self.encode_map(scene_graph, map_idx, map_env)
        past_feat = self.encode_past(scene_graph)
        mu, var = self.prior(scene_graph, map_feat, past_feat)
        out = {'prior_out': (mu, var), 'map_feat': map_feat, 'past_feat': past_feat}
        if self._has(scene_graph, 'future'):
            future_feat = self.encode_future(scene_graph)
            out['posterior_out'] = self.encoder(scene_graph, map_feat, past_feat, future_feat)
        return out

    def rsample(self, mean, var):
        return mean + torch.randn_like(mean) * torch.sqrt(var)

    def _replicated(self, scene_graph, map_idx, NS):
        """NS copies of the scene batch laid end to end (copy s owns agents [s*NA, (s+1)*NA)): scenes are independent in the
        model, so the reference's (NA, NS, ...) sample axis (traffic_model.py:352-353) is just more scenes for the kernels."""
        srcs = self._sources(scene_graph, map_idx)
        rep = self._cache_get(scene_graph, '_strive_replicated', srcs, extra=int(NS))
        if rep is None:
            class _G(object):
                pass
            g = _G()
            NA = scene_graph.past.size(0)
            ptr = scene_graph.ptr.detach().to(torch.int64)
            S = ptr.numel() - 1
            off = (torch.arange(NS, device=ptr.device, dtype=torch.int64) * NA).view(NS, 1)
            g.ptr = torch.cat([(ptr[:-1].view(1, S) + off).reshape(-1), torch.tensor([NS * NA], device=ptr.device, dtype=torch.int64)])
            g.past = scene_graph.past.detach().repeat(NS, 1, 1)
            g.lw = scene_graph.lw.detach().repeat(NS, 1)
            g.sem = scene_graph.sem.detach().repeat(NS, 1)
            g.batch = torch.repeat_interleave(torch.arange(NS * S, device=ptr.device), (g.ptr[1:] - g.ptr[:-1]))
            rep = (g, map_idx.detach().repeat(NS))
            self._cache_put(scene_graph, '_strive_replicated', srcs, rep, extra=int(NS))
        return rep

    def sample_batched(self, scene_graph, map_idx, map_env, num_samples, include_mean=False, nfuture=None):
        """reference :319-370.  The NS sampled rollouts run as ONE kernel rollout over NS x NA agents (no grad)."""
        NA = scene_graph.past.size(0)
        NS = int(num_samples)
        emb = self.embed(scene_graph, map_idx, map_env)
        mu, var = emb['prior_out']
        z = self.rsample(mu.unsqueeze(0).expand(NS, NA, -1), var.unsqueeze(0).expand(NS, NA, -1))
        if include_mean:
            z[-1] = mu
        g_rep, midx_rep = self._replicated(scene_graph, map_idx, NS)
        emb_rep = {'map_feat': emb['map_feat'].detach().repeat(NS, 1), 'past_feat': emb['past_feat'].detach().repeat(NS, 1)}
        with torch.no_grad():
            fut = self.decode_embedding(z.reshape(NS * NA, -1).contiguous(), emb_rep, g_rep, midx_rep, map_env, nfuture=nfuture)['future_pred']
        FT = fut.size(1)
        fut = fut.view(NS, NA, FT, 4).transpose(0, 1).contiguous()
        dist = torch.distributions.Normal(mu.unsqueeze(0), torch.sqrt(var).unsqueeze(0))
        return {'prior_out': (mu, var), 'z_samp': z.transpose(0, 1), 'future_pred': fut,
                'z_logprob': dist.log_prob(z).sum(-1).transpose(0, 1),
                'z_mdist': torch.norm((z - mu.unsqueeze(0)) / torch.sqrt(var).unsqueeze(0), dim=-1).transpose(0, 1)}
