"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Third-party shims that let the UNMODIFIED reference sources under /root/reference/src
import and run on CPU in the build container (recipe: SURVEY.md App. B).  Used by
oracle/gen_golden.py to produce the committed fixtures in tests/golden/ and by
oracle self-checks.  /root/reference does not exist on the GPU box, so nothing in
tests -m gpu / bench.py / smoke() may import this module.

What is shimmed (all absent from this image, requirements.txt:1-14 of the reference):
  * matplotlib, nuscenes, pyquaternion  -> empty stub modules (only imported, never
    called on the latent-optimisation path).
  * torch_geometric.nn.MessagePassing   -> ~40-line gather / max-aggregate / update
    restatement of PyG 1.7.1 `propagate` for flow='source_to_target', aggr='max'
    (call site: reference src/models/interaction_net.py:136).
  * torch_geometric.data.{Data,Batch}   -> attribute dict supporting `in`.
  * numpy aliases np.int/np.bool/np.float removed in numpy 2 (adv_gen_nusc.py:258).
"""
import sys
import types
import inspect

import numpy as np
import torch

REF_SRC = '/root/reference/src'


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class MessagePassing(torch.nn.Module):
    """aggr='max', flow='source_to_target': edge_index[0]=source j, edge_index[1]=target i."""

    def __init__(self, aggr='max', flow='source_to_target'):
        super().__init__()
        assert aggr == 'max' and flow == 'source_to_target'

    def propagate(self, edge_index, **kw):
        msg_args = inspect.signature(self.message).parameters
        src, dst = edge_index[0], edge_index[1]
        call = {}
        for name in msg_args:
            base, suffix = name[:-2], name[-2:]
            v = kw.get(base)
            if v is None:
                call[name] = None
            elif suffix == '_i':
                call[name] = v.index_select(0, dst)
            else:
                call[name] = v.index_select(0, src)
        msg = self.message(**call)
        N = kw['x'].size(0)
        D = msg.size(1)
        out = torch.zeros((N, D), dtype=msg.dtype, device=msg.device)
        if msg.size(0) > 0:
            idx = dst.view(-1, 1).expand(-1, D)
            out = out.scatter_reduce(0, idx, msg, 'amax', include_self=False)
            has_in = torch.zeros(N, dtype=torch.bool, device=msg.device)
            has_in[dst] = True
            out = torch.where(has_in.view(-1, 1), out, torch.zeros_like(out))
        upd_args = [p for p in inspect.signature(self.update).parameters if p != 'aggr_out']
        return self.update(out, **{k: kw.get(k) for k in upd_args})


class Data(object):
    """Minimal torch_geometric.data.Data / Batch stand-in (attribute bag with `in`)."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def __contains__(self, k):
        return k in self.__dict__ and self.__dict__[k] is not None

    def keys(self):
        return list(self.__dict__.keys())

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(device))
        return self

    @property
    def num_graphs(self):
        return int(self.ptr.numel() - 1)


def install():
    """Install the shims and put the reference on sys.path. Idempotent."""
    if getattr(install, '_done', False):
        return
    for alias, ty in (('int', int), ('bool', bool), ('float', float)):
        if not hasattr(np, alias):
            setattr(np, alias, ty)
    mpl = _stub('matplotlib', use=lambda *a, **k: None)
    _stub('matplotlib.pyplot')
    _stub('matplotlib.patches')
    _stub('matplotlib.collections')
    _stub('matplotlib.animation')
    mpl.pyplot = sys.modules['matplotlib.pyplot']
    _stub('nuscenes')
    _stub('nuscenes.nuscenes', NuScenes=object)
    _stub('nuscenes.map_expansion')
    _stub('nuscenes.map_expansion.map_api', NuScenesMap=object)
    _stub('nuscenes.map_expansion.arcline_path_utils', discretize_lane=None)
    _stub('nuscenes.utils')
    _stub('nuscenes.utils.splits', create_splits_scenes=None)
    _stub('pyquaternion', Quaternion=object)
    tg = _stub('torch_geometric')
    tg.nn = _stub('torch_geometric.nn', MessagePassing=MessagePassing)
    tg.data = _stub('torch_geometric.data', Data=Data, Batch=Data)
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    install._done = True


def make_map_env(raster, dx):
    """Synthetic NuScenesMapEnv without the nuScenes devkit."""
    install()
    from datasets.map_env import NuScenesMapEnv
    me = object.__new__(NuScenesMapEnv)
    me.nusc_raster = raster
    me.nusc_dx = dx
    me.map_list = ['synthetic-%d' % i for i in range(raster.size(0))]
    me.bounds = [-17.0, -38.5, 60.0, 38.5]
    me.L = 256
    me.W = 256
    me.num_layers = raster.size(1)
    me.device = raster.device
    return me


def make_ref_model(nfuture=20, npast=4, nclasses=2):
    install()
    from models.traffic_model import TrafficModel
    from datasets.utils import MeanStdNormalizer, NUSC_BIKE_PARAMS
    model = TrafficModel(npast, nfuture, 256, nclasses, conv_channel_in=4)
    state_mean = torch.tensor([0.0, 0.0, 0.0, 0.0, 1.802009, -0.000037])
    state_std = torch.tensor([15.0, 15.0, 1.0, 1.0, 3.507907, 0.055684])
    att_mean = torch.tensor([4.844294, 2.021752])
    att_std = torch.tensor([1.084860, 0.299647])
    model.set_normalizer(MeanStdNormalizer(state_mean, state_std))
    model.set_att_normalizer(MeanStdNormalizer(att_mean, att_std))
    model.set_bicycle_params(NUSC_BIKE_PARAMS)
    return model
