"""Device-side containers the C-ABI calls take: packed model handle, scene batch, raster map.

Host code stays PyTorch (north_star): these classes only own CUDA tensors and build the C structs.
"""
import ctypes as C

import torch

from . import _cabi
from .weights import pack_decode_weights, pack_tc_weights

CROP_BOUNDS = (-17.0, -38.5, 60.0, 38.5)     # reference src/datasets/map_env.py:23


def _i32(t, device):
    return t.to(device=device, dtype=torch.int32).contiguous()


class DeviceModel(object):
    """Opaque packed-weights handle (strive_model_create / strive_model_destroy)."""

    def __init__(self, state_dict, num_classes, device):
        L = _cabi.lib()
        blob, sizes = pack_decode_weights(state_dict, num_classes)
        want = (C.c_int64 * 256)()
        n = C.c_int(0)
        _cabi.check(L.strive_model_layout(num_classes, want, 256, C.byref(n)))
        if n.value != len(sizes) or list(want[:n.value]) != sizes:
            raise RuntimeError('strive_b200: weight packing does not match the library layout')
        self.blob = blob.to(device)
        self.num_classes = num_classes
        arr = (C.c_int64 * len(sizes))(*sizes)
        h = C.c_void_p()
        tcb = pack_tc_weights(state_dict)
        if tcb.numel() != L.strive_model_tc_bytes():
            raise RuntimeError('strive_b200: tensor-core weight packing size mismatch')
        with torch.cuda.device(self.blob.device):       # the library works on the CURRENT device: make it the model's
            _cabi.check(L.strive_model_create(_cabi.dptr(self.blob), self.blob.numel(), arr, len(sizes), num_classes, C.byref(h)))
            self.handle = h
            self.tc_blob = tcb.to(device)
            _cabi.check(L.strive_model_set_tc_weights(h, _cabi.dptr(self.tc_blob), self.tc_blob.numel()))
            # mma.sync fragment packs of the edge MLP, packed on the device from the blob (csrc/edge_mma.cuh)
            self.edge_frags = torch.empty(L.strive_model_edge_frag_bytes(), dtype=torch.uint8, device=device)
            _cabi.check(L.strive_model_set_edge_frags(h, _cabi.dptr(self.edge_frags), self.edge_frags.numel(), _cabi.stream_ptr()))

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                _cabi.lib().strive_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class MapEnv(object):
    """Raster store with the attribute surface of the reference NuScenesMapEnv that the hot path touches
    (src/datasets/map_env.py:21-203): nusc_raster (M,C,H,W) uint8, nusc_dx (M,2) float64, bounds, L, W,
    get_map_crop(scene_graph, map_idx).  Rasterising nuScenes itself is out of scope (no devkit)."""

    def __init__(self, nusc_raster, nusc_dx, bounds=CROP_BOUNDS, L=256, W=256, device='cuda'):
        if L != 256 or W != 256:
            raise RuntimeError('strive_b200: the map encoder is built for 256x256 crops (map_obs_size_pix=256)')
        self.device = torch.device(device)
        self.nusc_raster = nusc_raster.to(self.device, dtype=torch.uint8).contiguous()
        self.nusc_dx = nusc_dx.to(self.device, dtype=torch.float64).contiguous()
        self.bounds = list(bounds)
        self.L, self.W = L, W
        self.num_layers = int(self.nusc_raster.size(1))
        self.map_list = ['map-%d' % i for i in range(self.nusc_raster.size(0))]
        # nuscenes_utils.py:219-220, computed by torch so the values are bit-identical to the reference's
        self._lin_l = torch.linspace(bounds[0], bounds[2], L, dtype=torch.float32).to(self.device)
        self._lin_w = torch.linspace(bounds[1], bounds[3], W, dtype=torch.float32).to(self.device)
        M, Cc, H, Wd = self.nusc_raster.shape
        if int(self.nusc_raster.max()) > 1:
            raise RuntimeError('strive_b200: the map raster must be binary (0/1 layers as produced by get_map_mask, map_env.py:106-118)')
        # one byte per pixel, bit c = layer c: the crop gather then touches 1 byte instead of C scattered bytes
        # rows padded to a multiple of 16 bytes: crop_pack stages raster rows with 16-byte cp.async
        Wp = (Wd + 15) // 16 * 16
        self._packed = torch.zeros((M, H, Wp), dtype=torch.uint8, device=self.device)
        for c in range(min(Cc, 8)):
            self._packed[:, :, :Wd] |= (self.nusc_raster[:, c] << c)
        self.cstruct = _cabi.StriveMap(_cabi.dptr(self.nusc_raster), _cabi.dptr(self.nusc_dx), M, Cc, H, Wd,
                                       _cabi.dptr(self._lin_l), _cabi.dptr(self._lin_w), _cabi.dptr(self._packed), Wp)

    def crop_poses(self, pose_un, mapixes):
        """(N,4) unnormalised poses -> (N,C,256,256) uint8 (reference get_map_obs, nuscenes_utils.py:236-264)."""
        n = pose_un.size(0)
        out = torch.empty((n, self.num_layers, self.L, self.W), dtype=torch.uint8, device=self.device)
        pose_c, mix = pose_un.contiguous().float(), _i32(mapixes, self.device)
        with torch.cuda.device(out.device):
            _cabi.check(_cabi.lib().strive_map_crop(C.byref(self.cstruct), _cabi.dptr(pose_c), _cabi.dptr(mix), n, _cabi.dptr(out), _cabi.stream_ptr()))
        return out

    def get_map_crop(self, scene_graph, map_idx, bounds=None, L=None, W=None):
        """map_env.py:168-203; scene_graph.pos is UNNORMALISED (N,4)."""
        if bounds is not None or L is not None or W is not None:
            raise RuntimeError('strive_b200: crop geometry overrides are not supported')
        mapixes = map_idx[scene_graph.batch]
        return self.crop_poses(scene_graph.pos, mapixes)


class SceneBatch(object):
    """Device copy of what the hot path reads from the drivers' torch_geometric Batch
    (src/datasets/nuscenes_dataset.py:609-687): past[:, -1], lw, sem, ptr, batch + map_idx per scene."""

    def __init__(self, scene_graph, map_idx, device='cuda'):
        dev = torch.device(device)
        self.device = dev
        ptr_h = scene_graph.ptr.detach().cpu().to(torch.int64)
        self.ptr_host = ptr_h
        self.NA = int(ptr_h[-1])
        self.S = int(ptr_h.numel() - 1)
        sizes = ptr_h[1:] - ptr_h[:-1]
        if self.S < 1 or int(sizes.min()) < 1:
            raise RuntimeError('strive_b200: empty scene in batch')
        self.max_n = int(sizes.max())
        E_expected = int((sizes * (sizes - 1)).sum())
        ei = getattr(scene_graph, 'edge_index', None)
        if ei is not None and int(ei.size(1)) != E_expected:
            raise RuntimeError('strive_b200: edge_index is not the full directed clique per scene '
                               '(%d edges, expected %d)' % (int(ei.size(1)), E_expected))
        self.ptr = _i32(ptr_h, dev)
        self.scene_of = torch.repeat_interleave(torch.arange(self.S), sizes).to(dev, torch.int32).contiguous()
        self.map_idx = _i32(map_idx, dev)
        self.past_last = scene_graph.past[:, -1, :].detach().to(dev, torch.float32).contiguous()
        self.lw = scene_graph.lw.detach().to(dev, torch.float32).contiguous()
        self.sem = scene_graph.sem.detach().to(dev, torch.float32).contiguous()
        self.NC = int(self.sem.size(1))
        self.cstruct = _cabi.StriveScene(self.NA, self.S, self.max_n, self.NC, _cabi.dptr(self.ptr), _cabi.dptr(self.scene_of),
                                         _cabi.dptr(self.map_idx), _cabi.dptr(self.past_last), _cabi.dptr(self.lw),
                                         _cabi.dptr(self.sem))
        self.agent_map = self.map_idx[self.scene_of.long()].contiguous()
        self.ego_mask = torch.zeros(self.NA, dtype=torch.bool, device=dev)
        self.ego_mask[self.ptr[:-1].long()] = True
