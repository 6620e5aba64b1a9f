"""ctypes binding of include/strive_b200.h (the C-ABI of libstrive_b200.so).

The product path has NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised
(the reference drivers catch RuntimeError to skip a batch, src/refine_traffic_optim.py:381-388).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libstrive_b200.so')
_lib = None

c_f32p = C.c_void_p
STRIVE_TERMS = 16
LOSS_AVOID, LOSS_ADV, LOSS_MATCH = 1, 2, 4


class StriveScene(C.Structure):
    _fields_ = [('num_agents', C.c_int32), ('num_scenes', C.c_int32), ('max_scene_agents', C.c_int32),
                ('num_classes', C.c_int32), ('ptr', C.c_void_p), ('scene_of', C.c_void_p), ('map_idx', C.c_void_p),
                ('past_last', C.c_void_p), ('lw', C.c_void_p), ('sem', C.c_void_p)]


class StriveMap(C.Structure):
    _fields_ = [('raster', C.c_void_p), ('dx', C.c_void_p), ('M', C.c_int32), ('C', C.c_int32), ('H', C.c_int32),
                ('W', C.c_int32), ('lin_l', C.c_void_p), ('lin_w', C.c_void_p), ('packed', C.c_void_p), ('packed_pitch', C.c_int32)]


class StriveLossCfg(C.Structure):
    _fields_ = [('kind', C.c_int32), ('traj_unnormalized', C.c_int32), ('num_groups', C.c_int32),
                ('group_agent_ptr', C.c_void_p), ('group_of', C.c_void_p), ('agent_map', C.c_void_p),
                ('group_zrows', C.c_void_p), ('group_match_rows', C.c_void_p),
                ('cblock_ptr', C.c_void_p), ('cblock_of', C.c_void_p),
                ('w_coll_veh', C.c_float), ('w_coll_env', C.c_float), ('w_motion_prior', C.c_float), ('w_init_z', C.c_float),
                ('w_coll_veh_plan', C.c_float), ('w_init_z_atk', C.c_float), ('w_motion_prior_atk', C.c_float),
                ('w_adv_crash', C.c_float), ('w_match_ext', C.c_float), ('w_motion_prior_ext', C.c_float),
                ('veh_coll_buffer', C.c_float), ('single_veh_idx', C.c_int32), ('crash_min_t', C.c_int32),
                ('use_infront', C.c_int32), ('crash_min_infront', C.c_float),
                ('attack_mask', C.c_void_p), ('adv_min_out', C.c_void_p),
                ('env_L', C.c_void_p), ('env_W', C.c_void_p), ('env_lin_l', C.c_void_p), ('env_lin_w', C.c_void_p),
                ('circ_cx', C.c_void_p), ('lw_un', C.c_void_p), ('adv_own_pred', C.c_int32), ('reserved0', C.c_int32)]


EXPORTS = ['strive_last_error', 'strive_abi_version', 'strive_struct_layout', 'strive_profile_enable', 'strive_profile_report', 'strive_tc_selftest', 'strive_tc_selftest_pair', 'strive_tc_trace', 'strive_tc_debug', 'strive_model_layout', 'strive_model_create', 'strive_model_destroy', 'strive_model_tc_bytes', 'strive_model_set_tc_weights', 'strive_mapenc_set_impl', 'strive_mapenc_set_split', 'strive_mapenc_set_pair', 'strive_model_edge_frag_bytes', 'strive_model_set_edge_frags', 'strive_edge_set_impl', 'strive_set_pdl',
           'strive_mapenc_workspace_bytes', 'strive_mapenc_fwd', 'strive_map_crop', 'strive_decode_tape_bytes',
           'strive_decode_fwd', 'strive_decode_bwd', 'strive_decode_bwd_pair', 'strive_decode_tape_read', 'strive_loss_workspace_bytes',
           'strive_loss_fwd_bwd', 'strive_adam_step', 'strive_adam_step_dev', 'strive_on_layer_frac', 'strive_line_layer', 'strive_veh_iou_hits']


def lib():
    """Load libstrive_b200.so (built in-tree by __graft_entry__.build()); raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError('strive_b200: %s is missing -- the CUDA extension has not been built '
                           '(run `python -c "import __graft_entry__ as g; g.build()"`); there is no CPU fallback' % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    L.strive_last_error.restype = C.c_char_p
    L.strive_abi_version.restype = C.c_int
    L.strive_model_layout.argtypes = [C.c_int, C.POINTER(i64), C.c_int, C.POINTER(C.c_int)]
    L.strive_model_create.argtypes = [vp, i64, C.POINTER(i64), C.c_int, C.c_int, C.POINTER(vp)]
    L.strive_model_destroy.argtypes = [vp]
    L.strive_model_destroy.restype = None
    L.strive_model_tc_bytes.restype = i64
    L.strive_model_set_tc_weights.argtypes = [vp, vp, i64]
    L.strive_mapenc_set_impl.argtypes = [C.c_int]
    L.strive_mapenc_set_split.argtypes = [C.c_int]
    L.strive_mapenc_set_pair.argtypes = [C.c_int]
    L.strive_model_edge_frag_bytes.restype = i64
    L.strive_model_set_edge_frags.argtypes = [vp, vp, i64, vp]
    L.strive_edge_set_impl.argtypes = [C.c_int]
    L.strive_set_pdl.argtypes = [C.c_int]
    L.strive_mapenc_workspace_bytes.argtypes = [i32]
    L.strive_mapenc_workspace_bytes.restype = i64
    L.strive_mapenc_fwd.argtypes = [vp, C.POINTER(StriveMap), vp, vp, i32, vp, vp, i64, vp]
    L.strive_map_crop.argtypes = [C.POINTER(StriveMap), vp, vp, i32, vp, vp]
    L.strive_decode_tape_bytes.argtypes = [i32, i32]
    L.strive_decode_tape_bytes.restype = i64
    L.strive_decode_fwd.argtypes = [vp, C.POINTER(StriveScene), C.POINTER(StriveMap), vp, vp, vp, vp, i32, vp, vp, i64, vp]
    L.strive_decode_bwd.argtypes = [vp, C.POINTER(StriveScene), i32, vp, vp, vp, vp, i64, vp]
    L.strive_decode_bwd_pair.argtypes = [vp, C.POINTER(StriveScene), i32, vp, vp, vp, vp, vp, vp, i64, vp]
    L.strive_decode_tape_read.argtypes = [vp, i32, i32, C.c_char_p, i32, vp, vp]
    L.strive_loss_workspace_bytes.argtypes = [i32, i32, i32]
    L.strive_loss_workspace_bytes.restype = i64
    L.strive_loss_fwd_bwd.argtypes = [C.POINTER(StriveLossCfg), C.POINTER(StriveScene), C.POINTER(StriveMap), i32,
                                      vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]
    L.strive_adam_step.argtypes = [vp, vp, vp, vp, vp, i64, i32, f32, f32, f32, f32, vp]
    L.strive_adam_step_dev.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, i64, vp, f32, f32, f32, f32, vp]
    L.strive_on_layer_frac.argtypes = [C.POINTER(StriveMap), i32, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp]
    L.strive_line_layer.argtypes = [C.POINTER(StriveMap), i32, vp, vp, vp, vp, i32, i32, vp, vp, vp]
    L.strive_veh_iou_hits.argtypes = [vp, vp, i32, vp, vp, i32, i32, C.c_double, vp, vp, vp]
    L.strive_struct_layout.argtypes = [C.POINTER(i64), C.c_int]
    L.strive_profile_enable.argtypes = [C.c_int]
    L.strive_tc_selftest.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.strive_tc_selftest_pair.argtypes = [vp, vp, vp, vp, i32, vp]
    L.strive_tc_trace.argtypes = [vp, C.c_int]
    L.strive_tc_debug.argtypes = [C.c_int]
    L.strive_profile_report.argtypes = [C.c_char_p, i64]
    L.strive_profile_report.restype = i64
    if L.strive_abi_version() != 1:
        raise RuntimeError('strive_b200: ABI version mismatch')
    _verify_layout(L)
    if os.environ.get('STRIVE_MAPENC_SPLIT') is not None:   # development switch: half-chunk pipeline of the map encoder
        L.strive_mapenc_set_split(int(os.environ['STRIVE_MAPENC_SPLIT']))
    if os.environ.get('STRIVE_MAPENC_PAIR') is not None:    # A/B switch: conv3 on CTA pairs (tcgen05 cta_group::2)
        L.strive_mapenc_set_pair(int(os.environ['STRIVE_MAPENC_PAIR']))
    if os.environ.get('STRIVE_PDL') is not None:      # development switch: bit 0 rollout kernels, bit 1 map-encoder kernels
        L.strive_set_pdl(int(os.environ['STRIVE_PDL']))
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise RuntimeError('strive_b200 [%d]: %s' % (rc, lib().strive_last_error().decode()))


def dptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('strive_b200: expected a CUDA tensor (no CPU fallback)')
    if not t.is_contiguous():
        raise RuntimeError('strive_b200: expected a contiguous tensor')
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError('strive_b200: expected dtype %s, got %s' % (dtype, t.dtype))
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _verify_layout(L):
    buf = (C.c_int64 * 16)()
    n = L.strive_struct_layout(buf, 16)
    mine = [C.sizeof(StriveScene), StriveScene.ptr.offset, StriveScene.sem.offset, C.sizeof(StriveMap), StriveMap.lin_l.offset,
            C.sizeof(StriveLossCfg), StriveLossCfg.group_agent_ptr.offset, StriveLossCfg.w_coll_veh.offset,
            StriveLossCfg.attack_mask.offset, StriveLossCfg.lw_un.offset]
    if n != len(mine) or list(buf[:n]) != mine:
        raise RuntimeError('strive_b200: ctypes struct layout %s does not match the library %s' % (mine, list(buf[:max(n, 0)])))


def tc_trace(reset=True):
    """[4][8] cycle counters of the tensor-core conv pipelines (see include/strive_b200.h: strive_tc_trace)."""
    buf = (C.c_uint64 * 32)()
    if lib().strive_tc_trace(C.cast(buf, C.c_void_p), int(bool(reset))) != 0:
        raise RuntimeError('strive_tc_trace failed')
    return [[int(buf[k * 8 + j]) for j in range(8)] for k in range(4)]


def profile_enable(on):
    lib().strive_profile_enable(int(bool(on)))


def profile_report():
    """{kernel_name: (launches, total_ms)} since the last report (synchronises the device)."""
    buf = C.create_string_buffer(1 << 16)
    n = lib().strive_profile_report(buf, len(buf))
    if n < 0:
        raise RuntimeError('strive_b200: profile report buffer too small')
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split()
        out[name] = (int(cnt), float(ms))
    return out


def set_edge_impl(impl):
    """Edge-phase kernels of the rollout: 3 / True (default) = tcgen05 forward + backward (scenes up to 129 agents), 2 = tcgen05 forward +
    mma.sync backward, 1 = mma.sync TF32 kernels both ways, 0 / False = fp32 SIMT kernels (A/B verification)."""
    if isinstance(impl, bool):
        impl = 3 if impl else 0
    lib().strive_edge_set_impl(int(impl))


def set_mapenc_impl(tensor_core):
    """True (default): tcgen05 map encoder; False: fp32 SIMT kernels (A/B verification)."""
    lib().strive_mapenc_set_impl(int(bool(tensor_core)))
