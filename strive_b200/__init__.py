"""strive_b200 -- B200-native (sm_100a) implementation of STRIVE's latent-optimisation hot path.

Public surface (mirrors the reference modules the drivers import):
    strive_b200.TrafficModel            <-> src/models/traffic_model.py TrafficModel (decode side)
    strive_b200.MapEnv                  <-> src/datasets/map_env.py NuScenesMapEnv (raster store + crop)
    strive_b200.losses.{AvoidCollLoss, AdvGenLoss, TgtMatchingLoss}  <-> src/losses/adv_gen_nusc.py
    strive_b200.optim.refine_traffic_optim / RefineLoop              <-> src/refine_traffic_optim.py:146-226
    strive_b200.metrics.{compute_coll_rate_env, check_*_veh_coll, determine_feasibility_nusc}  <-> success / plausibility checks
    strive_b200.scenario_io.{prepare_output_dict, read_adv_scenes}      <-> src/utils/scenario_gen.py:189-254, src/datasets/utils.py:10-38
    strive_b200.train.TrafficModelTrainer                              <-> src/train_traffic.py:101-112 (step) + data-parallel all-reduce
The compute path is hand-written CUDA in strive_b200/csrc behind the C-ABI of include/strive_b200.h; importing the
package does not need a GPU, calling it does (there is no CPU fallback).
"""
from .runtime import MapEnv, SceneBatch, DeviceModel          # noqa: F401
from .traffic_model import TrafficModel, MeanStdNormalizer, NUSC_BIKE_PARAMS, STATE_MEAN, STATE_STD, ATT_MEAN, ATT_STD  # noqa: F401
from . import losses, optim, synth, metrics, scenario_io      # noqa: F401

__all__ = ['TrafficModel', 'MapEnv', 'SceneBatch', 'DeviceModel', 'MeanStdNormalizer', 'losses', 'optim', 'synth', 'metrics']


def make_model(nfuture=20, npast=4, nclasses=2, state_dict=None, device='cuda'):
    """TrafficModel with the nuScenes car/truck normalisers and bicycle parameters set (what the drivers do at
    refine_traffic_optim.py:450-487)."""
    import torch
    m = TrafficModel(npast, nfuture, 256, nclasses)
    m.set_normalizer(MeanStdNormalizer(torch.tensor(STATE_MEAN), torch.tensor(STATE_STD)))
    m.set_att_normalizer(MeanStdNormalizer(torch.tensor(ATT_MEAN), torch.tensor(ATT_STD)))
    m.set_bicycle_params(NUSC_BIKE_PARAMS)
    if state_dict is not None:
        missing, unexpected = m.load_state_dict(state_dict, strict=False)
        if unexpected:
            raise RuntimeError('unexpected keys in state_dict: %s' % unexpected)
    return m.to(device)
