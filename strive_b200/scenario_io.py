"""Scenario files written after the adversarial / solution optimisation and read back by the evaluation tools: the JSON schema
of the reference (writer src/utils/scenario_gen.py:189-254 `prepare_output_dict`, reader src/datasets/utils.py:10-38
`read_adv_scenes`), so scenarios produced with strive_b200 open in the reference's `eval_adv_gen.py` / `viz_scenario_dir.py` and
vice versa.  Plain host code; trajectories are stored UNNORMALISED, lists of Python floats."""
import glob
import json
import os

import torch

# key -> (argument, whether it is a normalised state trajectory)
_TRAJ_KEYS = (('fut_init', 'init_fut_traj'), ('fut_adv', 'adv_fut_traj'), ('fut_internal_ego', 'internal_ego_traj'), ('fut_sol', 'sol_fut_traj'))


def _lst(t):
    return t.detach().cpu().numpy().tolist()


def prepare_output_dict(scene_graph, map_idx, map_env, dt, model, init_fut_traj, adv_fut_traj, sol_fut_traj=None, attack_agt=None,
                        attack_t=None, adv_z=None, sol_z=None, prior_distrib=None, attack_bike_params=None, internal_ego_traj=None):
    """Same signature and output as the reference writer; key order follows it as well (json.dump keeps insertion order)."""
    nrm, att = model.get_normalizer(), model.get_att_normalizer()
    given = dict(init_fut_traj=init_fut_traj, adv_fut_traj=adv_fut_traj, internal_ego_traj=internal_ego_traj, sol_fut_traj=sol_fut_traj)
    out = {'N': int(init_fut_traj.size(0)), 'dt': dt, 'map': map_env.map_list[map_idx]}
    out['lw'] = _lst(att.unnormalize(scene_graph.lw))
    out['sem'] = _lst(scene_graph.sem)
    out['past'] = _lst(nrm.unnormalize(scene_graph.past_gt))
    for key, arg in _TRAJ_KEYS:
        if given[arg] is not None:
            out[key] = _lst(nrm.unnormalize(given[arg]))
    if attack_agt is not None:
        out['attack_agt'] = int(attack_agt)
    if attack_t is not None:
        out['attack_t'] = int(attack_t)
    if adv_z is not None:
        out['z_adv'] = _lst(adv_z)
    if sol_z is not None:
        out['z_sol'] = _lst(sol_z)
    if prior_distrib is not None:
        out['z_prior'] = {'mean': _lst(prior_distrib[0]), 'var': _lst(prior_distrib[1])}
    if attack_bike_params is not None:
        out['attack_bike_prof'] = _lst(attack_bike_params)
    return out


def write_scenario(path, out_dict):
    with open(path, 'w') as f:
        json.dump(out_dict, f)


def read_adv_scenes(scene_path):
    """Every *.json of a directory -> list of scene dicts with tensors (reference reader semantics: `scene_fut` is the adversarial
    future, `attack_t` / `sem` only when present)."""
    scenes = []
    for fpath in sorted(glob.glob(os.path.join(scene_path, '*.json'))):
        with open(fpath, 'r') as f:
            j = json.load(f)
        if j is None:
            continue
        sc = {'name': os.path.basename(fpath)[:-5], 'map': j['map'], 'dt': j['dt'], 'veh_att': torch.tensor(j['lw']),
              'scene_past': torch.tensor(j['past']), 'scene_fut': torch.tensor(j['fut_adv'])}
        if 'attack_t' in j:
            sc['attack_t'] = j['attack_t']
        if 'sem' in j:
            sc['sem'] = torch.tensor(j['sem'])
        scenes.append(sc)
    return scenes
