"""Scene sharding of the latent-optimisation loops over ranks (one process per GPU).

Scenes are independent in the model (edges never cross scenes, reference src/datasets/nuscenes_dataset.py:678-687) but
the losses couple the scenes of one reference batch: every `.mean()` is over the colliding pairs / valid hits / agents of
the whole batch (src/losses/adv_gen_nusc.py:317-336, 229-250) and `get_coll_point` sizes its grid from batch means
(src/datasets/nuscenes_utils.py:351-354).  A rank therefore owns WHOLE loss-normalisation groups (= the batches the
reference driver would have formed, src/refine_traffic_optim.py:291-310); no collective runs inside the loop and the
result of every group is what the reference computes for that batch on its own, whatever the world size.
The only communication is the final gather of the optimised rows.
"""
import torch


def group_costs(ptr, group_scene_ptr, FT):
    """Work estimate per group: sum over its scenes of n_s^2 * FT (all-pairs edges and collision tests dominate)."""
    ptr = [int(v) for v in ptr]
    costs = []
    for g in range(len(group_scene_ptr) - 1):
        c = 0
        for s in range(int(group_scene_ptr[g]), int(group_scene_ptr[g + 1])):
            n = ptr[s + 1] - ptr[s]
            c += n * n * int(FT)
        costs.append(c)
    return costs


def partition_groups(costs, world):
    """Longest-processing-time greedy assignment of groups to ranks; deterministic (ties -> lower group id, lower rank).
    Returns `world` ascending lists of group ids; every group appears exactly once."""
    order = sorted(range(len(costs)), key=lambda g: (-costs[g], g))
    load = [0] * world
    out = [[] for _ in range(world)]
    for g in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(g)
        load[r] += costs[g]
    return [sorted(o) for o in out]


def clique_edges(ptr):
    """Full directed clique per scene, (2,E) int64 -- the edge set the reference dataset builds (nuscenes_dataset.py:678-687)."""
    src, dst = [], []
    for s in range(len(ptr) - 1):
        a, b = int(ptr[s]), int(ptr[s + 1])
        n = b - a
        if n > 1:
            i = torch.arange(a, b).repeat_interleave(n)
            j = torch.arange(a, b).repeat(n)
            keep = i != j
            src.append(j[keep])
            dst.append(i[keep])
    if not src:
        return torch.zeros((2, 0), dtype=torch.long)
    return torch.stack([torch.cat(src), torch.cat(dst)])


PER_AGENT = ('past', 'lw', 'sem', 'z', 'prior_mu', 'prior_var', 'map_feat', 'past_feat', 'future_gt', 'future_vis', 'past_vis')
PER_SCENE = ('map_idx', 'ext_future')


def group_rows(ptr, group_scene_ptr, groups):
    """Agent rows of the full batch owned by `groups` (ascending group ids), in shard order."""
    rows = []
    for g in groups:
        a, b = int(ptr[int(group_scene_ptr[g])]), int(ptr[int(group_scene_ptr[g + 1])])
        rows.append(torch.arange(a, b))
    return torch.cat(rows) if rows else torch.zeros(0, dtype=torch.long)


def shard_scenes(scene, group_scene_ptr, groups):
    """Sub-batch holding the scenes of `groups` (ascending group ids), renumbered from 0.

    scene: dict with 'ptr' (S+1) and per-agent / per-scene tensors (keys in PER_AGENT / PER_SCENE that are present).
    Returns (sub_scene, local_group_scene_ptr, agent_index) where agent_index (NA_local) int64 are the rows of the full
    batch this shard owns, in shard order."""
    ptr = scene['ptr']
    scenes = []
    local_gptr = [0]
    for g in groups:
        ss = list(range(int(group_scene_ptr[g]), int(group_scene_ptr[g + 1])))
        scenes += ss
        local_gptr.append(local_gptr[-1] + len(ss))
    rows = [torch.arange(int(ptr[s]), int(ptr[s + 1])) for s in scenes]
    agent_index = torch.cat(rows) if rows else torch.zeros(0, dtype=torch.long)
    sizes = [int(ptr[s + 1]) - int(ptr[s]) for s in scenes]
    new_ptr = torch.zeros(len(scenes) + 1, dtype=ptr.dtype)
    if sizes:
        new_ptr[1:] = torch.cumsum(torch.tensor(sizes), 0)
    sub = {'ptr': new_ptr,
           'batch': torch.repeat_interleave(torch.arange(len(scenes)), torch.tensor(sizes, dtype=torch.long)) if sizes
           else torch.zeros(0, dtype=torch.long),
           'edge_index': clique_edges(new_ptr)}
    sidx = torch.tensor(scenes, dtype=torch.long)
    for k in PER_AGENT:
        if k in scene:
            sub[k] = scene[k][agent_index.to(scene[k].device)]
    for k in PER_SCENE:
        if k in scene:
            sub[k] = scene[k][sidx.to(scene[k].device)]
    return sub, local_gptr, agent_index


def gather_rows(local_rows, agent_index, NA, dst=0, all_index=None, rows_per_rank=None):
    """Collects per-agent result rows of every rank on rank `dst` in the order of the unsharded batch.
    local_rows (NA_local, ...) on any device; returns the (NA, ...) CPU tensor on `dst`, None elsewhere.
    With torch.distributed uninitialised (single process) it is a local scatter.

    The partition is deterministic, so every rank can know every rank's rows: with `all_index` (list of index tensors, needed on
    `dst`) and `rows_per_rank` the gather is point-to-point sends of exactly-sized tensors (device tensors over NCCL / NVLink,
    CPU tensors over gloo) -- no pickling, no size exchange.  Without them it falls back to gather_object."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        rows = local_rows.detach().cpu()
        out = torch.empty((NA,) + tuple(rows.shape[1:]), dtype=rows.dtype)
        out[agent_index] = rows
        return out
    rank, world = dist.get_rank(), dist.get_world_size()
    if rows_per_rank is None or (rank == dst and all_index is None):
        rows = local_rows.detach().cpu()
        parts = [None] * world if rank == dst else None
        dist.gather_object((agent_index.cpu(), rows), parts, dst=dst)
        if rank != dst:
            return None
    else:
        nccl = dist.get_backend() == 'nccl'
        rows = local_rows.detach()
        rows = rows.cuda() if nccl else rows.cpu()
        rows = rows.contiguous()
        tail = tuple(rows.shape[1:])
        if rank != dst:
            if rows_per_rank[rank] > 0:
                dist.send(rows, dst)
            return None
        bufs, reqs = [], []
        for r in range(world):
            if r == dst or rows_per_rank[r] == 0:
                bufs.append(rows if r == dst else rows.new_zeros((0,) + tail))
                continue
            b = torch.empty((rows_per_rank[r],) + tail, dtype=rows.dtype, device=rows.device)
            reqs.append(dist.irecv(b, r))
            bufs.append(b)
        for q in reqs:
            q.wait()
        parts = [(all_index[r], bufs[r].cpu()) for r in range(world)]
        rows = rows.cpu()
    out = torch.empty((NA,) + tuple(rows.shape[1:]), dtype=rows.dtype)
    seen = torch.zeros(NA, dtype=torch.bool)
    for idx, r in parts:
        out[idx] = r
        seen[idx] = True
    if not bool(seen.all()):
        raise RuntimeError('strive_b200.shard.gather_rows: %d agent rows were not produced by any rank' % int((~seen).sum()))
    return out
