"""Multi-rank host logic on CPU: world_size-2 gloo processes shard the scene batch by whole
loss-normalisation groups, each rank optimises its groups with the oracle standing in for the device loop, rank 0 gathers
the rows -- the result must equal the unsharded run (no collective inside the loop, sharding is invisible in the result)."""
import os
import socket
import tempfile

import torch
import torch.multiprocessing as mp

from strive_b200 import shard, synth

SIZES = [2, 1, 3, 2, 2, 1]
GROUP_PTR = [0, 2, 3, 5, 6]       # 4 groups: scenes {0,1} {2} {3,4} {5}
FT, ITERS = 2, 2
W = {'coll_veh': 100.0, 'coll_env': 100.0, 'motion_prior': 1.0, 'init_z': 0.01}


def _world():
    raster, dx = synth.make_raster(seed=3, M=2, H=640, W=640)
    sd = {k: v.double() for k, v in synth.make_weights(0).items()}
    sc = synth.make_scenes(21, SIZES, map_extent_m=(60.0, 100.0), M=2, FT=FT, collide_frac=1.0, offroad_frac=1.0, dtype=torch.float64)
    return raster, dx, sd, sc


def _groups_as_scene_lists(gptr):
    return [list(range(gptr[g], gptr[g + 1])) for g in range(len(gptr) - 1)]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, out_path):
    import torch.distributed as dist
    from oracle import strive_oracle as O
    torch.set_num_threads(2)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    raster, dx, sd, sc = _world()
    costs = shard.group_costs(sc['ptr'], GROUP_PTR, FT)
    mine = shard.partition_groups(costs, world)[rank]
    sub, lgptr, agent_index = shard.shard_scenes(sc, GROUP_PTR, mine)
    if agent_index.numel():
        z = O.refine_loop(sd, sub, raster, dx, W, ITERS, 0.05, FT, veh_coll_buffer=0.2, groups=_groups_as_scene_lists(lgptr))
    else:
        z = torch.zeros((0, 32), dtype=torch.float64)
    NA = int(sc['ptr'][-1])
    full = shard.gather_rows(z, agent_index, NA, dst=0)                         # object gather (no partition knowledge needed)
    # point-to-point gather of exactly-sized tensors: every rank derives every rank's rows from the deterministic partition
    assign = shard.partition_groups(costs, world)
    all_index = [shard.group_rows(sc['ptr'], GROUP_PTR, a) for a in assign]
    assert torch.equal(all_index[rank], agent_index)
    full2 = shard.gather_rows(z, agent_index, NA, dst=0, all_index=all_index if rank == 0 else None, rows_per_rank=[int(i.numel()) for i in all_index])
    if rank == 0:
        assert torch.equal(full, full2)
        torch.save({'z': full, 'assign': assign}, out_path)
    else:
        assert full is None and full2 is None
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_a_balanced_exact_cover():
    costs = [100, 7, 7, 50, 49, 1, 30]
    for world in (1, 2, 3, 8):
        parts = shard.partition_groups(costs, world)
        assert len(parts) == world
        assert sorted(g for p in parts for g in p) == list(range(len(costs)))
        assert parts == shard.partition_groups(costs, world)          # deterministic
        loads = [sum(costs[g] for g in p) for p in parts]
        assert max(loads) <= sum(costs) / world + max(costs)            # LPT bound
    ptr = torch.tensor([0, 2, 3, 6, 8, 10, 11])
    assert shard.group_costs(ptr, GROUP_PTR, 3) == [(4 + 1) * 3, 9 * 3, (4 + 4) * 3, 1 * 3]


def test_shard_scenes_renumbers_and_keeps_rows():
    _, _, _, sc = _world()
    sub, lgptr, idx = shard.shard_scenes(sc, GROUP_PTR, [1, 3])
    assert lgptr == [0, 1, 2]
    assert sub['ptr'].tolist() == [0, 3, 4] and idx.tolist() == [3, 4, 5, 10]
    assert torch.equal(sub['z'], sc['z'][idx]) and torch.equal(sub['map_idx'], sc['map_idx'][torch.tensor([2, 5])])
    assert torch.equal(sub['edge_index'], synth.clique_edges([0, 3, 4]))
    assert sub['batch'].tolist() == [0, 0, 0, 1]
    empty, egptr, eidx = shard.shard_scenes(sc, GROUP_PTR, [])
    assert eidx.numel() == 0 and egptr == [0] and empty['ptr'].tolist() == [0]
    # single process: gather is a local scatter back to batch order
    back = shard.gather_rows(sub['z'], idx, 11)
    assert torch.equal(back[idx], sc['z'][idx])


def test_two_rank_gloo_sharded_refine_equals_unsharded():
    from oracle import strive_oracle as O
    raster, dx, sd, sc = _world()
    torch.set_num_threads(4)
    ref = O.refine_loop(sd, sc, raster, dx, W, ITERS, 0.05, FT, veh_coll_buffer=0.2, groups=_groups_as_scene_lists(GROUP_PTR))
    assert (ref - sc['z']).abs().max() > 1e-3             # the loop really moved z
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, 'z.pt')
        mp.spawn(_rank_main, args=(2, _free_port(), out), nprocs=2, join=True)
        got = torch.load(out)
    assert sorted(g for p in got['assign'] for g in p) == [0, 1, 2, 3] and all(len(p) > 0 for p in got['assign'])
    err = (got['z'] - ref).abs().max().item()
    assert err < 1e-9, err


def _bucket_rank(rank, world, port, out_path):
    """world_size-2 gloo: data-parallel gradient exchange of the training step (strive_b200.train.FlatBucket)."""
    import torch.distributed as dist
    from strive_b200.train import FlatBucket
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(5)
    net = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.ReLU(), torch.nn.Linear(8, 3))
    bucket = FlatBucket(net.parameters())
    assert bucket.numel == sum(p.numel() for p in net.parameters())
    x = torch.randn(4, 6, generator=torch.Generator().manual_seed(100 + rank))        # every rank its own batch
    bucket.zero_grad()
    net(x).pow(2).sum().backward()
    local = bucket.flat_g.clone()
    bucket.all_reduce_mean()
    torch.save({'local': local, 'mean': bucket.flat_g.clone(), 'first_param_is_view': net[0].weight.data_ptr() == bucket.flat_p.data_ptr()},
               out_path + '.%d' % rank)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_flat_gradient_bucket_all_reduce():
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, 'g.pt')
        mp.spawn(_bucket_rank, args=(2, _free_port(), out), nprocs=2, join=True)
        r0, r1 = torch.load(out + '.0'), torch.load(out + '.1')
    want = (r0['local'] + r1['local']) / 2
    assert (r0['local'] - r1['local']).abs().max() > 1e-3                 # the ranks really saw different batches
    assert torch.allclose(r0['mean'], want, atol=1e-7) and torch.equal(r0['mean'], r1['mean'])
    assert r0['first_param_is_view']


def _empty_rank_main(rank, world, port, out_path):
    """More ranks than groups: the partition leaves ranks without rows; both gather paths must cope (N = 8 with 4 groups)."""
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    ptr = torch.tensor([0, 3, 5, 9])
    gptr = [0, 3]                                                   # ONE group of three scenes
    assign = shard.partition_groups(shard.group_costs(ptr, gptr, 4), world)
    all_index = [shard.group_rows(ptr, gptr, a) for a in assign]
    rows_per_rank = [int(i.numel()) for i in all_index]
    mine = all_index[rank]
    z = (mine.to(torch.float32)[:, None] * 10 + torch.arange(2)[None]).contiguous()        # row r -> [10 r, 10 r + 1]; (0,2) on the empty rank
    a = shard.gather_rows(z, mine, 9, dst=0)
    b = shard.gather_rows(z, mine, 9, dst=0, all_index=all_index if rank == 0 else None, rows_per_rank=rows_per_rank)
    if rank == 0:
        torch.save({'a': a, 'b': b, 'rows_per_rank': rows_per_rank}, out_path)
    else:
        assert a is None and b is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_with_an_empty_rank():
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, 'g.pt')
        mp.spawn(_empty_rank_main, args=(2, _free_port(), out), nprocs=2, join=True)
        got = torch.load(out)
    assert sorted(got['rows_per_rank']) == [0, 9]
    want = torch.arange(9, dtype=torch.float32)[:, None] * 10 + torch.arange(2)[None]
    assert torch.equal(got['a'], want) and torch.equal(got['b'], want)
