"""world_size-2 gloo test (CPU) of the host-side multi-GPU logic: contiguous batch shards with no
data-path collective and the all-gather of stacked trajectory outputs (config 5)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kinematics_jl_b200.planning import gather_packed, gather_stacked, shard_range


def test_shard_ranges_tile_the_batch():
    for n in (0, 1, 7, 4096 * 64, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, n_wp, n_dof, n_coll = 6, 4, 3, 5
    a, b = shard_range(P, rank, world)
    g = torch.Generator().manual_seed(0)
    vals = torch.rand((P, n_wp, n_coll), generator=g, dtype=torch.float64)
    grads = torch.rand((P, n_wp, n_dof, n_coll), generator=g, dtype=torch.float64)
    v_all, g_all = gather_stacked(vals[a:b], grads[a:b])
    ok = torch.equal(v_all, vals) and torch.equal(g_all, grads)
    # the packed form: one SoA slab [vals rows | grads rows] per rank, ONE collective, views of the gathered slab
    Pl = b - a
    slab = torch.cat([vals[a:b].permute(2, 0, 1).reshape(n_coll, Pl * n_wp),
                      grads[a:b].permute(3, 2, 0, 1).reshape(n_coll * n_dof, Pl * n_wp)]).contiguous()
    v_pk, g_pk = gather_packed(slab, n_wp, n_dof, n_coll)
    ok = ok and v_pk.shape == (world, Pl, n_wp, n_coll) and g_pk.shape == (world, Pl, n_wp, n_dof, n_coll)
    ok = ok and torch.equal(v_pk.reshape(P, n_wp, n_coll), vals) and torch.equal(g_pk.reshape(P, n_wp, n_dof, n_coll), grads)
    t = torch.tensor([1.0 + rank])                      # max-over-ranks reduction used by bench.py
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[rank] = bool(ok) and float(t) == float(world)
    dist.destroy_process_group()


def test_gather_stacked_world2_gloo():
    world = 2
    port = 29500 + os.getpid() % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert all(out[r] for r in range(world))
