"""CPU suite, part 3: the N > 1 path (member sharding + the single final gather) with
world_size-2 gloo process groups -- the step path itself has no collective to test."""

import os

import numpy as np
import pytest


def test_shard_ranges_partition_the_ensemble():
    from continuum_robot_b200.sharding import shard_range, shard_sizes

    for B in (1, 7, 64, 65536, 65537):
        for W in (1, 2, 3, 8):
            r = [shard_range(B, k, W) for k in range(W)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[k][1] == r[k + 1][0] for k in range(W - 1))
            assert max(shard_sizes(B, W)) - min(shard_sizes(B, W)) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, B, out):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from continuum_robot_b200 import ensembles as ens
        from continuum_robot_b200.sharding import gather_members, local_slice, shard_range

        e = ens.config3(B, 4, seed=11)
        x = np.concatenate([e.q0, e.v0], axis=1)
        lo, hi = shard_range(B, rank, world)
        mine = torch.from_numpy(local_slice(x))
        assert mine.shape[0] == hi - lo
        mine = mine * 2.0 + 1.0  # stand-in for the (collective-free) per-shard integration
        full = gather_members(mine, B, dst=0)
        if rank == 0:
            np.save(out, full.numpy())
        else:
            assert full is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [10, 7])
def test_final_gather_world_size_2_gloo(tmp_path, B):
    import torch.multiprocessing as mp

    from continuum_robot_b200 import ensembles as ens

    out = str(tmp_path / "full.npy")
    port = 29500 + (os.getpid() % 2000) + B
    mp.spawn(_worker, args=(2, port, B, out), nprocs=2, join=True)
    e = ens.config3(B, 4, seed=11)
    ref = np.concatenate([e.q0, e.v0], axis=1) * 2.0 + 1.0
    assert np.array_equal(np.load(out), ref)
