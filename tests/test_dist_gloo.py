"""The N>1 path on CPU: world_size-2 gloo.  Batch sharding (ragged and empty shards included)
plus the single all-reduce of the flat score buffer must give the single-process answer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dct_pruning_b200.generate import rank_slice
from oracle import reference_port as port


def free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def worker(rank, world, port_no, batch_sizes, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port_no), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from dct_pruning_b200 import dist as ddist
    r, _, w = ddist.init_from_env(backend='gloo')
    assert (r, w) == (rank, world)
    C = 6
    flat = torch.zeros(C + 1, dtype=torch.float64)
    for b, bs in enumerate(batch_sizes):
        g = torch.Generator().manual_seed(1000 + b)
        x = torch.relu(torch.randn(bs, C, 8, 8, generator=g))
        lo, hi = rank_slice(bs, rank, world)
        if hi > lo:
            flat[:C] += torch.from_numpy(port.energy_scipy64(x[lo:hi].numpy()).sum(0))
        flat[C] += hi - lo
    ddist.allreduce_sums(flat)
    np.save(os.path.join(out_dir, 'rank%d.npy' % rank), flat.numpy())
    ddist.shutdown()


@pytest.mark.parametrize('batch_sizes', [[4, 4], [5, 3, 1], [1]])
def test_two_rank_sharding_matches_single_process(tmp_path, batch_sizes):
    world = 2
    mp.spawn(worker, args=(world, free_port(), batch_sizes, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(os.path.join(str(tmp_path), 'rank%d.npy' % r)) for r in range(world)]
    np.testing.assert_array_equal(got[0], got[1])
    C = 6
    want = np.zeros(C)
    for b, bs in enumerate(batch_sizes):
        g = torch.Generator().manual_seed(1000 + b)
        x = torch.relu(torch.randn(bs, C, 8, 8, generator=g))
        want += port.energy_scipy64(x.numpy()).sum(0)
    np.testing.assert_allclose(got[0][:C], want, rtol=1e-13)
    assert got[0][C] == sum(batch_sizes)


def test_rank_slice_partitions_every_batch():
    for n in (0, 1, 5, 8, 256):
        for world in (1, 2, 3, 4, 8):
            parts = [rank_slice(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
