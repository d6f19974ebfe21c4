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


def session_worker(rank, world, port_no, out_dir, give_layout_to_all):
    """ScoreSession's end-of-run collective with an EMPTY shard on rank 1 (batch size smaller than the world size):
    the rank whose hooks never fired must still enter the all-reduce - with zeros and count 0 - instead of raising before it."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port_no), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from dct_pruning_b200 import dist as ddist
    from dct_pruning_b200.hooks import ScoreSession
    from dct_pruning_b200.zoo import get_network
    ddist.init_from_env(backend='gloo')
    torch.manual_seed(0)
    net = get_network('resnet_56').eval()
    session = ScoreSession(net, 'resnet_56')
    if give_layout_to_all:
        session.plan_layout(torch.zeros(1, 3, 32, 32))           # shape-only forward: slots for all 55 sites, in site order
        assert session.used == 2032 and all(s is not None for s in session.slots)
    if rank == 0 and give_layout_to_all:                         # this rank "scored" 3 images (the kernels themselves need a GPU)
        session.flat[:session.used] = torch.arange(session.used, dtype=torch.float64) + 1.0
        session.images = [3] * len(session.sites)
    result = 'ok'
    try:
        session.register()                                       # (refuses a multi-rank run without a planned layout)
        session.remove()
        n_images = session.reduce_sums()
        np.save(os.path.join(out_dir, 'flat%d.npy' % rank), session.flat[:session.used + 1].numpy())
        assert n_images == 3.0
    except RuntimeError as e:
        result = 'raised: %s' % e
    with open(os.path.join(out_dir, 'result%d.txt' % rank), 'w') as f:
        f.write(result)
    dist.barrier()                                               # nobody is left behind in a collective
    ddist.shutdown()


def test_empty_shard_enters_the_all_reduce(tmp_path):
    mp.spawn(session_worker, args=(2, free_port(), str(tmp_path), True), nprocs=2, join=True)
    assert [open(os.path.join(str(tmp_path), 'result%d.txt' % r)).read() for r in range(2)] == ['ok', 'ok']
    a, b = (np.load(os.path.join(str(tmp_path), 'flat%d.npy' % r)) for r in range(2))
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(a[:-1], np.arange(2032) + 1.0)
    assert a[-1] == 3.0


def test_multi_rank_run_without_a_planned_layout_fails_on_every_rank_up_front(tmp_path):
    """Without plan_layout a rank with an empty shard would hold no accumulator and could not enter the all-reduce (the others
    would wait in it forever): the session refuses to register its hooks on every rank, before any forward pass."""
    mp.spawn(session_worker, args=(2, free_port(), str(tmp_path), False), nprocs=2, join=True)
    res = [open(os.path.join(str(tmp_path), 'result%d.txt' % r)).read() for r in range(2)]
    assert all(r.startswith('raised') and 'plan_layout' in r for r in res), res
