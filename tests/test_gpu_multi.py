"""The real N>1 path: one process per GPU, NCCL, batch sharding with `rank_slice`, ONE all-reduce of the flat fp64 sums.
Needs two GPUs (skipped otherwise; the round's multi-GPU run is `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).

The result must not depend on the rank count: identical kept-channel sets and scores within 1e-6 of the single-process
run on the same seeded inputs - also when a rank's shard of every batch is empty (batch size 1 over two ranks)."""
import json
import os
import socket
import types

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

RATE = '[0.]+[0.18]*29'


def free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def score(net_name, batch, limit, device):
    from dct_pruning_b200.generate import imp_score
    from dct_pruning_b200.zoo import get_network
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    torch.manual_seed(0)
    net = get_network(net_name).to(device).eval()
    args = types.SimpleNamespace(net=net_name, limit=limit, batch_size=batch, seed_base=1000)
    return imp_score(net, args, write=False)


def worker(rank, world, port_no, net_name, batch, limit, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port_no), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from dct_pruning_b200 import dist as ddist
    r, local, w = ddist.init_from_env(backend='nccl')
    files = score(net_name, batch, limit, torch.device('cuda', local))
    if r == 0:
        np.savez(os.path.join(out_dir, 'scores.npz'), **files)
    ddist.shutdown()


@pytest.mark.parametrize('batch,limit', [(8, 2), (5, 2), (1, 3)])
def test_two_ranks_over_nccl_match_one_process(lib, cuda_device, tmp_path, batch, limit):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    from dct_pruning_b200.topk import kept_channels
    want = score('resnet_56', batch, limit, cuda_device)
    mp.spawn(worker, args=(2, free_port(), 'resnet_56', batch, limit, str(tmp_path)), nprocs=2, join=True)
    got = dict(np.load(os.path.join(str(tmp_path), 'scores.npz')))
    assert sorted(got) == sorted(want)
    for stem in want:
        w, g = want[stem].astype(np.float64), got[stem].astype(np.float64)
        assert ((w == 0) == (g == 0)).all(), stem
        # (a rank's slice of a batch goes through cuDNN as its own, smaller batch: other algorithms, last-bit differences in the
        #  activations; the scoring itself adds nothing, see test_gpu_properties.test_rank_count_independence_emulated)
        np.testing.assert_allclose(g, w, rtol=1e-6, atol=0, err_msg=stem)
    kw = kept_channels('resnet_56', RATE, want, device=cuda_device)
    kg = kept_channels('resnet_56', RATE, got, device=cuda_device)
    assert len(kw) == len(kg) == 45
    for (sw, iw), (sg, ig) in zip(kw, kg):
        assert sw.stem == sg.stem and np.array_equal(iw, ig), sw.stem
    with open(os.path.join(str(tmp_path), 'summary.json'), 'w') as f:
        json.dump({'batch': batch, 'limit': limit, 'files': len(want)}, f)
