"""The command line mirrors /root/reference/importance_generation.py:8-21 (same flags and defaults); the GPU test runs
it end to end: score files, kept-channel sets and the pruned state dict."""
import json
import os

import numpy as np
import pytest
import torch

from dct_pruning_b200 import cli


def test_flags_and_defaults_match_the_reference():
    args = cli.build_parser().parse_args([])
    assert (args.dataset, args.data_dir, args.batch_size, args.pretrain_dir, args.limit, args.net) == \
        ('cifar10', './data', 128, 'checkpoints/googlenet.pt', 5, 'googlenet')
    with pytest.raises(SystemExit):
        cli.build_parser().parse_args(['--net', 'alexnet'])
    with pytest.raises(SystemExit):
        cli.build_parser().parse_args(['--dataset', 'mnist'])


@pytest.mark.gpu
def test_cli_writes_scores_selections_and_pruned_weights(lib, cuda_device, tmp_path, monkeypatch):
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
        monkeypatch.delenv(k, raising=False)
    out = tmp_path / 'importance_score'
    pruned = tmp_path / 'pruned.pt'
    rate = '[0.]+[0.18]*29'
    cli.main(['--net', 'resnet_56', '--batch_size', '8', '--limit', '2', '--out_root', str(out), '--compress_rate', rate,
              '--save_pruned', str(pruned), '--pretrain_dir', str(tmp_path / 'none.pt')])
    d = out / 'resnet_56_limit2'
    assert len([f for f in os.listdir(d) if f.endswith('.npy')]) == 55
    sel = json.load(open(d / 'kept_channels.json'))['selections']
    assert len(sel) == 45 and all(len(s['select_index']) == s['k'] for s in sel)
    state = torch.load(pruned)
    w = state['layer1.0.conv1.weight']
    assert w.shape[0] == int(16 * (1 - 0.18))
    first = next(s for s in sel if s['conv'] == 'layer1.0.conv1.weight')
    scores = np.load(d / (first['file'] + '.npy'))
    assert sorted(np.argsort(scores, kind='stable')[len(scores) - first['k']:].tolist()) == first['select_index']
