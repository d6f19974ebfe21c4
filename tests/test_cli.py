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
              '--save_pruned', str(pruned), '--pretrain_dir', str(tmp_path / 'none.pt'), '--synthetic', '--random_init'])
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


def test_missing_checkpoint_or_dataset_is_an_error_unless_asked_for(tmp_path):
    """The reference raises without a checkpoint (importance_generation.py:54-56) and scores real images; noise scores must
    not reach importance_score/ by accident."""
    import types
    from dct_pruning_b200.zoo import get_network
    net = get_network('resnet_56')
    args = types.SimpleNamespace(net='resnet_56', pretrain_dir=str(tmp_path / 'none.pt'), random_init=False, seed=0)
    with pytest.raises(FileNotFoundError):
        cli.load_checkpoint(net, args)
    args.random_init = True
    assert cli.load_checkpoint(net, args) is False
    data_args = types.SimpleNamespace(dataset='cifar10', data_dir=str(tmp_path), batch_size=4, synthetic=False, seed=0)
    with pytest.raises(SystemExit):
        cli.build_loader(data_args)
    data_args.synthetic = True
    assert cli.build_loader(data_args) is None


def test_checkpoint_formats_follow_the_reference(tmp_path):
    """importance_generation.py:24-53: u2netp keeps only keys present in the model, resnet_50 files are bare state dicts,
    densenet_40/resnet_110 were saved from DataParallel ('module.' prefix), the rest carry 'state_dict'."""
    import types
    from dct_pruning_b200.zoo import get_network
    torch.manual_seed(1)
    src = get_network('resnet_110')
    path = tmp_path / 'r110.pt'
    torch.save({'state_dict': {'module.' + k: v for k, v in src.state_dict().items()}}, path)
    dst = get_network('resnet_110')
    assert cli.load_checkpoint(dst, types.SimpleNamespace(net='resnet_110', pretrain_dir=str(path), random_init=False, seed=0))
    assert all(torch.equal(a, b) for a, b in zip(src.state_dict().values(), dst.state_dict().values()))
    torch.manual_seed(2)
    u = get_network('u2netp')
    state = dict(u.state_dict())
    first = next(iter(state))
    state['not.in.the.model'] = torch.zeros(3)            # extra key: dropped
    del state[first]                                      # missing key: the model keeps what it was built with
    path = tmp_path / 'u2netp.pth'
    torch.save(state, path)
    dst = get_network('u2netp')
    keep = dst.state_dict()[first].clone()
    assert cli.load_checkpoint(dst, types.SimpleNamespace(net='u2netp', pretrain_dir=str(path), random_init=False, seed=0))
    assert torch.equal(dst.state_dict()[first], keep)
    others = [k for k in u.state_dict() if k != first]
    assert all(torch.equal(u.state_dict()[k], dst.state_dict()[k]) for k in others)


def test_dataset_loaders_follow_the_reference_layout(tmp_path):
    """data.load_data on miniature ImageFolder / DUTS trees: batch shapes, dtypes and the {'image': ...} sample form the
    batch drivers expect (common.py:312-332); seeded sampling repeats."""
    import types
    from PIL import Image
    from dct_pruning_b200.data import load_data
    from dct_pruning_b200.generate import _images_of
    rng = np.random.default_rng(0)
    for cls in ('n01', 'n02'):
        d = tmp_path / 'imagenet' / 'ILSVRC2012_img_train' / cls
        d.mkdir(parents=True)
        for i in range(3):
            Image.fromarray(rng.integers(0, 255, (260, 300, 3), dtype=np.uint8)).save(d / ('%d.JPEG' % i))
    args = types.SimpleNamespace(dataset='imagenet', data_dir=str(tmp_path / 'imagenet'), batch_size=4, workers=0)
    loader, val = load_data(args, seed=3)
    x, y = next(iter(loader))
    assert x.shape == (4, 3, 224, 224) and x.dtype == torch.float32 and val is None
    x2, _ = next(iter(load_data(args, seed=3)[0]))
    assert torch.equal(x, x2)
    d = tmp_path / 'duts' / 'DUTS-TR' / 'DUTS-TR-Image'
    d.mkdir(parents=True)
    for i in range(3):
        Image.fromarray(rng.integers(0, 255, (200 + 30 * i, 250, 3), dtype=np.uint8)).save(d / ('im%d.jpg' % i))
    Image.fromarray(rng.integers(0, 255, (210, 190), dtype=np.uint8)).save(d / 'gray.jpg')
    args = types.SimpleNamespace(dataset='DUTS', data_dir=str(tmp_path / 'duts'), batch_size=4)
    sample = next(iter(load_data(args, seed=1)[0]))
    img = _images_of(sample)
    assert img.shape == (4, 3, 288, 288) and img.dtype == torch.float32
    with pytest.raises(FileNotFoundError):
        load_data(types.SimpleNamespace(dataset='cifar10', data_dir=str(tmp_path), batch_size=4))


@pytest.mark.gpu
def test_prune_entry_selects_from_existing_score_files(lib, cuda_device, tmp_path, monkeypatch):
    """--imp_score DIR --compress_rate ...: kept sets and the filled pruned net from files on disk, no --limit needed
    (prune_cifar10.py:70-71,83-88; SURVEY C-8); same selections as the scoring run's own kept_channels.json."""
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
        monkeypatch.delenv(k, raising=False)
    out = tmp_path / 'importance_score'
    rate = '[0.]+[0.18]*29'
    cli.main(['--net', 'resnet_56', '--batch_size', '8', '--limit', '1', '--out_root', str(out), '--compress_rate', rate,
              '--synthetic', '--random_init'])
    d = out / 'resnet_56_limit1'
    want = {s['file']: s['select_index'] for s in json.load(open(d / 'kept_channels.json'))['selections']}
    pruned = tmp_path / 'p.pt'
    net, kept = cli.prune_main(['--net', 'resnet_56', '--imp_score', str(d), '--compress_rate', rate, '--random_init',
                                '--save_pruned', str(pruned)])
    assert {s.stem: [int(i) for i in idx] for s, idx in kept} == want
    assert torch.load(pruned)['layer1.0.conv1.weight'].shape[0] == int(16 * (1 - 0.18))


def test_score_op_flag_and_session_argument():
    """--score_op selects one of the per-slice reductions the reference keeps beside the DCT (utils/common.py:267-269); the default
    is the DCT; an unknown op is refused before anything touches a device."""
    from dct_pruning_b200.cli import build_parser
    from dct_pruning_b200.hooks import ScoreSession
    from dct_pruning_b200.zoo import get_network
    p = build_parser()
    assert p.parse_args([]).score_op == 'dct2'
    for op in ('dct2', 'rank', 'rank_sq', 'dct3'):
        assert p.parse_args(['--score_op', op]).score_op == op
    with pytest.raises(SystemExit):
        p.parse_args(['--score_op', 'svd'])
    net = get_network('vgg_16_bn')
    assert ScoreSession(net, 'vgg_16_bn', op='rank').op == 'rank'
    with pytest.raises(ValueError):
        ScoreSession(net, 'vgg_16_bn', op='svd')
    with pytest.raises(RuntimeError):                        # no CPU path for the alternative ops either
        s = ScoreSession(net, 'vgg_16_bn', op='rank')
        s.score(0, torch.zeros(1, 4, 8, 8))
