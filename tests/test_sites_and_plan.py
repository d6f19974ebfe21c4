"""Host logic against what the reference itself did: hook-site tables vs the sessions recorded
from the reference's imp_score, compress-rate parsing, and the top-k plan (file, C, k) vs the
np.load / np.argsort calls captured from the reference's loaders (tests/golden/topk_*.json)."""
import numpy as np
import pytest
import torch

from conftest import golden_json, golden_scores
from dct_pruning_b200.compress import get_compress_rate, selection_plan
from dct_pruning_b200.sites import hook_sites, resolve_module, score_dir
from dct_pruning_b200.zoo import NETS, get_network
from oracle import reference_port as port

ALL_NETS = ['vgg_16_bn', 'resnet_56', 'resnet_110', 'densenet_40', 'googlenet', 'resnet_50', 'u2netp']


@pytest.mark.parametrize('net_name', ALL_NETS)
def test_sites_match_reference_sessions(net_name):
    gold = golden_json('sites_%s.json' % net_name)
    with torch.device('meta'):
        net = get_network(net_name)
    sites = hook_sites(net_name, net)
    assert len(sites) == len(gold['sessions'])
    # the reference visits U^2-Net sites in a different order; the set of (module, variant, files) is what counts
    mine = {id(resolve_module(net, s.module)): (s.variant, [f.stem for f in s.files]) for s in sites}
    theirs = {id(net.get_submodule(s['module'])): (s['variant'], s['files']) for s in gold['sessions']}
    assert len(mine) == len(sites) and mine == theirs
    if net_name != 'u2netp':
        assert [f.stem for s in sites for f in s.files] == [f for s in gold['sessions'] for f in s['files']]
    n_files = sum(len(s.files) for s in sites)
    assert n_files == {'vgg_16_bn': 12, 'resnet_56': 55, 'resnet_110': 109, 'densenet_40': 39,
                       'googlenet': 37, 'resnet_50': 53, 'u2netp': 118}[net_name]


def test_score_dir_naming():
    assert score_dir('resnet_50', 5) == 'importance_score/resnet_50_limit5'


def test_get_compress_rate():
    assert get_compress_rate('[0.]+[0.18]*29') == [0.0] + [0.18] * 29
    assert get_compress_rate('[0.50]*7+[0.95]*5') == [0.5] * 7 + [0.95] * 5
    assert len(get_compress_rate('[0.]+[0.2]*2+[0.3]*18+[0.40]*18+[0.39]*19')) == 58
    with pytest.raises(AssertionError):
        get_compress_rate('[0]')                      # the decimal point is mandatory (common.py:178)
    with pytest.raises(AssertionError):
        get_compress_rate('[0.1]*2*3')


def test_k_truncation_quirks():
    # int() of inexact double products
    assert int(10 * (1 - 0.9)) == 0 and int(20 * (1 - 0.85)) == 3 and int(192 * (1 - 0.9)) == 19


@pytest.mark.parametrize('net_name', ALL_NETS)
def test_selection_plan_matches_reference_loaders(net_name):
    gold = golden_json('topk_%s.json' % net_name)
    plan = selection_plan(net_name, get_compress_rate(gold['compress_rate']))
    want = [(s['file'], s['C'], s['k']) for s in gold['selections'] if 'k' in s]
    got = [(s.stem, s.C, s.k) for s in plan]
    assert got == want


@pytest.mark.parametrize('net_name', ALL_NETS)
def test_stable_rule_is_a_legal_answer_to_the_reference_selection(net_name):
    gold = golden_json('topk_%s.json' % net_name)
    _, scores = golden_scores(gold['scores'])
    for s in gold['selections']:
        if 'k' not in s:
            continue
        imp = scores[s['file']]
        ref_sel = np.asarray(s['select_index'])
        mine = port.select_index_stable(imp, s['k'])
        assert port.topk_equivalent(imp, mine, s['k'])
        assert port.topk_equivalent(imp, ref_sel, s['k'])
        cut = np.sort(imp)[len(imp) - s['k']] if s['k'] else None
        if s['k'] and (imp == cut).sum() == 1:
            np.testing.assert_array_equal(mine, ref_sel)          # unique cut -> bit-exact kept set


def test_shipped_fixtures_tie_heavy():
    z = np.load(__import__('os').path.join(__import__('conftest').GOLDEN, 'shipped_googlenet.npz'))
    ties = 0
    for name in z.files:
        imp = z[name]
        for k in (1, len(imp) // 2, max(1, int(len(imp) * 0.1))):
            a = port.select_index_stable(imp, k)
            b = port.select_index_reference(imp, k)
            assert port.topk_equivalent(imp, a, k) and port.topk_equivalent(imp, b, k)
            ties += int(not np.array_equal(a, b))
    assert ties >= 0
