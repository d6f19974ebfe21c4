"""Pruned-weight transfer: the op plan this repo derives for a net must reproduce, tensor for tensor
and bit for bit, what the unmodified reference loaders wrote (tests/golden/transfer_<net>.json, produced by
tests/golden/make_golden.py from /root/reference/utils/load_models.py).  CPU part: plan + oracle gather.
GPU part (-m gpu): the same through the C ABI (dctp_gather_weight) with the kept ids from the GPU top-k."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from dct_pruning_b200 import transfer
from dct_pruning_b200.compress import get_compress_rate
from dct_pruning_b200.zoo import get_network
from oracle import load_models_port as port

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def load_case(net):
    with open(os.path.join(GOLDEN, 'transfer_%s.json' % net)) as f:
        gold = json.load(f)
    with open(os.path.join(GOLDEN, 'topk_%s.json' % net)) as f:
        topk = json.load(f)
    assert topk['compress_rate'] == gold['compress_rate'] and topk['scores'] == gold['scores']
    kept = {s['file']: np.asarray(s['select_index'], dtype=np.int64) for s in topk['selections'] if 'select_index' in s}
    rates = get_compress_rate(gold['compress_rate'])
    torch.manual_seed(0)
    orig = get_network(net).eval()
    torch.manual_seed(0)
    pruned = get_network(net, rates).eval()
    return gold, kept, orig, pruned


def digests(state):
    out = {}
    for k, v in state.items():
        a = v.detach().cpu().contiguous().numpy()
        out[k] = [list(a.shape), str(a.dtype), hashlib.sha256(a.tobytes()).hexdigest()]
    return out


def init_digest(d):
    h = hashlib.sha256()
    for k in d:
        h.update(k.encode())
        h.update(d[k][2].encode())
    return h.hexdigest()


def check_against_golden(gold, before, after):
    assert len(after) == gold['n_tensors']
    changed = {k: after[k] for k in after if after[k][2] != before[k][2]}
    assert sorted(changed) == sorted(gold['changed'])
    for k, want in gold['changed'].items():
        assert changed[k] == want, k
    same = hashlib.sha256(''.join(k + after[k][2] for k in after if after[k][2] == before[k][2]).encode()).hexdigest()
    assert same == gold['unchanged_digest']


def cpu_gather(w, sel_out, sel_in):
    as_np = lambda v: None if v is None else np.asarray(v, dtype=np.int64)          # noqa: E731
    return torch.from_numpy(port.gather_numpy(w.numpy(), as_np(sel_out), as_np(sel_in)))


@pytest.mark.parametrize('net', transfer.SUPPORTED)
def test_plan_reproduces_reference_loader(net):
    gold, kept, orig, pruned = load_case(net)
    before = digests(pruned.state_dict())
    assert init_digest(before) == gold['pruned_init_digest']       # this repo's pruned constructor == the reference's
    ori = {k: v.clone() for k, v in orig.state_dict().items()}
    plan = transfer.transfer_plan(net, pruned, {k: tuple(v.shape) for k, v in ori.items()})
    state = transfer.apply_plan(plan, ori, dict(pruned.state_dict()), kept, gather=cpu_gather)
    pruned.load_state_dict(state)
    check_against_golden(gold, before, digests(pruned.state_dict()))


def test_oracle_gather_forms_agree():
    rng = np.random.default_rng(0)
    w = rng.standard_normal((7, 5, 3, 3)).astype(np.float32)
    so, si = np.array([0, 2, 6]), np.array([1, 4])
    for a, b in ((so, si), (so, None), (None, si), (None, None)):
        np.testing.assert_array_equal(port.copy_loops(w, a, b), port.gather_numpy(w, a, b))
    v = rng.standard_normal(9).astype(np.float32)
    np.testing.assert_array_equal(port.copy_loops(v, so), port.gather_numpy(v, so))


def test_unsupported_net_is_an_error():
    with pytest.raises(ValueError):
        transfer.transfer_plan('alexnet', get_network('vgg_16_bn'), {})


def test_gather_has_no_cpu_path():
    with pytest.raises(RuntimeError):
        transfer.gather_weight(torch.zeros(4, 4, 3, 3))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize('shape,k_out,k_in', [((64, 32, 3, 3), 40, 20), ((16, 3, 7, 7), 9, None), ((128, 256, 1, 1), None, 100),
                                              ((10, 10), 4, 7), ((33,), 5, None), ((5, 4, 3, 3), None, None), ((6, 6, 3, 3), 0, 3)])
def test_gather_kernel_matches_oracle(lib, cuda_device, shape, k_out, k_in):
    rng = np.random.default_rng(sum(shape))
    w = rng.standard_normal(shape).astype(np.float32)
    so = None if k_out is None else np.sort(rng.choice(shape[0], k_out, replace=False)).astype(np.int64)
    si = None if k_in is None else np.sort(rng.choice(shape[1], k_in, replace=False)).astype(np.int64)
    got = transfer.gather_weight(torch.from_numpy(w).to(cuda_device),
                                 None if so is None else torch.from_numpy(so).to(cuda_device),
                                 None if si is None else torch.from_numpy(si).to(cuda_device))
    np.testing.assert_array_equal(got.cpu().numpy(), port.gather_numpy(w, so, si))
    if len(shape) == 4 and (k_out or 1) * (k_in or 1) <= 1000:
        np.testing.assert_array_equal(got.cpu().numpy(), port.copy_loops(w, so, si))


@pytest.mark.gpu
def test_gather_bad_index_is_reported(lib, cuda_device):
    from dct_pruning_b200 import _lib
    w = torch.zeros(4, 4, 3, 3, device=cuda_device)
    transfer.gather_weight(w, torch.tensor([0, 9], device=cuda_device))
    with pytest.raises(_lib.DctpError):
        _lib.check(lib.dctp_check(None))
    _lib.check(lib.dctp_check(None))                               # the status word was cleared


@pytest.mark.gpu
@pytest.mark.parametrize('net', transfer.SUPPORTED)
def test_device_transfer_reproduces_reference_loader(lib, cuda_device, net):
    from dct_pruning_b200.topk import kept_channels
    gold, kept_ref, orig, pruned = load_case(net)
    before = digests(pruned.state_dict())
    scores = np.load(os.path.join(GOLDEN, 'scores_%s.npz' % gold['scores']))
    scores = {k: scores[k] for k in scores.files if k != '__meta__'}
    kept = dict(kept_ref)
    for sel, ids in kept_channels(net, gold['compress_rate'], scores, device=cuda_device):     # GPU top-k (stable tie rule)
        assert sel.stem in kept_ref
        if len(set(scores[sel.stem].tolist())) == len(scores[sel.stem]):   # no ties: must be the reference's own selection
            np.testing.assert_array_equal(ids, kept_ref[sel.stem])
            kept[sel.stem] = ids
    pruned = pruned.to(cuda_device)
    plan = transfer.transfer_weights(net, pruned, orig.state_dict(), kept)
    assert any(op.kind == 'gather' for op in plan)
    check_against_golden(gold, before, digests(pruned.state_dict()))


@pytest.mark.gpu
def test_score_select_rebuild_fill_on_device(lib, cuda_device):
    """load_model's non-resume branch in one call: must equal scoring, top-k and transfer done step by step."""
    from dct_pruning_b200.prune import pruned_model
    from dct_pruning_b200.topk import kept_channels
    torch.backends.cudnn.allow_tf32 = False
    rate = '[0.]+[0.18]*29'
    torch.manual_seed(0)
    orig = get_network('resnet_56').to(cuda_device).eval()
    net, scores, kept = pruned_model('resnet_56', rate, orig, limit=1, batch_size=8, seed=0)
    assert len(scores) == 55 and len(kept) == 45
    torch.manual_seed(0)
    again = get_network('resnet_56', get_compress_rate(rate)).to(cuda_device).eval()
    transfer.transfer_weights('resnet_56', again, orig.state_dict(), kept_channels('resnet_56', rate, scores, device=cuda_device))
    assert digests(net.state_dict()) == digests(again.state_dict())
    with torch.no_grad():
        y = net(torch.randn(2, 3, 32, 32, device=cuda_device))
    assert y.shape == (2, 10) and torch.isfinite(y).all()


def load_iter_case():
    with open(os.path.join(GOLDEN, 'transfer_googlenet_iter.json')) as f:
        gold = json.load(f)
    kept = {s['file']: np.asarray(s['select_index'], dtype=np.int64) for s in gold['selections'] if 'select_index' in s}
    torch.manual_seed(0)
    net_a = get_network('googlenet', gold['origin_rates']).eval()
    torch.manual_seed(0)
    net_b = get_network('googlenet', gold['rates']).eval()
    return gold, kept, net_a, net_b


def test_iterative_round_reproduces_reference_loader():
    """prune_dynamic.py:150-154: the source net is itself pruned; GoogLeNet's loader then takes the source's rates (`cpr`)."""
    from dct_pruning_b200.compress import selection_plan
    gold, kept, net_a, net_b = load_iter_case()
    before = digests(net_b.state_dict())
    assert init_digest(before) == gold['pruned_init_digest']
    plan_sel = selection_plan('googlenet', gold['rates'], gold['origin_rates'])
    want = [(s['file'], s['C'], s['k']) for s in gold['selections']]
    assert [(s.stem, s.C, s.k) for s in plan_sel] == want
    ori = {k: v.clone() for k, v in net_a.state_dict().items()}
    plan = transfer.transfer_plan('googlenet', net_b, {k: tuple(v.shape) for k, v in ori.items()}, origin_rates=gold['origin_rates'])
    state = transfer.apply_plan(plan, ori, dict(net_b.state_dict()), kept, gather=cpu_gather)
    net_b.load_state_dict(state)
    check_against_golden(gold, before, digests(net_b.state_dict()))


@pytest.mark.gpu
def test_iterative_round_on_device(lib, cuda_device):
    from dct_pruning_b200.prune import pruned_model
    gold, kept_ref, net_a, net_b = load_iter_case()
    before = digests(net_b.state_dict())
    scores = np.load(os.path.join(GOLDEN, 'scores_%s.npz' % gold['scores']))
    scores = {k: scores[k] for k in scores.files}
    net, _, kept = pruned_model('googlenet', gold['rates'], net_a.to(cuda_device), scores=scores, seed=0, origin_rates=gold['origin_rates'])
    if all(np.array_equal(ids, kept_ref[sel.stem]) for sel, ids in kept):       # (no tie fell on a cut: same selections as the reference)
        check_against_golden(gold, before, digests(net.state_dict()))
    assert len(kept) == len(gold['selections'])


@pytest.mark.parametrize('net', transfer.SUPPORTED)
def test_plan_reproduces_reference_loader_second_rate(net):
    """A second compress rate per net (other k patterns), selections captured from the reference loader's own argsort calls."""
    from dct_pruning_b200.compress import selection_plan
    with open(os.path.join(GOLDEN, 'transfer_%s_alt.json' % net)) as f:
        gold = json.load(f)
    kept = {s['file']: np.asarray(s['select_index'], dtype=np.int64) for s in gold['selections'] if 'select_index' in s}
    rates = get_compress_rate(gold['compress_rate'])
    assert [(s.stem, s.C, s.k) for s in selection_plan(net, rates)] == [(s['file'], s['C'], s['k']) for s in gold['selections']]
    torch.manual_seed(0)
    orig = get_network(net).eval()
    torch.manual_seed(0)
    pruned = get_network(net, rates).eval()
    before = digests(pruned.state_dict())
    assert init_digest(before) == gold['pruned_init_digest']
    ori = {k: v.clone() for k, v in orig.state_dict().items()}
    plan = transfer.transfer_plan(net, pruned, {k: tuple(v.shape) for k, v in ori.items()})
    pruned.load_state_dict(transfer.apply_plan(plan, ori, dict(pruned.state_dict()), kept, gather=cpu_gather))
    check_against_golden(gold, before, digests(pruned.state_dict()))
