"""Segmented top-k kernel: bit-exact against stable argsort (the documented tie rule), a legal
answer to the reference's own selection always, identical to it whenever the cut is unique.
Inputs: selections captured from the reference's loaders (tests/golden/topk_*.json), the
tie-heavy score files the reference ships, and adversarial vectors."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_json, golden_scores
from oracle import reference_port as port

pytestmark = pytest.mark.gpu

ALL_NETS = ['vgg_16_bn', 'resnet_56', 'resnet_110', 'densenet_40', 'googlenet', 'resnet_50', 'u2netp']


@pytest.mark.parametrize('net_name', ALL_NETS)
def test_kept_channels_match_reference_loaders(lib, cuda_device, net_name):
    from dct_pruning_b200.topk import kept_channels
    gold = golden_json('topk_%s.json' % net_name)
    _, scores = golden_scores(gold['scores'])
    kept = kept_channels(net_name, gold['compress_rate'], scores, device=cuda_device)
    sels = [s for s in gold['selections'] if 'k' in s]
    assert len(kept) == len(sels)
    exact = 0
    for (sel, idx), ref in zip(kept, sels):
        imp = scores[ref['file']]
        assert sel.stem == ref['file'] and sel.k == ref['k'] and idx.dtype == np.int64
        np.testing.assert_array_equal(idx, port.select_index_stable(imp, ref['k']))        # bit-exact vs the rule
        assert port.topk_equivalent(imp, idx, ref['k'])
        cut = np.sort(imp)[len(imp) - ref['k']] if ref['k'] else None
        if ref['k'] == 0 or (imp == cut).sum() == 1:
            np.testing.assert_array_equal(idx, np.asarray(ref['select_index'], np.int64))  # bit-exact vs the reference
            exact += 1
    assert exact > 0


def test_shipped_tie_heavy_files(lib, cuda_device):
    from dct_pruning_b200.topk import topk_segmented
    z = np.load(os.path.join(GOLDEN, 'shipped_googlenet.npz'))
    vecs, offsets, ks = [], [0], []
    for name in z.files:
        imp = z[name]
        for k in sorted({0, 1, len(imp) // 2, max(1, int(len(imp) * 0.1)), len(imp) - 1, len(imp)}):
            vecs.append(imp)
            offsets.append(offsets[-1] + len(imp))
            ks.append(k)
    flat = torch.from_numpy(np.concatenate(vecs)).to(cuda_device)
    out = topk_segmented(flat, offsets, ks)
    for imp, k, got in zip(vecs, ks, out):
        got = got.cpu().numpy()
        np.testing.assert_array_equal(got, port.select_index_stable(imp, k))
        assert port.topk_equivalent(imp, port.select_index_reference(imp, k), k)


def test_adversarial_vectors(lib, cuda_device):
    from dct_pruning_b200.topk import select_index, topk_segmented
    rng = np.random.default_rng(0)
    cases = [np.zeros(17, np.float32), np.ones(300, np.float32),
             np.array([0.0, -0.0, 0.0, 1.0, -0.0], np.float32),
             np.array([np.nan, 1.0, np.inf, -np.inf, 2.0, np.nan], np.float32),
             rng.integers(0, 4, 2048).astype(np.float32),
             rng.standard_normal(8192).astype(np.float32),
             rng.integers(0, 3, 10000).astype(np.float32),           # longer than the shared-memory key buffer
             np.array([5.0], np.float32)]
    for imp in cases:
        for k in sorted({0, 1, len(imp) // 3, len(imp)}):
            got = select_index(imp, k, device=cuda_device)
            want = port.select_index_stable(np.where(imp == 0, 0.0, imp).astype(np.float32), k)
            np.testing.assert_array_equal(got, want, err_msg='C=%d k=%d' % (len(imp), k))
    assert topk_segmented(torch.zeros(0, device=cuda_device), [0], []) == []
    empty = topk_segmented(torch.zeros(4, device=cuda_device), [0, 0, 4], [0, 2])     # an empty segment among others
    assert empty[0].numel() == 0 and empty[1].cpu().tolist() == [2, 3]
