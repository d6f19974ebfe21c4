"""End-to-end importance generation on the GPU against the REAL reference's outputs.

tests/golden/scores_*.npz were produced by the reference's imp_score on CPU (one forward sweep
per hook site).  Here the same seeded nets and inputs go through cuDNN (fp32, TF32 off) with all
hook sites live in ONE sweep and the CUDA scoring kernels behind them.  Tolerance: relative 1e-4
on every live channel (north_star); channels the reference scores as exactly 0 (dead post-ReLU)
must be 0 here too unless cuDNN's forward rounds a pre-activation across zero, which is counted
and bounded rather than hidden.
"""
import types

import numpy as np
import pytest
import torch

from conftest import golden_scores

pytestmark = pytest.mark.gpu

CASES = ['vgg_16_bn_b3_l2', 'resnet_56_b2_l2', 'resnet_110_b1_l1', 'densenet_40_b2_l1', 'googlenet_b2_l1',
         'resnet_50_s64_b2_l1', 'resnet_50_s224_b1_l1', 'u2netp_s64_b1_l2', 'u2netp_s144_b1_l1']


def generate(tag, device, path='auto'):
    from dct_pruning_b200.generate import imp_score
    from dct_pruning_b200.zoo import get_network
    meta, want = golden_scores(tag)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(meta['seed'])
    net = get_network(meta['net']).to(device).eval()
    args = types.SimpleNamespace(net=meta['net'], limit=meta['limit'], batch_size=meta['batch'],
                                 input_side=meta['side'], seed_base=meta['batch_seed_base'])
    got = imp_score(net, args, write=False, path=path)
    return want, got


@pytest.mark.parametrize('tag', CASES)
def test_scores_match_reference(lib, cuda_device, tag):
    want, got = generate(tag, cuda_device)
    assert sorted(want) == sorted(got)
    flipped = total = 0
    for stem in want:
        w, g = want[stem].astype(np.float64), got[stem].astype(np.float64)
        assert g.shape == w.shape and got[stem].dtype == np.float32
        scale = max(w.max(), 1e-30)
        live = w > 1e-6 * scale                      # channels carrying real energy
        rel = np.abs(g[live] - w[live]) / w[live]
        assert rel.size == 0 or rel.max() < 1e-4, (stem, rel.max())
        # (near-)dead channels: absolute agreement at the layer's scale; exact zeros counted
        assert np.abs(g[~live] - w[~live]).max(initial=0) < 1e-6 * scale + 1e-12, stem
        flipped += int(((w == 0) != (g == 0)).sum())
        total += w.size
    assert flipped <= max(1, total // 200), 'cuDNN vs CPU forward flipped %d of %d dead channels' % (flipped, total)


def test_written_files_are_byte_compatible(lib, cuda_device, tmp_path):
    import os
    from dct_pruning_b200.generate import write_score_files
    want, got = generate('resnet_56_b2_l2', cuda_device)
    d = write_score_files(got, str(tmp_path / 'importance_score' / 'resnet_56_limit2'))
    names = sorted(os.listdir(d))
    assert names == sorted(s + '.npy' for s in want)
    for fn in names:
        raw = open(os.path.join(d, fn), 'rb').read()
        arr = np.load(os.path.join(d, fn))
        assert raw[:8] == b'\x93NUMPY\x01\x00' and len(raw) == 128 + 4 * arr.shape[0]
        assert arr.dtype == np.dtype('<f4') and arr.ndim == 1


def test_simt_and_umma_paths_agree_end_to_end(lib, cuda_device):
    _, a = generate('googlenet_b2_l1', cuda_device, path='auto')
    _, b = generate('googlenet_b2_l1', cuda_device, path='simt')
    for stem in a:
        np.testing.assert_allclose(a[stem], b[stem], rtol=3e-5, atol=0)
