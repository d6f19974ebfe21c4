"""The CPU oracle (oracle/) pinned against outputs of the REAL reference.

tests/golden/scores_*.npz were written by the reference's own `imp_score` (imported unmodified,
see tests/golden/make_golden.py).  Here the restatement in oracle/reference_port.py, driven by
this repo's site tables and this repo's nets, must reproduce them bit for bit: that pins the
port, the hook-site enumeration, the file naming and the zoo's seeded weights in one go.
The DCT arithmetic itself (absent third-party torch_dct) is cross-checked against scipy, cv2,
the explicit basis-matrix form and Parseval.
"""
import hashlib

import numpy as np
import pytest
import torch

from conftest import golden_scores
from dct_pruning_b200.generate import synthetic_batches
from dct_pruning_b200.sites import hook_sites
from dct_pruning_b200.zoo import get_network
from oracle import reference_port as port
from oracle import torch_dct_port as tdct


def state_digest(model):
    h = hashlib.sha256()
    for k, v in model.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().cpu().numpy().tobytes())
    return h.hexdigest()


def run_port(tag):
    meta, files = golden_scores(tag)
    torch.manual_seed(meta['seed'])
    net = get_network(meta['net']).eval()
    assert state_digest(net) == meta['weights_sha256'], 'zoo init diverged from the reference constructors'
    sessions = [(s.module, s.variant, [(f.stem, f.lo, f.hi) for f in s.files]) for s in hook_sites(meta['net'], net)]

    def batches():
        return [b[0] for b in synthetic_batches(meta['batch'], meta['side'], meta['limit'],
                                                seed_base=meta['batch_seed_base'])]
    return files, port.imp_score_port(net, sessions, batches, meta['limit'])


FAST = ['vgg_16_bn_b3_l2', 'resnet_56_b2_l2', 'densenet_40_b2_l1', 'googlenet_b2_l1']


@pytest.mark.parametrize('tag', FAST)
def test_port_reproduces_reference_bit_exact(tag):
    torch.set_num_threads(min(8, torch.get_num_threads()))
    want, got = run_port(tag)
    assert sorted(want) == sorted(got)
    for stem in want:
        assert got[stem].dtype == np.float32 and got[stem].shape == want[stem].shape
        np.testing.assert_array_equal(got[stem], want[stem], err_msg=stem)


@pytest.mark.slow
@pytest.mark.parametrize('tag', ['resnet_50_s64_b2_l1', 'u2netp_s64_b1_l2'])
def test_port_reproduces_reference_bit_exact_large(tag):
    want, got = run_port(tag)
    assert sorted(want) == sorted(got)
    for stem in want:
        np.testing.assert_array_equal(got[stem], want[stem], err_msg=stem)


# ----------------------------------------------------------------- the DCT arithmetic (unpinned by the reference)
def basis(n):
    k = np.arange(n)[:, None]
    m = np.arange(n)[None, :]
    c = np.cos(np.pi * (2 * m + 1) * k / (2 * n)) * np.sqrt(2.0 / n)
    c[0] *= np.sqrt(0.5)
    return c


@pytest.mark.parametrize('n', [1, 2, 4, 7, 8, 9, 10, 14, 16, 20, 28, 32, 56, 112, 144, 288, 320])
def test_torch_dct_port_matches_independent_transforms(n):
    from scipy.fft import dctn
    rng = np.random.default_rng(n)
    x = np.maximum(rng.standard_normal((n, n)), 0).astype(np.float32)
    got = tdct.dct_2d(torch.from_numpy(x), norm='ortho').numpy().astype(np.float64)
    ref = dctn(x.astype(np.float64), type=2, norm='ortho')
    scale = max(np.abs(ref).max(), 1e-30)
    assert np.abs(got - ref).max() / scale < 2e-5
    c = basis(n)
    np.testing.assert_allclose(c @ x.astype(np.float64) @ c.T, ref, atol=1e-10 * max(scale, 1))
    if n > 1:
        import cv2
        if n % 2 == 0:
            assert np.abs(cv2.dct(x).astype(np.float64) - ref).max() / scale < 2e-5
    e_ref = (ref ** 2).sum()
    assert abs((got ** 2).sum() - e_ref) <= 1e-5 * max(e_ref, 1e-30)
    assert abs(port.energy_parseval64(x) - e_ref) <= 1e-10 * max(e_ref, 1e-30)


def test_known_answers():
    n = 8
    z = tdct.dct_2d(torch.zeros(n, n), norm='ortho')
    assert float(z.abs().max()) == 0.0                                     # all-zero stays exactly zero
    const = tdct.dct_2d(torch.full((n, n), 3.0), norm='ortho').numpy()     # constant map: only DC, energy n*n*v^2
    assert abs(const[0, 0] - 3.0 * n) < 1e-5 and np.abs(const).sum() - abs(const[0, 0]) < 1e-4
    imp = torch.zeros(n, n)
    imp[2, 5] = 1.0
    zi = tdct.dct_2d(imp, norm='ortho').numpy()                            # impulse: outer product of basis columns
    c = basis(n)
    np.testing.assert_allclose(zi, np.outer(c[:, 2], c[:, 5]), atol=1e-6)
    mode = np.outer(c[3], c[1]).astype(np.float32)                         # a single cosine mode -> one coefficient
    zm = tdct.dct_2d(torch.from_numpy(mode), norm='ortho').numpy()
    assert abs(zm[3, 1] - 1.0) < 1e-5 and np.abs(zm).sum() - abs(zm[3, 1]) < 1e-4


def test_torch2dct_odd_rows_pad_keeps_energy():
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal((9, 9)).astype(np.float32))
    d = port.torch2dct(x)
    assert tuple(d.shape) == (10, 10)                                      # common.py:235-236 pads both axes
    e = float((d * d).sum())
    assert abs(e - float((x * x).sum())) <= 1e-5 * e


def test_running_mean_equals_sum_over_images():
    st = port.ScoreState()
    rng = np.random.default_rng(1)
    chunks = [torch.from_numpy(rng.random(5).astype(np.float32)) for _ in range(4)]
    for c in chunks:
        st.update(c * 3, 3)
    want = sum(c.double() * 3 for c in chunks) / 12
    np.testing.assert_allclose(st.feature_result.numpy(), want.numpy(), rtol=1e-6)


# ------------------------------------------------------------------ alternative scoring ops
def test_alt_oracle_dct3_matches_scipy_and_parseval():
    """oracle dct_3d (the 1-D port along three axes) == scipy's 3-D dctn; its energy == sum of squares (orthonormal)."""
    from scipy.fft import dctn
    from oracle import alt_ops_port as ao, torch_dct_port as dct
    x = torch.relu(torch.randn(2, 6, 9, 12, generator=torch.Generator().manual_seed(3)))
    for b in range(2):
        cube = dct.dct_3d(x[b], norm='ortho').numpy()
        np.testing.assert_allclose(cube, dctn(x[b].numpy(), norm='ortho'), atol=2e-5)
    e = ao.dct3_energy64(x.numpy())
    np.testing.assert_allclose(e, (x.double() ** 2).sum((1, 2, 3)).numpy(), rtol=1e-12)
    st = port.ScoreState()
    ao.hook_dct3(st)(None, None, x)
    assert st.feature_result.shape == (1,)
    np.testing.assert_allclose(float(st.feature_result[0]), e.mean(), rtol=1e-5)


def test_alt_oracle_rank_rule_matches_numpy():
    """The oracle's rank (torch.linalg.matrix_rank, successor of the removed torch.matrix_rank) applies the rule of the reference's
    line: S > S.max() * max(H, W) * eps.  numpy's matrix_rank states the same rule on LAPACK's singular values."""
    from oracle import alt_ops_port as ao
    g = torch.Generator().manual_seed(11)
    cases = [torch.zeros(5, 5), torch.eye(7), torch.randn(9, 3, generator=g) @ torch.randn(3, 9, generator=g),
             torch.relu(torch.randn(12, 12, generator=g) - 0.8), torch.randn(6, 10, generator=g)]
    for m in cases:
        s = np.linalg.svd(m.numpy(), compute_uv=False)
        by_rule = int((s > s.max(initial=0) * max(m.shape) * np.finfo(np.float32).eps).sum())
        assert ao.matrix_rank(m) == by_rule == int(np.linalg.matrix_rank(m.numpy()))
    x = torch.stack(cases[:1] + [torch.eye(5)])[None]
    st = port.ScoreState()
    ao.hook_rank(st)(None, None, x)
    assert st.feature_result.tolist() == [0.0, 5.0]
    st = port.ScoreState()
    ao.hook_rank(st, through_cnt_score=True)(None, None, x)
    assert st.feature_result.tolist() == [0.0, 25.0]
