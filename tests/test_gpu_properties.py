"""Size-independent properties at BASELINE's full sizes (ResNet-50@224 batch 256 layer shapes,
U^2-Netp@320 maps), where the CPU oracle would take hours:

  Parseval      the orthonormal DCT preserves energy: score == sum of squares (fp64 on the device)
  homogeneity   score(a*x) == a^2 * score(x) for a power-of-two a (exact in floating point)
  additivity    accumulating two half batches == one whole batch; splitting the batch over "ranks" likewise
  permutation   permuting images leaves per-channel sums unchanged up to fp64 summation order
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

FULL = [(256, 64, 56, 56), (256, 256, 56, 56), (256, 128, 28, 28), (256, 512, 28, 28), (256, 256, 14, 14),
        (256, 1024, 14, 14), (256, 512, 7, 7), (256, 2048, 7, 7), (256, 16, 32, 32), (256, 64, 8, 8),
        (12, 64, 160, 160), (4, 16, 320, 320), (12, 64, 80, 80), (12, 64, 40, 40), (12, 64, 20, 20), (12, 64, 10, 10)]


def activations(shape, device, seed=0):
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn(*shape, generator=g, device=device)
    x = torch.relu_(x)
    x[:, ::7] = 0.0
    return x


@pytest.mark.parametrize('shape', FULL)
def test_parseval_at_full_size(lib, cuda_device, shape):
    from dct_pruning_b200.ops import dct_energy
    x = activations(shape, cuda_device, seed=shape[1])
    acc, en, _ = dct_energy(x, want_energy=True)
    want = (x.double() ** 2).sum(dim=(2, 3))
    rel = ((en.double() - want).abs() / want.clamp_min(1e-30))[want > 0]
    assert float(rel.max()) < 5e-5          # worst single map of up to 524 288; the bar is 1e-4
    assert bool((en[want == 0] == 0).all())
    tot = want.sum(0)
    relc = ((acc - tot).abs() / tot.clamp_min(1e-30))[tot > 0]
    assert float(relc.max()) < 2e-5


@pytest.mark.parametrize('shape', [(64, 256, 56, 56), (64, 2048, 7, 7), (64, 512, 28, 28), (4, 64, 80, 80)])
def test_homogeneity_and_additivity(lib, cuda_device, shape):
    from dct_pruning_b200.ops import dct_energy
    x = activations(shape, cuda_device, seed=1)
    acc, en, _ = dct_energy(x, want_energy=True)
    _, en4, _ = dct_energy(x * 4.0, want_energy=True)
    assert torch.equal(en4, en * 16.0)                       # power-of-two scaling is exact end to end
    half = shape[0] // 2
    whole, _, _ = dct_energy(x)
    acc2, _, _ = dct_energy(x[:half])
    acc2, _, _ = dct_energy(x[half:], accum=acc2)
    np.testing.assert_allclose(acc2.cpu().numpy(), whole.cpu().numpy(), rtol=1e-13)
    # (with per-map energies requested a map's fp32 energy is formed first; without, the kernels add their fp64 shares of a
    #  map straight into the channel sum: the two agree to fp32 rounding of a map's energy)
    np.testing.assert_allclose(whole.cpu().numpy(), acc.cpu().numpy(), rtol=1e-6)
    perm = torch.randperm(shape[0], device=cuda_device)
    accp, enp, _ = dct_energy(x[perm].contiguous(), want_energy=True)
    assert torch.equal(enp, en[perm])                        # per-map energies are bit-reproducible
    np.testing.assert_allclose(accp.cpu().numpy(), acc.cpu().numpy(), rtol=1e-13)


def test_rank_count_independence_emulated(lib, cuda_device):
    """1/2/4/8-way batch sharding gives the same fp32 scores (each shard accumulates in fp64,
    shards are summed as the all-reduce would)."""
    from dct_pruning_b200.generate import rank_slice
    from dct_pruning_b200.ops import dct_energy, finalize
    x = activations((40, 96, 28, 28), cuda_device, seed=5)
    outs = []
    for world in (1, 2, 4, 8):
        total = torch.zeros(96, dtype=torch.float64, device=cuda_device)
        for r in range(world):
            lo, hi = rank_slice(40, r, world)
            acc, _, _ = dct_energy(x[lo:hi])
            total += acc
        outs.append(finalize(total, 40).cpu().numpy())
    for o in outs[1:]:
        np.testing.assert_array_equal(o, outs[0])
