"""TEST INFRASTRUCTURE ONLY - CPU oracle, never imported by the product path.

Restatement of the two alternative per-slice scoring lines the reference keeps commented out
beside the DCT in get_feature_hook (/root/reference/utils/common.py:262-277) and compares
against in chart*.py:

    :268  c = torch.tensor([torch.matrix_rank(output[i,j,:,:]).item() for i in range(a) for j in range(b)])
    :269  c = [dct.dct_3d(output[i,:,:,:], norm='ortho') for i in range(a)]

each followed by the unchanged tail of the hook (:271-277): cnt_score, view(a, -1), sum(0),
running mean.  `torch.matrix_rank` was removed from torch (this image: 2.11 raises); its
documented successor `torch.linalg.matrix_rank` applies the same default rule - singular values
of the fp32 matrix, rank = #{S > S.max() * max(rows, cols) * eps_fp32} - and is what runs here.

PARITY UNPINNED by the reference itself: the lines are comments, the reference ships no output
of either.  The rank rule is cross-checked against numpy's `matrix_rank` (same rule, LAPACK
gesdd) and dct_3d against scipy's 3-D `dctn(norm='ortho')` in tests/test_oracle.py.
"""
import numpy as np
import torch

from . import torch_dct_port as dct
from .reference_port import cnt_score


def matrix_rank(slice2d):
    """torch.matrix_rank(x) of common.py:268 (tol=None): count of singular values above S.max() * max(H, W) * eps."""
    return int(torch.linalg.matrix_rank(slice2d.float()).item())


def singular_values(slice2d):
    return torch.linalg.svdvals(slice2d.float())


def rank_gap(slice2d, factor=4.0):
    """True when no singular value lies within `factor` of the rank cut: there the rank does not depend on which
    backward-stable SVD computed it (singular values move by O(eps * S.max()) between algorithms, the cut sits at
    max(H, W) * eps * S.max())."""
    s = singular_values(slice2d)
    if s.numel() == 0 or float(s.max()) == 0.0:
        return True
    cut = float(s.max()) * max(slice2d.shape) * float(torch.finfo(torch.float32).eps)
    return not bool(((s > cut / factor) & (s < cut * factor)).any())


def hook_rank(state, through_cnt_score=False):
    """get_feature_hook with line 268 in place of line 267.  through_cnt_score=False is HRank's own hook (the rank tensor
    goes straight to view/sum); True leaves the reference's next line `c = cnt_score(c)` (:271) in place, which squares
    every entry (cnt_score: sum(d.mul(d)) of a 0-d tensor)."""
    def hook(module, inputs, output):
        a, b = output.shape[0], output.shape[1]
        c = torch.tensor([matrix_rank(output[i, j, :, :]) for i in range(a) for j in range(b)])
        if through_cnt_score:
            c = cnt_score([t for t in c])
        c = c.view(a, -1).float().sum(0)
        state.update(c, a)
    return hook


def hook_dct3(state):
    """get_feature_hook with line 269 in place of line 267: one 3-D coefficient cube per image, cnt_score makes ONE
    number of it, so the score vector of the site has a single entry."""
    def hook(module, inputs, output):
        a = output.shape[0]
        c = [dct.dct_3d(output[i, :, :, :], norm='ortho') for i in range(a)]
        c = cnt_score(c).view(a, -1).sum(0)
        state.update(c, a)
    return hook


def rank_values(x):
    """[B, C] int64 ranks of an NCHW tensor (per-slice torch rule)."""
    B, C = x.shape[:2]
    return torch.tensor([[matrix_rank(x[i, j]) for j in range(C)] for i in range(B)], dtype=torch.int64)


def dct3_energy64(x):
    """[B] float64 energies of the 3-D orthonormal DCT of every image (scipy, double precision)."""
    from scipy.fft import dctn
    x = np.asarray(x, dtype=np.float64)
    return np.array([np.sum(dctn(x[i], norm='ortho') ** 2) for i in range(x.shape[0])])
