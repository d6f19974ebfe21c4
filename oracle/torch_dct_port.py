"""TEST INFRASTRUCTURE ONLY - CPU oracle, never imported by the product path.

Restatement of the third-party DCT the reference calls but does not vendor:
``torch_dct.dct_2d(x, norm='ortho')`` (PyPI ``torch-dct``, GitHub zh217/torch-dct;
the reference pins no version - there is no requirements file - so this follows
the published 0.1.x algorithm).  Call site in the reference:
/root/reference/utils/common.py:28 (import) and :267 (per-slice call).

PARITY UNPINNED: the reference ships no known-answer vector for this routine and
the package itself is absent from this image, so the restatement is anchored on
(a) the algorithm as published, (b) agreement with two independent transforms
(`scipy.fft.dctn(norm='ortho')`, `cv2.dct`) and the explicit basis-matrix form,
checked in tests/test_oracle.py.

Algorithm (Makhoul's N-point DCT-II via one length-N complex FFT), fp32:
  v      = [x[0], x[2], ..., x[N-1 or N-2], ..., x[3], x[1]]   (evens, then odds reversed)
  V      = FFT(v)
  X[k]   = Re(V[k]) * cos(-pi k / 2N) - Im(V[k]) * sin(-pi k / 2N)
  ortho:   X[0] /= 2 sqrt(N);  X[k>0] /= 2 sqrt(N/2);  X *= 2
2-D = 1-D along the last axis, transpose, 1-D again, transpose back.
"""
import numpy as np
import torch


def dct(x, norm=None):
    shape = x.shape
    n = shape[-1]
    rows = x.contiguous().view(-1, n)
    v = torch.cat([rows[:, ::2], rows[:, 1::2].flip([1])], dim=1)
    spec = torch.view_as_real(torch.fft.fft(v, dim=1))
    ang = -torch.arange(n, dtype=x.dtype, device=x.device)[None, :] * np.pi / (2 * n)
    out = spec[:, :, 0] * torch.cos(ang) - spec[:, :, 1] * torch.sin(ang)
    if norm == 'ortho':
        out[:, 0] /= np.sqrt(n) * 2
        out[:, 1:] /= np.sqrt(n / 2) * 2
    return 2 * out.view(*shape)


def dct_2d(x, norm=None):
    once = dct(x, norm=norm)
    twice = dct(once.transpose(-1, -2), norm=norm)
    return twice.transpose(-1, -2)


def dct_3d(x, norm=None):
    """torch_dct.dct_3d: the 1-D transform along each of the last three axes in turn
    (call site in the reference: /root/reference/utils/common.py:269, commented out beside dct_2d)."""
    a = dct(x, norm=norm)
    b = dct(a.transpose(-1, -2), norm=norm)
    c = dct(b.transpose(-1, -3), norm=norm)
    return c.transpose(-1, -3).transpose(-1, -2)
