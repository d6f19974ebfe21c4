"""Bring-up check of one score kernel on one shape: energies and coefficients vs float64, first mismatches printed.
    python tools/diag_stack.py N [B C [path]]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dct_pruning_b200.ops import dct_energy            # noqa: E402
from scipy.fft import dctn                             # noqa: E402

n = int(sys.argv[1])
B, C = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (2, 5)
path = sys.argv[4] if len(sys.argv) > 4 else ('stack' if n > 8 else 'kron')
g = torch.Generator().manual_seed(n)
x = torch.relu(torch.randn(B, C, n, n, generator=g))
dev = torch.device('cuda', 0)
try:
    acc, en, co = dct_energy(x.to(dev), path=path, want_energy=True, want_coeff=True)
except Exception as e:                                 # noqa: BLE001
    print('N=%d FAILED: %s' % (n, e))
    sys.exit(0)
z = dctn(x.numpy().astype(np.float64), type=2, norm='ortho', axes=(-2, -1))
want = (z * z).sum((-2, -1))
en = en.cpu().numpy()
co = co.cpu().numpy()
rel = np.abs(en - want) / np.maximum(want, 1e-30)
cerr = np.abs(co - z).max() / np.abs(z).max()
print(path, 'N=%d B=%d C=%d: energy max rel err %.3e, coeff max err %.3e, accum ok %s' % (
    n, B, C, rel.max(), cerr, np.allclose(acc.cpu().numpy(), en.astype(np.float64).sum(0), rtol=1e-12)))
if rel.max() > 2e-5 or cerr > 2e-5:
    bad = np.argwhere(rel > 2e-5)
    print('  bad maps (b,c):', bad[:10].tolist(), 'of', rel.size)
    b, c = (bad[0] if len(bad) else (0, 0))
    d = np.abs(co[b, c] - z[b, c])
    print('  map', (int(b), int(c)), 'got energy', en[b, c], 'want', want[b, c])
    print('  coeff err by row u (max over v):', np.round(d.max(1) / np.abs(z[b, c]).max(), 4).tolist()[:16])
    print('  coeff err by col v (max over u):', np.round(d.max(0) / np.abs(z[b, c]).max(), 4).tolist()[:16])
    print('  got[0,:4]', co[b, c, 0, :4], 'want', z[b, c, 0, :4])
