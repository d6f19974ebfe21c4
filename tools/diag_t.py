"""Find how the TMEM-operand kernel permutes coefficients when it is wrong (bring-up aid)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dct_pruning_b200.ops import dct_energy
from scipy.fft import dctn
dev = torch.device('cuda', 0)
for n in (int(v) for v in sys.argv[1:]):
    g = torch.Generator().manual_seed(n)
    x = torch.relu(torch.randn(1, 2, n, n, generator=g))
    _, en, co = dct_energy(x.to(dev), path='tmem', want_energy=True, want_coeff=True)
    got = co.cpu().numpy()[0, 0].astype(np.float64)
    want = dctn(x.numpy()[0, 0].astype(np.float64), type=2, norm='ortho')
    tol = 1e-3 * np.abs(want).max()
    rows = []
    for u in range(n):
        d = np.abs(want - got[u][None, :]).max(axis=1)
        m = int(np.argmin(d))
        rows.append(m if d[m] < tol else -1)
    cols = []
    for v in range(n):
        d = np.abs(want - got[:, v][:, None]).max(axis=0)
        m = int(np.argmin(d))
        cols.append(m if d[m] < tol else -1)
    print('N', n, 'got row u == want row:', rows)
    print('N', n, 'got col v == want col:', cols)
    # is got = P want for a linear map? least squares got = M @ want
    M = got @ np.linalg.pinv(want)
    print(' residual of got = M @ want:', np.abs(M @ want - got).max(), ' M diag head', np.round(np.diag(M)[:10], 3))
    nz = [(i, j, round(M[i, j], 3)) for i in range(min(n, 10)) for j in range(n) if abs(M[i, j]) > 0.05]
    print(' M nonzeros (rows<10):', nz[:40])
if os.environ.get('DCTP_T_DUMP_STAGE') == '1':
    for n in (int(v) for v in sys.argv[1:]):
        g = torch.Generator().manual_seed(n)
        x = torch.relu(torch.randn(1, 2, n, n, generator=g))
        _, _, co = dct_energy(x.to(dev), path='tmem', want_coeff=True)
        got = co.cpu().numpy()[0, 0].astype(np.float64)          # [v][h] = (C X^T)
        kk = np.arange(n)[:, None]; mm = np.arange(n)[None, :]
        C = np.cos(np.pi * (2 * mm + 1) * kk / (2 * n)) * np.sqrt(2.0 / n); C[0] *= np.sqrt(0.5)
        want = C @ x.numpy()[0, 0].astype(np.float64).T
        bad = np.abs(got - want).max(axis=0) > 1e-3 * np.abs(want).max()
        print('stage-1 dump N', n, 'bad h columns:', np.nonzero(bad)[0].tolist())
        for h in np.nonzero(bad)[0][:4]:
            d = np.abs(want - got[:, h][:, None]).max(axis=0)
            print('   got column h=%d equals want column %d (err %.2e)' % (h, int(np.argmin(d)), d.min()))
