"""GPU bring-up diagnostics: every shape class x path, errors reported instead of raised, plus a
first per-shape timing table.  Run on the GPU box:  python tools/gpu_diag.py > gpurun_out/diag.log"""
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dct_pruning_b200 import _lib                      # noqa: E402
from dct_pruning_b200.ops import dct_energy            # noqa: E402


def check(shape, path, dev):
    g = torch.Generator().manual_seed(shape[2])
    x = torch.relu(torch.randn(*shape, generator=g))
    xd = x.to(dev)
    try:
        _, en, co = dct_energy(xd, path=path, want_energy=True, want_coeff=True)
        want = (x.double() ** 2).sum(dim=(2, 3)).numpy()
        en = en.cpu().numpy()
        e_err = np.abs(en - want).max() / want.max()
        from scipy.fft import dctn
        z = dctn(x.numpy().astype(np.float64), type=2, norm='ortho', axes=(-2, -1))
        c_err = np.abs(co.cpu().numpy() - z).max() / np.abs(z).max()
        bad_maps = int((np.abs(en - want) > 1e-4 * want.max()).sum())
        print('%-18s %-5s energy_err %.2e coeff_err %.2e bad_maps %d/%d' % (shape, path, e_err, c_err, bad_maps, en.size), flush=True)
        if c_err > 1e-3:
            got = co.cpu().numpy()
            bm = np.argwhere(np.abs(got - z).max(axis=(2, 3)) > 1e-3 * np.abs(z).max())
            print('   first bad maps (b,c):', bm[:8].tolist(), flush=True)
            b, c = bm[0]
            d = np.abs(got[b, c] - z[b, c]) > 1e-3 * np.abs(z).max()
            print('   bad coeff rows:', np.nonzero(d.any(1))[0][:16].tolist(), 'cols:', np.nonzero(d.any(0))[0][:16].tolist(), flush=True)
            print('   got[:3,:3]', got[b, c][:3, :3].tolist(), 'want', z[b, c][:3, :3].tolist(), flush=True)
    except Exception as e:                                # noqa: BLE001
        print('%-18s %-5s FAILED: %s' % (shape, path, e), flush=True)
        traceback.print_exc()


def timing(shape, path, dev, iters=10):
    x = torch.relu(torch.randn(*shape, device=dev))
    acc = torch.zeros(shape[1], dtype=torch.float64, device=dev)
    for _ in range(3):
        dct_energy(x, path=path, accum=acc, check=False)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(iters):
        dct_energy(x, path=path, accum=acc, check=False)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / iters
    gb = x.numel() * 4 / 1e9
    print('time %-22s %-5s %8.3f ms  %8.1f GB/s  (%.2f of 6554.6)' % (shape, path, ms, gb / (ms / 1e3), gb / (ms / 1e3) / 6554.6), flush=True)


def main():
    dev = torch.device('cuda', 0)
    lib = _lib.load()
    _lib.check(lib.dctp_init())
    print('SMs', lib.dctp_sm_count(), torch.cuda.get_device_name(0), flush=True)
    for n in (8, 16, 32, 64, 7, 14, 28, 56, 4, 10, 20, 24, 40, 48, 9, 3, 1, 72, 80, 128):
        B, C = (3, 37) if n <= 64 else (2, 3)
        for path in ('simt', 'umma'):
            check((B, C, n, n), path, dev)
            rc = lib.dctp_check(None)
            if rc:
                print('   dctp_check ->', rc, lib.dctp_last_error(), flush=True)
    for shape in ((1, 2, 144, 144), (1, 2, 320, 320), (2, 3, 32, 16)):
        check(shape, 'auto', dev)
    if '--time' in sys.argv:
        for shape in ((256, 256, 56, 56), (256, 64, 56, 56), (256, 512, 28, 28), (256, 1024, 14, 14), (256, 2048, 7, 7),
                      (256, 64, 32, 32), (256, 64, 8, 8), (64, 64, 80, 80), (32, 64, 128, 128)):
            for path in ('umma', 'simt'):
                try:
                    timing(shape, path, dev)
                except Exception as e:                    # noqa: BLE001
                    print('time', shape, path, 'FAILED', e, flush=True)
        timing((12, 64, 160, 160), 'simt', dev, iters=3)
        timing((4, 64, 320, 320), 'simt', dev, iters=2)


if __name__ == '__main__':
    main()
