"""Small launches of every kernel instantiation AUTO can pick (and the explicit paths), for compute-sanitizer:
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_cases.py
Each case is checked against Parseval so that a tool run is also a correctness run."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dct_pruning_b200 import _lib                      # noqa: E402
from dct_pruning_b200.ops import dct_energy, finalize  # noqa: E402
from dct_pruning_b200.topk import topk_segmented       # noqa: E402

dev = torch.device('cuda', 0)
lib = _lib.load()
_lib.check(lib.dctp_init())
g = torch.Generator().manual_seed(0)
seen = {}


def run(shape, path='auto', c_begin=0, c_count=None, strided=False):
    x = torch.relu(torch.randn(*shape, generator=g)).to(dev)
    if strided:                                             # batch-strided view (every other image of a larger tensor)
        big = torch.relu(torch.randn(shape[0] * 2, *shape[1:], generator=g)).to(dev)
        x = big[::2]
    acc, en, _ = dct_energy(x, c_begin=c_begin, c_count=c_count, path=path, want_energy=True)
    name = lib.dctp_last_kernel().decode().split(' (')[0]
    cc = x.shape[1] - c_begin if c_count is None else c_count
    want = (x[:, c_begin:c_begin + cc].double() ** 2).sum((2, 3))
    err = float(((en.double() - want).abs() / want.clamp_min(1e-30))[want > 0].max())
    acc2, _, _ = dct_energy(x, c_begin=c_begin, c_count=c_count, path=path)          # production path (no per-map energies)
    err2 = float(((acc2 - want.sum(0)).abs() / want.sum(0).clamp_min(1e-30)).max())
    assert err < 5e-5 and err2 < 5e-5, (shape, path, name, err, err2)
    seen[name] = seen.get(name, 0) + 1


for n in (1, 2, 3, 4, 5, 6, 7, 8):                          # Kronecker kernel: all eight instantiations, several tiles, tail tile
    run((3, 700, n, n))
for n in (10, 14, 16, 18, 20, 22, 28, 32, 40, 48, 56, 64):  # stacked-basis kernel: all six instantiations
    run((3, 150, n, n))
for n in (9, 13, 34, 50, 62):                               # sides it does not take: block-diagonal TMEM kernel / smem-operand kernel
    run((4, 37, n, n))
run((2, 9, 72, 72)); run((2, 5, 80, 80)); run((1, 3, 100, 100))          # 128-wide smem-operand kernel
run((2, 3, 96, 96)); run((1, 3, 160, 160)); run((1, 2, 320, 320))        # tiled large-map kernel
run((2, 48, 32, 32), c_begin=36, c_count=12)                # DenseNet window: per-map pointers
run((2, 30, 16, 16), c_begin=18, c_count=12)
run((3, 20, 8, 8), strided=True)
run((2, 3, 32, 16)); run((1, 2, 130, 130))                  # CUDA-core kernels
for p in ('umma', 'tmem', 'simt'):
    run((2, 21, 28, 28), path=p)
scores = torch.rand(5000, generator=g).to(dev)
kept = topk_segmented(scores, [0, 100, 1100, 5000], [10, 500, 3000])
assert [k.numel() for k in kept] == [10, 500, 3000]
out = finalize(torch.rand(77, dtype=torch.float64, generator=g).to(dev), 5)
w = torch.randn(16, 8, 9, generator=g).to(dev)
sel_o = torch.tensor([1, 3, 15], dtype=torch.int64, device=dev)
sel_i = torch.tensor([0, 7], dtype=torch.int64, device=dev)
o = torch.empty(3, 2, 9, device=dev)
_lib.check(lib.dctp_gather_weight(_lib.ptr(w), 16, 8, 9, _lib.ptr(sel_o), 3, _lib.ptr(sel_i), 2, _lib.ptr(o), _lib.current_stream()))
_lib.check(lib.dctp_check(None))
assert torch.equal(o, w[sel_o][:, sel_i])
torch.cuda.synchronize()
print('sanitize_cases ok:', json.dumps(seen) if False else seen)
