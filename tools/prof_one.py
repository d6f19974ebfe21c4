"""One shape, one path, a few launches: the command ncu wraps.
    python tools/prof_one.py B C H W [path] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dct_pruning_b200 import _lib                      # noqa: E402
from dct_pruning_b200.ops import dct_energy            # noqa: E402

B, C, H, W = (int(v) for v in sys.argv[1:5])
path = sys.argv[5] if len(sys.argv) > 5 else 'auto'
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 5
dev = torch.device('cuda', 0)
lib = _lib.load()
_lib.check(lib.dctp_init())
x = torch.relu(torch.randn(B, C, H, W, device=dev))
# rotate through enough copies that no launch finds its input in the 126 MB L2 (a single re-used tensor below that size
# measures L2, not HBM - it made 7x7 look 35 % faster on one kernel than it is inside a real step)
copies = [x] + ([] if os.environ.get('PROF_NOROT') else [x.clone() for _ in range(max(0, int(400e6 // (x.numel() * 4))))])   # PROF_NOROT=1: deliberately L2-resident
acc = torch.zeros(C, dtype=torch.float64, device=dev)
for _ in range(2):
    dct_energy(x, path=path, accum=acc, check=False)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for i in range(iters):
    dct_energy(copies[i % len(copies)], path=path, accum=acc, check=False)
t1.record()
torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / iters
print('%s %s: %.3f ms, %.1f GB/s' % ((B, C, H, W), path, ms, x.numel() * 4 / ms / 1e6))
_lib.check(lib.dctp_check(None))
