"""Time the alternative scoring ops (dctp_score_op) on ResNet-50-sized activations: maps/s of the Jacobi rank kernel and the
cost of dct3 next to dct2.  python tools/bench_alt_ops.py"""
import sys
import time

import torch

sys.path.insert(0, '.')
from dct_pruning_b200.ops import score_op  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    shapes = [(64, 256, 56, 56), (64, 512, 28, 28), (64, 1024, 14, 14), (64, 2048, 7, 7), (128, 64, 32, 32), (128, 64, 8, 8)]
    for shape in shapes:
        x = torch.relu(torch.randn(*shape, device=dev))
        for op in ('dct2', 'dct3', 'rank'):
            score_op(x, op)
            torch.cuda.synchronize()
            reps = 3
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                score_op(x, op, check=False)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            maps = shape[0] * shape[1]
            print('%-18s %-5s %9.3f ms  %8.2f M maps/s  %7.1f GB/s' % (shape, op, ms, maps / ms / 1e3, x.numel() * 4 / ms / 1e6), flush=True)
        # the CPU rule on the same slices (bounded sample): torch.linalg.matrix_rank per slice, as the reference's line would run it
        xs = x[:1, :64].cpu()
        t0 = time.time()
        for c in range(xs.shape[1]):
            torch.linalg.matrix_rank(xs[0, c])
        dt = time.time() - t0
        print('%-18s cpu matrix_rank per slice: %.1f us' % (shape, dt / xs.shape[1] * 1e6), flush=True)


if __name__ == '__main__':
    main()
