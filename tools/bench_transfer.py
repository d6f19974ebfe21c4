"""Pruned-weight transfer: device gather vs the reference's Python copy loops.
    python tools/bench_transfer.py [net]        # default resnet_50, README compress rate, synthetic scores"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dct_pruning_b200 import transfer                                   # noqa: E402
from dct_pruning_b200.compress import get_compress_rate, selection_plan  # noqa: E402
from dct_pruning_b200.topk import kept_channels                          # noqa: E402
from dct_pruning_b200.zoo import get_network                             # noqa: E402

RATES = {'vgg_16_bn': '[0.50]*7+[0.95]*5', 'resnet_56': '[0.]+[0.18]*29',
         'resnet_110': '[0.]+[0.2]*2+[0.3]*18+[0.40]*18+[0.39]*19', 'resnet_50': '[0.]+[0.1]*3+[0.4]*7+[0.4]*9',
         'densenet_40': '[0.]+[0.2]*12+[0.]+[0.2]*12+[0.]+[0.2]*12', 'googlenet': '[0.4]+[0.85]*2+[0.9]*5+[0.9]*2',
         'u2netp': '[0.40]*40'}
net = sys.argv[1] if len(sys.argv) > 1 else 'resnet_50'
dev = torch.device('cuda', 0)
rates = get_compress_rate(RATES[net])
torch.manual_seed(0)
orig = get_network(net).eval()
pruned = get_network(net, rates).eval().to(dev)
rng = np.random.default_rng(0)
scores = {s.stem: rng.random(s.C).astype(np.float32) for s in selection_plan(net, rates)}
ori_dev = {k: v.to(dev) for k, v in orig.state_dict().items()}
for _ in range(2):
    kept = kept_channels(net, rates, scores, device=dev)
    plan = transfer.transfer_weights(net, pruned, ori_dev, kept)
torch.cuda.synchronize()
t0 = time.perf_counter()
kept = kept_channels(net, rates, scores, device=dev)
plan = transfer.transfer_weights(net, pruned, ori_dev, kept)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
moved = sum(pruned.state_dict()[op.name].numel() * 4 for op in plan if op.kind == 'gather')
print('%s: top-k + %d ops (%d gathers, %.1f MB gathered) on the device: %.1f ms wall' %
      (net, len(plan), sum(op.kind == 'gather' for op in plan), moved / 1e6, dt * 1e3))

# the reference's way, for one mid-size convolution: one tensor assignment per (kept out, kept in) pair
name = max((op.name for op in plan if op.kind == 'gather' and op.inp and op.out and orig.state_dict()[op.name].dim() == 4),
           key=lambda n: pruned.state_dict()[n].numel())
op = next(o for o in plan if o.name == name)
kd = {sel.stem: ids for sel, ids in kept}
w = orig.state_dict()[name]
dst = torch.empty(len(kd[op.out]), len(kd[op.inp]), *w.shape[2:])
t0 = time.perf_counter()
for index_i, i in enumerate(kd[op.out]):
    for index_j, j in enumerate(kd[op.inp]):
        dst[index_i][index_j] = w[i][j]
loop_s = time.perf_counter() - t0
pairs_total = sum(len(kd[o.out] if o.out else [0] * orig.state_dict()[o.name].shape[0]) * len(kd[o.inp])
                  for o in plan if o.kind == 'gather' and o.inp and orig.state_dict()[o.name].dim() == 4)
print('reference copy loops on the host: %s %s -> %.2f s (%.1f us per pair); all %d pairs of the net ~ %.0f s'
      % (name, tuple(dst.shape), loop_s, loop_s / (dst.shape[0] * dst.shape[1]) * 1e6, pairs_total,
         pairs_total * loop_s / (dst.shape[0] * dst.shape[1])))
