"""Host-side cost of the hook plumbing on a launch-bound net: cProfile over forward passes with the hooks live.
    python tools/prof_hooks_host.py [net] [batch]"""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, '.')
from dct_pruning_b200.hooks import ScoreSession  # noqa: E402
from dct_pruning_b200.zoo import NET_INPUT, get_network  # noqa: E402

net_name = sys.argv[1] if len(sys.argv) > 1 else 'resnet_56'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device('cuda', 0)
torch.backends.cudnn.benchmark = True
net = get_network(net_name).to(dev).eval()
side = NET_INPUT[net_name][1]
x = torch.randn(B, 3, side, side, device=dev)
session = ScoreSession(net, net_name)
with torch.no_grad():
    for _ in range(5):
        net(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        net(x)
    torch.cuda.synchronize()
    bare = (time.perf_counter() - t0) / 50
    with session:
        for _ in range(5):
            net(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            net(x)
        torch.cuda.synchronize()
        hooked = (time.perf_counter() - t0) / 50
        print('%s batch %d: forward %.3f ms, with %d hooks %.3f ms (+%.1f us per hook), %d launches per pass' % (
            net_name, B, bare * 1e3, len(session.sites), hooked * 1e3, (hooked - bare) / len(session.sites) * 1e6, session.launches // 55))
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(20):
            net(x)
        torch.cuda.synchronize()
        pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
