"""Command line of the drop-in: same flags and defaults as the reference's
importance_generation.py:8-21, plus opt-in extras (--compress_rate to also emit the kept-channel
sets the prune_* scripts derive, --seed, --out_root).

    python importance_generation.py --net resnet_50 --dataset imagenet --batch_size 256 --limit 5
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 importance_generation.py --net resnet_50 ...
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

from . import dist as ddist
from .generate import imp_score
from .sites import score_dir
from .zoo import NETS, get_network


def build_parser():
    p = argparse.ArgumentParser(description='DCT importance-score generation (B200)')
    p.add_argument('--dataset', type=str, default='cifar10', choices=('cifar10', 'imagenet', 'DUTS'), help='dataset')
    p.add_argument('--data_dir', type=str, default='./data', help='dataset path (unused: inputs are seeded synthetic)')
    p.add_argument('--batch_size', type=int, default=128, help='Batch size for scoring.')
    p.add_argument('--pretrain_dir', type=str, default='checkpoints/googlenet.pt', help='load the model from the specified checkpoint')
    p.add_argument('--limit', type=int, default=5, help='The num of batch to get importance score.')
    p.add_argument('--net', type=str, default='googlenet', choices=tuple(NETS), help='net type')
    # opt-in extras
    p.add_argument('--compress_rate', type=str, default=None, help="e.g. '[0.]+[0.18]*29': also write kept_channels.json")
    p.add_argument('--save_pruned', type=str, default=None,
                   help='with --compress_rate: also build the pruned net, fill it from the scored one on the device '
                        '(what load_model does before fine-tuning) and torch.save its state dict here')
    p.add_argument('--seed', type=int, default=0, help='seed of the random-init weights when no checkpoint is found')
    p.add_argument('--out_root', type=str, default='importance_score')
    p.add_argument('--input_side', type=int, default=None, help='override the input resolution (e.g. 288 for DUTS crops)')
    return p


def load_checkpoint(net, args):
    """Checkpoint handling of importance_generation.py:24-56; random init when the file is absent."""
    if not os.path.isfile(args.pretrain_dir):
        print('checkpoint %r not found: scoring seeded random-init weights (seed %d)' % (args.pretrain_dir, args.seed))
        return False
    ckpt = torch.load(args.pretrain_dir, map_location='cpu')
    state = ckpt.get('state_dict', ckpt) if isinstance(ckpt, dict) else ckpt
    if args.net in ('densenet_40', 'resnet_110'):
        state = {k.replace('module.', ''): v for k, v in state.items()}
    net.load_state_dict(state)
    return True


def main(argv=None):
    args = build_parser().parse_args(argv)
    if not torch.cuda.is_available():
        sys.exit('importance generation needs a CUDA device (sm_100a); there is no CPU fallback')
    rank, local_rank, world = ddist.init_from_env()
    device = torch.device('cuda', local_rank)
    torch.backends.cudnn.allow_tf32 = False           # activations stay fp32-exact like the reference's
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(args.seed)
    net = get_network(args.net)
    load_checkpoint(net, args)
    net = net.to(device).eval()
    files = imp_score(net, args, out_root=args.out_root)
    if args.compress_rate and rank == 0:
        from .topk import kept_channels
        kept = kept_channels(args.net, args.compress_rate, files, device=device)
        path = os.path.join(score_dir(args.net, args.limit, args.out_root), 'kept_channels.json')
        with open(path, 'w') as f:
            json.dump({'net': args.net, 'compress_rate': args.compress_rate,
                       'selections': [{'file': s.stem, 'conv': s.conv, 'C': s.C, 'k': s.k,
                                       'select_index': [int(i) for i in idx]} for s, idx in kept]}, f)
        print('kept-channel sets ->', path)
        if args.save_pruned:
            from .compress import get_compress_rate
            from .transfer import transfer_weights
            torch.manual_seed(args.seed)
            pruned = get_network(args.net, get_compress_rate(args.compress_rate)).to(device).eval()
            transfer_weights(args.net, pruned, net.state_dict(), kept)
            torch.save({k: v.cpu() for k, v in pruned.state_dict().items()}, args.save_pruned)
            print('pruned %s (%d parameters, was %d) ->' % (args.net, sum(p.numel() for p in pruned.parameters()),
                                                           sum(p.numel() for p in net.parameters())), args.save_pruned)
    if rank == 0:
        print('The importance score of %s has generated completed!' % args.net)
    ddist.shutdown()


if __name__ == '__main__':
    main()
