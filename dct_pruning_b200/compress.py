"""``--compress_rate`` mini-language and the kept-channel plan it implies.

* `get_compress_rate` mirrors /root/reference/utils/common.py:164-181: terms joined
  by '+', each holding exactly one decimal literal (the decimal point is mandatory)
  and at most one ``*N`` repeat.  Malformed input raises AssertionError, like the
  reference's asserts.
* `selection_plan` lists, for a net and a rate list, every top-k selection the
  reference's loaders perform (utils/load_models.py:17-772): which score file is
  read, its length C and the kept count k.  k is never recomputed from the rates in
  the loaders; it is read back as the pruned model's conv width
  (load_models.py:33,98,261,307,403,465,517,611), so the plan instantiates the
  original and the pruned net on the 'meta' device (shapes only) and walks the convs
  in the same order the loaders do.
"""
import re
from dataclasses import dataclass
from typing import List

import torch
import torch.nn as nn

_RATE = re.compile(r'\d+\.\d*')
_REPEAT = re.compile(r'\*\d+')


class CompressRateError(AssertionError, ValueError):
    pass


def get_compress_rate(spec):
    """'[0.]+[0.18]*29' -> [0.0, 0.18, ... x29].  Accepts the string or an args namespace."""
    text = spec if isinstance(spec, str) else spec.compress_rate
    rates = []
    for term in text.split('+'):
        repeat = 1
        found = _REPEAT.findall(term)
        if found:
            if len(found) != 1:
                raise CompressRateError('more than one repeat in term %r' % term)
            repeat = int(found[0].replace('*', ''))
        value = _RATE.findall(term)
        if len(value) != 1:
            raise CompressRateError('term %r must hold exactly one decimal literal' % term)
        rates += [float(value[0])] * repeat
    return rates


@dataclass(frozen=True)
class Selection:
    stem: str      # score file name without '.npy'
    C: int         # channels scored (length of the file)
    k: int         # channels kept
    conv: str      # state_dict name of the conv whose filters are gathered


def _shapes(net_name, rates, origin_rates=None):
    from .zoo import get_network
    with torch.device('meta'):
        orig = get_network(net_name, [0.] * 100 if origin_rates is None else list(origin_rates))
        pruned = get_network(net_name, list(rates))
    ow = {n: m.weight.shape for n, m in orig.named_modules() if isinstance(m, nn.Conv2d)}
    pw = {n: m.weight.shape for n, m in pruned.named_modules() if isinstance(m, nn.Conv2d)}
    return orig, pruned, ow, pw


def selection_plan(net_name, rates, origin_rates=None) -> List[Selection]:
    """The selections the reference's loader for `net_name` performs when it fills the net built with `rates` from the
    net built with `origin_rates` (None: the unpruned net; a list: an already pruned one, prune_dynamic.py:150-154)."""
    orig, pruned, ow, pw = _shapes(net_name, rates, origin_rates)
    plan = []

    def consider(conv, stem, even_if_full=False):
        C, k = ow[conv][0], pw[conv][0]
        if C != k or even_if_full:
            plan.append(Selection(stem, C, k, conv + '.weight'))
        return C != k

    if net_name in ('vgg_16_bn', 'densenet_40'):         # load_models.py:24-41, 394-409
        for cnt, conv in enumerate(ow, start=1):
            consider(conv, 'imp_conv%d' % cnt)
    elif net_name in ('resnet_56', 'resnet_110'):         # load_models.py:82-104
        cnt = 1
        for stage in range(3):
            for b in range(9 if net_name == 'resnet_56' else 18):
                for l in (1, 2):
                    cnt += 1
                    consider('layer%d.%d.conv%d' % (stage + 1, b, l), 'imp_conv%d' % cnt)
    elif net_name == 'resnet_50':                         # load_models.py:459-523
        consider('conv1', 'imp_conv1')
        cnt = 2
        for stage, repeat in enumerate(orig.num_blocks):
            for b in range(repeat):
                base = 'layer%d.%d.' % (stage + 1, b)
                order = ['conv1', 'conv2', 'downsample.0', 'conv3'] if b == 0 else ['conv1', 'conv2', 'conv3']
                for name in order:
                    consider(base + name, 'imp_conv%d' % cnt)
                    cnt += 1
    elif net_name == 'googlenet':                         # load_models.py:172-354
        cnt = 0
        for name, module in orig.named_modules():
            if name == 'pre_layers':
                cnt += 1
                consider('pre_layers.0', 'imp_conv%d' % cnt)
            elif name.startswith('inception_') and '.' not in name:
                cnt += 1
                consider(name + '.branch3x3.3', 'imp_conv%d_n3x3' % cnt)
                consider(name + '.branch5x5.3', 'imp_conv%d_n5x5' % cnt)
                consider(name + '.branch5x5.6', 'imp_conv%d_n5x5' % cnt)
    elif net_name == 'u2netp':                            # load_models.py:595-748
        # once one conv has been pruned (`last_select_index is not None`, load_models.py:647,689,725)
        # the loader argsorts every later stage conv too, even at k == C (it then keeps all channels)
        side, pruned_before = 0, False
        for conv in ow:
            if conv == 'outconv':
                break
            if conv.startswith('side'):
                side += 1
                consider(conv, 'net.side%d' % side)
            else:
                stage, block = conv.split('.')[:2]
                pruned_before |= consider(conv, 'net.%s.%s.relu_s1' % (stage, block), even_if_full=pruned_before)
    else:
        raise ValueError('the network name you have entered is not supported yet: %r' % (net_name,))
    return plan
