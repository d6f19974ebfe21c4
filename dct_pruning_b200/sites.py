"""Hook-site enumeration: which tensor of which module is scored, and which
``importance_score/<net>_limit<N>/*.npy`` file(s) each score vector lands in.

Mirrors the per-network branches of ``imp_score`` in
/root/reference/utils/common.py:384-977 (vgg :384-397, resnet_56 :400-437,
densenet_40 :440-476, googlenet :479-517, resnet_110 :520-554, resnet_50 :557-607,
u2netp :610-977), including the load-bearing quirks:
VGG ``relucfg`` ids that land on MaxPool modules, ResNet-50's duplicated block-0
``relu3`` file, GoogLeNet's ``imp_conv1_.npy`` and ``filters_p`` slicing,
DenseNet's last-12-channels window, U^2-Net's attribute-path file names.

The reference runs one forward sweep per site; here all sites are live in a
single sweep (one forward per batch), which is result-identical on fixed inputs.
"""
from dataclasses import dataclass
from typing import Optional, Tuple

VARIANT_OUTPUT = 'O'      # get_feature_hook: all channels of the module output      (common.py:262)
VARIANT_LAST12 = 'D'      # get_feature_hook_densenet: last 12 channels of the output (common.py:280)
VARIANT_INPUT = 'I'       # get_feature_hook_u2net_input: all channels of input[0]    (common.py:296)

DENSENET_WINDOW = 12      # hard-coded in the reference (common.py:285)


@dataclass(frozen=True)
class ScoreFile:
    stem: str                     # file name without '.npy'
    lo: Optional[int] = None      # channel slice of the site's score vector (None = all)
    hi: Optional[int] = None


@dataclass(frozen=True)
class Site:
    module: str                   # attribute path below the net ('features.6', 'layer1.0.relu1')
    variant: str
    files: Tuple[ScoreFile, ...]


def _whole(stem):
    return (ScoreFile(stem),)


def _resnet_cifar(blocks_per_stage):
    sites = [Site('relu', VARIANT_OUTPUT, _whole('imp_conv1'))]
    cnt = 1
    for stage in range(3):
        for j in range(blocks_per_stage):
            for relu in ('relu1', 'relu2'):
                cnt += 1
                sites.append(Site('layer%d.%d.%s' % (stage + 1, j, relu), VARIANT_OUTPUT,
                                  _whole('imp_conv%d' % cnt)))
    return sites


def _vgg(net):
    return [Site('features.%d' % cov_id, VARIANT_OUTPUT, _whole('imp_conv%d' % (i + 1)))
            for i, cov_id in enumerate(net.relucfg)]


def _densenet():
    sites = []
    for i in range(3):
        for j in range(12):
            sites.append(Site('dense%d.%d.relu' % (i + 1, j),
                              VARIANT_OUTPUT if j == 0 else VARIANT_LAST12,
                              _whole('imp_conv%d' % (13 * i + j + 1))))
        if i < 2:
            sites.append(Site('trans%d.relu' % (i + 1), VARIANT_LAST12,
                              _whole('imp_conv%d' % (13 * (i + 1)))))
    sites.append(Site('relu', VARIANT_LAST12, _whole('imp_conv39')))
    return sites


GOOGLENET_SITES = ['pre_layers', 'inception_a3', 'maxpool1', 'inception_a4', 'inception_b4',
                   'inception_c4', 'inception_d4', 'maxpool2', 'inception_a5', 'inception_b5']
GOOGLENET_BRANCHES = ['n1x1', 'n3x3', 'n5x5', 'pool_planes']


def _googlenet(net):
    sites = [Site(GOOGLENET_SITES[0], VARIANT_OUTPUT, _whole('imp_conv1_'))]
    for idx in range(1, len(GOOGLENET_SITES)):
        widths = net.filters_p[idx - 1]
        files, lo = [], 0
        for tp, w in zip(GOOGLENET_BRANCHES, widths):
            files.append(ScoreFile('imp_conv%d_%s' % (idx + 1, tp), lo, lo + w))
            lo += w
        sites.append(Site(GOOGLENET_SITES[idx], VARIANT_OUTPUT, tuple(files)))
    return sites


def _resnet50(net):
    sites = [Site('maxpool', VARIANT_OUTPUT, _whole('imp_conv1'))]
    cnt = 1
    for stage, repeat in enumerate(net.num_blocks):
        for j in range(repeat):
            base = 'layer%d.%d.' % (stage + 1, j)
            for relu in ('relu1', 'relu2'):
                cnt += 1
                sites.append(Site(base + relu, VARIANT_OUTPUT, _whole('imp_conv%d' % cnt)))
            stems = []
            if j == 0:                      # shortcut conv gets its own copy of the same vector
                cnt += 1
                stems.append('imp_conv%d' % cnt)
            cnt += 1
            stems.append('imp_conv%d' % cnt)
            sites.append(Site(base + 'relu3', VARIANT_OUTPUT, tuple(ScoreFile(s) for s in stems)))
    return sites


U2NETP_STAGE_DEPTH = {'stage1': 7, 'stage2': 6, 'stage3': 5, 'stage4': 4, 'stage5': 4, 'stage6': 4,
                      'stage5d': 4, 'stage4d': 4, 'stage3d': 5, 'stage2d': 6, 'stage1d': 7}


def _u2netp():
    sites = []
    for stage, depth in U2NETP_STAGE_DEPTH.items():
        names = (['rebnconvin'] + ['rebnconv%d' % n for n in range(1, depth + 1)]
                 + ['rebnconv%dd' % n for n in range(depth - 1, 0, -1)])
        for name in names:
            path = '%s.%s.relu_s1' % (stage, name)
            sites.append(Site(path, VARIANT_OUTPUT, _whole('net.' + path)))
    for k in range(1, 7):
        sites.append(Site('side%d' % k, VARIANT_INPUT, _whole('net.side%d' % k)))
    return sites


def hook_sites(net_name, net):
    """Sites for `net_name`; `net` supplies the few attributes the reference reads off the model
    (``relucfg``, ``filters_p``, ``num_blocks``)."""
    if net_name == 'vgg_16_bn':
        return _vgg(net)
    if net_name == 'resnet_56':
        return _resnet_cifar(9)
    if net_name == 'resnet_110':
        return _resnet_cifar(18)
    if net_name == 'densenet_40':
        return _densenet()
    if net_name == 'googlenet':
        return _googlenet(net)
    if net_name == 'resnet_50':
        return _resnet50(net)
    if net_name == 'u2netp':
        return _u2netp()
    raise ValueError('the network name you have entered is not supported yet: %r' % (net_name,))


def resolve_module(net, path):
    mod = net
    for part in path.split('.'):
        mod = mod[int(part)] if part.isdigit() else getattr(mod, part)
    return mod


def score_dir(net_name, limit, root='importance_score'):
    """``importance_score/<net>_limit<limit>`` (common.py:368-371)."""
    import os
    return os.path.join(root, '%s_limit%d' % (net_name, limit))
