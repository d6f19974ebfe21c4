"""CIFAR-10 networks whose layers the importance path hooks.

These are *scaffolding*, not the product: the CNN forward pass stays in
PyTorch/cuDNN.  They exist because the scoring path addresses hook sites by
module attribute (``net.features[6]``, ``net.layer2[3].relu1``,
``net.inception_a4`` ...), so a drop-in needs nets whose attribute names,
channel arithmetic and parameter-creation order (hence seeded random init)
match the reference's constructors:

  vgg_16_bn    /root/reference/models/cifar10/vgg.py:9-56
  resnet_56/110 /root/reference/models/cifar10/resnet.py:4-170
  densenet_40  /root/reference/models/cifar10/densenet.py:13-125
  googlenet    /root/reference/models/cifar10/googlenet.py:8-223

``kept = int(C * (1 - rate))`` is evaluated in Python doubles exactly as the
reference does: the truncation of inexact products is what
defines k for the top-k selection.
"""
import math
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F


def _kept(c, rate):
    return int(c * (1 - rate))


# --------------------------------------------------------------------------- VGG
VGG_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512]
VGG_RELUCFG = [2, 6, 9, 13, 16, 19, 23, 26, 29, 33, 36, 39]


class VGG(nn.Module):
    def __init__(self, compress_rate, num_classes=10):
        super().__init__()
        self.relucfg = VGG_RELUCFG
        self.compress_rate = list(compress_rate) + [0.0]
        feats = nn.Sequential()
        cin, nconv = 3, 0
        for pos, width in enumerate(VGG_CFG):
            if width == 'M':
                feats.add_module('pool%d' % pos, nn.MaxPool2d(kernel_size=2, stride=2))
                continue
            cout = _kept(width, self.compress_rate[nconv])
            nconv += 1
            feats.add_module('conv%d' % pos, nn.Conv2d(cin, cout, kernel_size=3, padding=1))
            feats.add_module('norm%d' % pos, nn.BatchNorm2d(cout))
            feats.add_module('relu%d' % pos, nn.ReLU(inplace=True))
            cin = cout
        self.features = feats
        self.classifier = nn.Sequential(OrderedDict([
            ('linear1', nn.Linear(VGG_CFG[-2], VGG_CFG[-1])),
            ('norm1', nn.BatchNorm1d(VGG_CFG[-1])),
            ('relu1', nn.ReLU(inplace=True)),
            ('linear2', nn.Linear(VGG_CFG[-1], num_classes)),
        ]))

    def forward(self, x):
        x = F.avg_pool2d(self.features(x), 2)
        return self.classifier(x.flatten(1))


def vgg_16_bn(compress_rate):
    return VGG(compress_rate)


# ------------------------------------------------------------------- ResNet-56/110
def resnet_cifar_channels(compress_rate, depth):
    """(block output widths incl. stem, per-block mid widths); resnet.py:4-30."""
    n = (depth - 2) // 6
    widths = [16] + [16] * n + [32] * n + [64] * n
    out_rate = [compress_rate[0]] + [compress_rate[1]] * n + [compress_rate[2]] * n + [0.] * n
    mid_rate = compress_rate[3:]
    overall = [_kept(w, r) for w, r in zip(widths, out_rate)]
    mid = [_kept(widths[i], mid_rate[i - 1]) for i in range(1, len(widths))]
    return overall, mid


class _ChannelPadShortcut(nn.Module):
    """Option-A shortcut: stride-2 subsample (when strided) + zero channel pad."""

    def __init__(self, inplanes, planes, strided):
        super().__init__()
        gap = planes - inplanes
        self.pad = (0, 0, 0, 0, gap // 2, gap - gap // 2)
        self.strided = strided

    def forward(self, x):
        if self.strided:
            x = x[:, :, ::2, ::2]
        return F.pad(x, self.pad, "constant", 0)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, midplanes, inplanes, planes, stride=1):
        super().__init__()
        self.inplanes, self.planes, self.stride = inplanes, planes, stride
        self.conv1 = nn.Conv2d(inplanes, midplanes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(midplanes)
        self.relu1 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(midplanes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.relu2 = nn.ReLU(inplace=True)
        self.shortcut = nn.Sequential()
        if stride != 1 or inplanes != planes:
            self.shortcut = _ChannelPadShortcut(inplanes, planes, stride != 1)

    def forward(self, x):
        out = self.relu1(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        out += self.shortcut(x)
        return self.relu2(out)


class ResNetCifar(nn.Module):
    def __init__(self, depth, compress_rate, num_classes=10):
        super().__init__()
        assert (depth - 2) % 6 == 0, 'depth should be 6n+2'
        n = (depth - 2) // 6
        self.num_layer = depth
        self.overall_channel, self.mid_channel = resnet_cifar_channels(compress_rate, depth)
        self.conv1 = nn.Conv2d(3, self.overall_channel[0], 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(self.overall_channel[0])
        self.relu = nn.ReLU(inplace=True)
        self.layers = nn.ModuleList()
        at = 1
        for stage, stride in enumerate((1, 2, 2)):
            blocks = []
            for b in range(n):
                blocks.append(BasicBlock(self.mid_channel[at - 1], self.overall_channel[at - 1],
                                         self.overall_channel[at], stride if b == 0 else 1))
                at += 1
            setattr(self, 'layer%d' % (stage + 1), nn.Sequential(*blocks))
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        head = nn.Linear(64 * BasicBlock.expansion, num_classes)
        if depth == 56:
            self.fc = head
        else:
            self.linear = head

    def forward(self, x):
        x = self.relu(self.bn1(self.conv1(x)))
        x = self.layer3(self.layer2(self.layer1(x)))
        x = self.avgpool(x).flatten(1)
        return self.fc(x) if self.num_layer == 56 else self.linear(x)


def resnet_56(compress_rate):
    return ResNetCifar(56, compress_rate)


def resnet_110(compress_rate):
    return ResNetCifar(110, compress_rate)


# --------------------------------------------------------------------- DenseNet-40
class DenseBasicBlock(nn.Module):
    def __init__(self, inplanes, outplanes, dropRate=0):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.conv1 = nn.Conv2d(inplanes, outplanes, 3, padding=1, bias=False)
        self.dropRate = dropRate

    def forward(self, x):
        out = self.conv1(self.relu(self.bn1(x)))
        if self.dropRate > 0:
            out = F.dropout(out, p=self.dropRate, training=self.training)
        return torch.cat((x, out), 1)


class Transition(nn.Module):
    def __init__(self, inplanes, outplanes):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.conv1 = nn.Conv2d(inplanes, outplanes, 1, bias=False)

    def forward(self, x):
        return F.avg_pool2d(self.conv1(self.relu(self.bn1(x))), 2)


class DenseNet(nn.Module):
    def __init__(self, compress_rate, depth=40, num_classes=10, growthRate=12, compressionRate=1):
        super().__init__()
        assert (depth - 4) % 3 == 0, 'depth should be 3n+4'
        n = (depth - 4) // 3
        self.compress_rate = compress_rate
        self.covcfg = [3 * i + 1 for i in range(12 * 3 + 2 + 1)]
        self.growthRate, self.dropRate = growthRate, 0
        self.inplanes = growthRate * 2
        self.conv1 = nn.Conv2d(3, self.inplanes, 3, padding=1, bias=False)
        self.dense1 = self._dense(n, compress_rate[1:n + 1])
        self.trans1 = self._transition(compressionRate, compress_rate[n + 1])
        self.dense2 = self._dense(n, compress_rate[n + 2:2 * n + 2])
        self.trans2 = self._transition(compressionRate, compress_rate[2 * n + 2])
        self.dense3 = self._dense(n, compress_rate[2 * n + 3:3 * n + 3])
        self.bn = nn.BatchNorm2d(self.inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.avgpool = nn.AvgPool2d(8)
        self.fc = nn.Linear(self.inplanes, num_classes)
        # the reference re-draws every conv after construction, in modules() order
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / fan))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _dense(self, blocks, rates):
        seq = []
        for i in range(blocks):
            grow = _kept(self.growthRate, rates[i])
            seq.append(DenseBasicBlock(self.inplanes, grow, self.dropRate))
            self.inplanes += grow
        return nn.Sequential(*seq)

    def _transition(self, compressionRate, rate):
        cin = self.inplanes
        cout = int(math.floor(self.inplanes * (1 - rate) // compressionRate))
        self.inplanes = cout
        return Transition(cin, cout)

    def forward(self, x):
        x = self.dense1(self.conv1(x))
        x = self.dense2(self.trans1(x))
        x = self.dense3(self.trans2(x))
        x = self.avgpool(self.relu(self.bn(x)))
        return self.fc(x.flatten(1))


def densenet_40(compress_rate):
    return DenseNet(compress_rate, depth=40)


# ----------------------------------------------------------------------- GoogLeNet
GOOGLENET_FILTERS = [
    [64, 128, 32, 32], [128, 192, 96, 64], [192, 208, 48, 64],
    [160, 224, 64, 64], [128, 256, 64, 64], [112, 288, 64, 64],
    [256, 320, 128, 128], [256, 320, 128, 128], [384, 384, 128, 128],
]
GOOGLENET_MID = [[96, 16], [128, 32], [96, 16], [112, 24], [128, 24],
                 [144, 32], [160, 32], [160, 32], [192, 48]]
GOOGLENET_BLOCKS = ['a3', 'b3', 'a4', 'b4', 'c4', 'd4', 'e4', 'a5', 'b5']


def _conv_bn_relu(cin, cout, k, tag, pad=0):
    conv = nn.Conv2d(cin, cout, kernel_size=k, padding=pad)
    conv.tmp_name = tag
    return [conv, nn.BatchNorm2d(cout), nn.ReLU(True)]


class Inception(nn.Module):
    def __init__(self, in_planes, n1x1, n3x3red, n3x3, n5x5red, n5x5, pool_planes,
                 tmp_name, cp_rate, last=False):
        super().__init__()
        self.tmp_name = tmp_name
        self.n1x1, self.n3x3, self.n5x5, self.pool_planes = n1x1, n3x3, n5x5, pool_planes
        if n1x1:
            self.branch1x1 = nn.Sequential(*_conv_bn_relu(in_planes, n1x1, 1, tmp_name))
        if n3x3:
            out3 = n3x3 if last else int(n3x3 * cp_rate)
            self.branch3x3 = nn.Sequential(*(_conv_bn_relu(in_planes, n3x3red, 1, tmp_name)
                                             + _conv_bn_relu(n3x3red, out3, 3, tmp_name, 1)))
        if n5x5 > 0:
            out5 = n5x5 if last else int(n5x5 * cp_rate)
            inner = int(n5x5 * cp_rate)
            self.branch5x5 = nn.Sequential(*(_conv_bn_relu(in_planes, n5x5red, 1, tmp_name)
                                             + _conv_bn_relu(n5x5red, inner, 3, tmp_name, 1)
                                             + _conv_bn_relu(inner, out5, 3, tmp_name, 1)))
        if pool_planes > 0:
            self.branch_pool = nn.Sequential(nn.MaxPool2d(3, stride=1, padding=1),
                                             *_conv_bn_relu(in_planes, pool_planes, 1, tmp_name))

    def forward(self, x):
        return torch.cat([self.branch1x1(x), self.branch3x3(x),
                          self.branch5x5(x), self.branch_pool(x)], 1)


class GoogLeNet(nn.Module):
    def __init__(self, compress_rate, block=Inception):
        super().__init__()
        first = 192
        self.pre_layers = nn.Sequential(*_conv_bn_relu(3, first, 3, 'pre_layer', 1))
        self.filters = [row[:] for row in GOOGLENET_FILTERS]
        self.filters_p = [row[:] for row in GOOGLENET_FILTERS]
        for i, row in enumerate(self.filters_p):
            row[1] = _kept(row[1], compress_rate[i + 1])
            row[2] = _kept(row[2], compress_rate[i + 1])
        keep = [1 - r for r in compress_rate]
        f, m = self.filters, GOOGLENET_MID
        cin = [first] + [f[i][0] + int(f[i][1] * keep[i + 1]) + int(f[i][2] * keep[i + 1]) + f[i][3]
                         for i in range(8)]
        # reference quirk kept: inception_b3 is tagged 'a4' (googlenet.py:162)
        tags = ['a3', 'a4', 'a4', 'b4', 'c4', 'd4', 'e4', 'a5', 'b5']

        def make(i):
            return block(cin[i], f[i][0], m[i][0], f[i][1], m[i][1], f[i][2], f[i][3],
                         tags[i], keep[i + 1], **({'last': True} if i == 8 else {}))

        self.inception_a3 = make(0)
        self.inception_b3 = make(1)
        self.maxpool1 = nn.MaxPool2d(3, stride=2, padding=1)
        self.maxpool2 = nn.MaxPool2d(3, stride=2, padding=1)
        for i in range(2, 9):
            setattr(self, 'inception_' + GOOGLENET_BLOCKS[i], make(i))
        self.avgpool = nn.AvgPool2d(8, stride=1)
        self.linear = nn.Linear(sum(f[-1]), 10)

    def forward(self, x):
        x = self.inception_b3(self.inception_a3(self.pre_layers(x)))
        x = self.maxpool1(x)
        for tag in ('a4', 'b4', 'c4', 'd4', 'e4'):
            x = getattr(self, 'inception_' + tag)(x)
        x = self.maxpool2(x)
        x = self.inception_b5(self.inception_a5(x))
        return self.linear(self.avgpool(x).flatten(1))


def googlenet(compress_rate):
    return GoogLeNet(compress_rate)
