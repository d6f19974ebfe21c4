"""ResNet-50 for 224x224 inputs, hook-site compatible with the reference.

Scaffolding for the scoring path (the forward stays PyTorch/cuDNN); mirrors
the attribute names, channel arithmetic and parameter-creation order of
/root/reference/models/imagenet/resnet.py:3-151 so that
``net.layerK[j].relu{1,2,3}`` and ``net.maxpool`` resolve and seeded init
matches.
"""
import torch.nn as nn

STAGE_REPEAT = [3, 4, 6, 3]
STAGE_OUT = [64] + [256] * 3 + [512] * 4 + [1024] * 6 + [2048] * 3


def resnet50_channels(compress_rate):
    """(stem + 16 block output widths, 16 bottleneck mid widths); resnet.py:7-27."""
    out_rate = [compress_rate[0]]
    for i in range(len(STAGE_REPEAT) - 1):
        out_rate += [compress_rate[i + 1]] * STAGE_REPEAT[i]
    out_rate += [0.] * STAGE_REPEAT[-1]
    mid_rate = compress_rate[len(STAGE_REPEAT):]
    overall = [int(c * (1 - r)) for c, r in zip(STAGE_OUT, out_rate)]
    mid = [int(STAGE_OUT[i] // 4 * (1 - mid_rate[i - 1])) for i in range(1, len(STAGE_OUT))]
    return overall, mid


class Bottleneck(nn.Module):
    def __init__(self, midplanes, inplanes, planes, stride=1, is_downsample=False):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, midplanes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(midplanes)
        self.relu1 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(midplanes, midplanes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(midplanes)
        self.relu2 = nn.ReLU(inplace=True)
        self.conv3 = nn.Conv2d(midplanes, planes, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes)
        self.relu3 = nn.ReLU(inplace=True)
        self.stride, self.inplanes, self.planes, self.midplanes = stride, inplanes, planes, midplanes
        self.is_downsample, self.expansion = is_downsample, 4
        if is_downsample:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride, bias=False),
                                            nn.BatchNorm2d(planes))

    def forward(self, x):
        out = self.relu1(self.bn1(self.conv1(x)))
        out = self.relu2(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        out += self.downsample(x) if self.is_downsample else x
        return self.relu3(out)


class ResNet50(nn.Module):
    def __init__(self, compress_rate, num_classes=1000):
        super().__init__()
        overall, mid = resnet50_channels(compress_rate)
        self.num_blocks = STAGE_REPEAT
        self.conv1 = nn.Conv2d(3, overall[0], 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(overall[0])
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        at = 1
        for stage, repeat in enumerate(STAGE_REPEAT):
            blocks = nn.ModuleList()
            setattr(self, 'layer%d' % (stage + 1), blocks)
            for j in range(repeat):
                first = j == 0
                blocks.append(Bottleneck(mid[at - 1], overall[at - 1], overall[at],
                                         stride=(1 if stage == 0 else 2) if first else 1,
                                         is_downsample=first))
                at += 1
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(2048, num_classes)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        for stage in (self.layer1, self.layer2, self.layer3, self.layer4):
            for block in stage:
                x = block(x)
        return self.fc(self.avgpool(x).flatten(1))


def resnet_50(compress_rate):
    return ResNet50(compress_rate)
