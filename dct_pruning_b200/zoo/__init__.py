"""Network zoo the importance path hooks into (scaffolding; forward stays in PyTorch).

`get_network(name, compress_rate)` mirrors /root/reference/utils/common.py:31-54
minus the implicit `.cuda()`: device placement is the caller's decision.
"""
from .cifar import vgg_16_bn, resnet_56, resnet_110, densenet_40, googlenet
from .imagenet import resnet_50
from .duts import u2netp

NETS = {
    'vgg_16_bn': vgg_16_bn, 'resnet_56': resnet_56, 'resnet_110': resnet_110,
    'densenet_40': densenet_40, 'googlenet': googlenet, 'resnet_50': resnet_50,
    'u2netp': u2netp,
}

# input side and dataset each net is scored on (importance_generation.py:9, common.py:57-161)
NET_INPUT = {
    'vgg_16_bn': ('cifar10', 32), 'resnet_56': ('cifar10', 32), 'resnet_110': ('cifar10', 32),
    'densenet_40': ('cifar10', 32), 'googlenet': ('cifar10', 32),
    'resnet_50': ('imagenet', 224), 'u2netp': ('DUTS', 320),
}


def get_network(name, compress_rate=None):
    if name not in NETS:
        raise ValueError('the network name you have entered is not supported yet: %r' % (name,))
    return NETS[name](compress_rate=[0.] * 100 if compress_rate is None else compress_rate)
