import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu on the GPU box')
    config.addinivalue_line('markers', 'slow: CPU test that takes more than a few seconds')


def golden_scores(tag):
    z = np.load(os.path.join(GOLDEN, 'scores_%s.npz' % tag))
    meta = json.loads(str(z['__meta__']))
    return meta, {k: z[k] for k in z.files if k != '__meta__'}


def golden_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope='session')
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda', 0)


@pytest.fixture(scope='session')
def lib(cuda_device):
    import __graft_entry__ as entry
    from dct_pruning_b200 import _lib
    from dct_pruning_b200.build import is_stale
    if is_stale():
        entry.build()
    l = _lib.load()
    _lib.check(l.dctp_init())
    return l
