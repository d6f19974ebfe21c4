"""The C-ABI boundary on a box without a GPU: the library builds, loads, exports every symbol
include/dctp.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, REPO


@pytest.fixture(scope='module')
def library():
    from dct_pruning_b200.build import LIB_PATH, build_library
    build_library()
    return ctypes.CDLL(LIB_PATH)


def declared_symbols():
    text = open(os.path.join(REPO, 'include', 'dctp.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(dctp_[a-z_0-9]+)\s*\(', text)))


def test_header_symbols_all_exported(library):
    names = declared_symbols()
    assert len(names) >= 12
    for name in names:
        assert hasattr(library, name), name


def test_binding_covers_header():
    from dct_pruning_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()


def test_version_and_path_selection(library):
    assert library.dctp_version() == 100
    library.dctp_path_for.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_longlong]
    assert library.dctp_path_for(56, 56, 56) == 1          # tensor-core path
    assert library.dctp_path_for(7, 7, 7) == 1
    assert library.dctp_path_for(80, 80, 80) == 4          # from side 80 (multiples of 16) the tiled kernel is faster
    assert library.dctp_path_for(72, 72, 72) == 1
    assert library.dctp_path_for(128, 128, 128) == 4
    assert library.dctp_path_for(100, 100, 100) == 1
    assert library.dctp_path_for(32, 16, 16) == 2          # non-square -> CUDA cores
    assert library.dctp_path_for(56, 56, 60) == 2          # strided rows -> CUDA cores
    assert library.dctp_path_for(320, 320, 320) == 4        # tiled tensor-core kernel for large maps
    assert library.dctp_path_for(130, 130, 130) == 2        # side not a multiple of 16 -> CUDA cores
    assert library.dctp_path_for(56, 56, 56) in (1, 3)


def test_no_cpu_fallback(library):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    library.dctp_last_error.restype = ctypes.c_char_p
    assert library.dctp_init() == -2                        # DCTP_E_CUDA
    assert library.dctp_last_error()
    x = np.ones((1, 1, 4, 4), np.float32)
    out = np.zeros(1, np.float32)
    rc = library.dctp_score_host(x.ctypes.data_as(ctypes.c_void_p), 1, 1, 4, 4, 0, 1,
                                 out.ctypes.data_as(ctypes.c_void_p), 0)
    assert rc == -2 and out[0] == 0.0


def test_python_path_refuses_cpu_tensors():
    import torch
    from dct_pruning_b200.hooks import ScoreSession
    from dct_pruning_b200.zoo import get_network
    net = get_network('vgg_16_bn').eval()
    with ScoreSession(net, 'vgg_16_bn'):
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            with torch.no_grad():
                net(torch.zeros(1, 3, 32, 32))


def test_npy_bytes_match_shipped_files(tmp_path):
    from dct_pruning_b200.generate import write_score_files
    for fn in ('shipped_googlenet_limit5_imp_conv2_n3x3.npy', 'shipped_googlenet_limit5_imp_conv1_.npy',
               'shipped_googlenet_limit1_imp_conv2_n5x5.npy'):
        raw = open(os.path.join(GOLDEN, fn), 'rb').read()
        arr = np.load(os.path.join(GOLDEN, fn))
        assert arr.dtype == np.float32 and len(raw) == 128 + 4 * arr.shape[0]
        write_score_files({'x': arr}, str(tmp_path))
        assert open(os.path.join(str(tmp_path), 'x.npy'), 'rb').read() == raw
    # the hand-written writer against np.save itself, also where the shape's digits push the header over a 64-byte boundary
    import io
    from dct_pruning_b200.generate import npy_bytes
    for n in (1, 9, 12, 64, 999, 1000, 22720, 10 ** 6):
        v = np.arange(n, dtype=np.float32) * 0.5
        ref = io.BytesIO()
        np.save(ref, v)
        assert npy_bytes(v) == ref.getvalue(), n
