"""CPU-side checks of the C ABI: the library builds, loads, and exports every symbol include/tha4_b200.h declares.
No compute calls (there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    g.build_cuda()
    from tha4_b200 import _lib
    return _lib.load_library()


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'tha4_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(tha4_[a-z0-9_]+)\s*\(', text)))


def test_header_and_binding_agree():
    from tha4_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared_symbols()


def test_library_exports_every_declared_symbol(lib):
    for name in _declared_symbols():
        assert hasattr(lib, name), name


def test_base_grid_matches_oracle(lib, oracle_clib):
    # the one entry point that is pure host code: must be bit-identical to the oracle's restatement
    for size in (128, 192, 256, 512):
        a = (ctypes.c_float * size)()
        b = (ctypes.c_float * size)()
        assert lib.tha4_base_grid(size, a) == 0
        oracle_clib.tha4o_base_grid(size, b)
        assert list(a) == list(b)


def test_context_creation_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from tha4_b200._lib import Context, Tha4Error
    with pytest.raises(Tha4Error):
        Context(torch.device('cuda:0'))
    with pytest.raises(Tha4Error):
        Context(torch.device('cpu'))
