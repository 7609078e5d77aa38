"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/*.h declares, and
compute entry points fail loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'tensorrec_b200.h')


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(trk_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for required in ('trk_csr_gather_reduce_f32', 'trk_csr_project_biases_f32', 'trk_score_f32', 'trk_rank_full',
                     'trk_score_topk_f16x3', 'trk_score_dense_f16x3', 'trk_topk_merge', 'trk_last_error'):
        assert required in names


def test_library_loads_and_exports_every_declared_symbol():
    from tensorrec_b200 import _lib
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), 'libtensorrec_b200.so does not export %s' % name
        assert name in _lib.SIGNATURES, 'ctypes binding misses %s' % name
    assert set(_lib.SIGNATURES) == set(declared_symbols())
    assert lib.trk_version() == 2000
    assert lib.trk_score_topk_max_k(128) >= 10 and lib.trk_score_topk_max_k(96) == 0
    assert lib.trk_rank_full_workspace_bytes(10, 100) == 0
    assert lib.trk_rank_full_workspace_bytes(3, 5000) == 2 * 3 * 2 * 4096 * 8     # two ping-pong key buffers


def test_argument_errors_are_reported_through_the_abi():
    from tensorrec_b200 import _lib
    lib = _lib.load()
    rc = lib.trk_csr_gather_reduce_f32(None, None, None, None, 4, 4, 4, 0, None, None, 0, None, None, None, None)
    assert rc == _lib.TRK_ERR_ARG
    assert 'null' in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, 'trk_csr_gather_reduce_f32')


def test_no_cpu_fallback_without_a_cuda_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    from tensorrec_b200 import kernels
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        kernels.require_cuda()
    import scipy.sparse as sp
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        kernels.DeviceCSR.from_scipy(sp.eye(3, format='csr', dtype=np.float32))
    # and the raw ABI reports a CUDA error instead of computing on the host
    from tensorrec_b200 import _lib
    lib = _lib.load()
    buf = (ctypes.c_float * 16)()
    idx = (ctypes.c_int32 * 16)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    q = ctypes.cast(idx, ctypes.c_void_p)
    rc = lib.trk_csr_project_biases_f32(q, q, p, p, 2, p, None)
    assert rc == _lib.TRK_ERR_CUDA


def test_integration_stub_argtypes_match_the_binding_table():
    """The ctypes stub INTEGRATION.md shows a maintainer of the reference must agree with the shipped table."""
    import ctypes
    import re
    from tensorrec_b200 import _lib
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'INTEGRATION.md')).read()
    env = {'_p': ctypes.c_void_p, '_i32': ctypes.c_int32, '_i64': ctypes.c_int64, 'ctypes': ctypes}
    found = re.findall(r'^_lib\.(trk_\w+)\.argtypes = (.+)$', text, flags=re.M)
    assert len(found) >= 10
    for name, expr in found:
        shown = eval(expr, env)
        shipped = _lib.SIGNATURES[name][1]
        assert len(shown) == len(shipped), name
        for a, b in zip(shown, shipped):
            assert ctypes.sizeof(a) == ctypes.sizeof(b) and (a is ctypes.c_void_p) == (b is ctypes.c_void_p), name
