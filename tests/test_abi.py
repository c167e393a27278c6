"""The C-ABI library: builds for sm_100a, loads, exports every symbol include/qgb200.h declares, and refuses to
run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope='module')
def lib():
    from pyqg_generative_b200 import build, _lib
    build.build()
    return _lib.load()


def header_functions():
    text = open(os.path.join(ROOT, 'include', 'qgb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(qgb_[a-z_0-9]+)\s*\(', text)))


def test_header_symbols_are_exported_and_bound(lib):
    from pyqg_generative_b200 import _lib
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), 'libqgb200.so does not export %s' % n
        assert n in _lib.SYMBOLS, 'python binding lacks %s' % n
    assert set(_lib.SYMBOLS) == set(names)


def test_config_struct_layout(lib):
    from pyqg_generative_b200 import _lib
    cfg = _lib.QgbConfig()
    lib.qgb_default_config(ctypes.byref(cfg))
    assert (cfg.nx, cfg.members, cfg.dt, cfg.rek, cfg.U1) == (64, 1, 7200.0, 5.787e-7, 0.025)
    assert ctypes.sizeof(cfg) == 16 + 10 * 8
    assert lib.qgb_version().startswith(b'qgb200')


def test_sass_is_sm100a(lib):
    import shutil
    import subprocess
    from pyqg_generative_b200 import _lib
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    out = subprocess.run([cuobjdump, '-lelf', _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


def test_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from pyqg_generative_b200 import _lib
    cfg = _lib.QgbConfig()
    lib.qgb_default_config(ctypes.byref(cfg))
    h = ctypes.c_void_p()
    rc = lib.qgb_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == _lib.QGB_ECUDA and not h
    assert b'no CPU fallback' in lib.qgb_last_error(None)
    with pytest.raises(RuntimeError):
        from pyqg_generative_b200.tools.stochastic_pyqg import EnsembleQGModel
        EnsembleQGModel(members=1)
