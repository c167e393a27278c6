import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run on the B200 box with -m gpu)')


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def golden_state_dict(name):
    """(state_dict of torch tensors, x_std, y_std) from a weights_*.npz fixture."""
    import torch
    d = golden(name)
    sd = {k: torch.as_tensor(d[k]) for k in d.files if k not in ('x_std', 'y_std')}
    return sd, d['x_std'], d['y_std']


def write_model_folder(tmp_path, kind):
    """Materialise a reference-format model folder (``*.pt`` + json scalers) from the committed fixtures."""
    import json
    import torch
    files = {'gan': [('weights_gan.npz', 'G.pt')], 'vae': [('weights_vae.npz', 'decoder.pt')],
             'gz': [('weights_gz_mean.npz', 'net_mean.pt'), ('weights_gz_var.npz', 'net_var.pt')],
             'ols': [('weights_gz_mean.npz', 'net.pt')]}[kind]
    folder = str(tmp_path / kind)
    os.makedirs(folder, exist_ok=True)
    for npz, pt in files:
        sd, xs, ys = golden_state_dict(npz)
        torch.save(sd, os.path.join(folder, pt))
    for name, std in (('x_scale.json', xs), ('y_scale.json', ys)):
        s = str(np.asarray(std, 'float64').reshape(1, 2, 1, 1).tolist())
        with open(os.path.join(folder, name), 'w') as f:
            json.dump(dict(mean=str(np.zeros((1, 2, 1, 1)).tolist()), std=s), f)
    return folder


@pytest.fixture(scope='session')
def emu_lib():
    """Host emulation of the CUDA phase programs (tests/emu), built on demand with g++."""
    import ctypes
    so = os.path.join(ROOT, 'tests', 'emu', 'libqgb_emu.so')
    src = os.path.join(ROOT, 'tests', 'emu', 'emu.cpp')
    deps = [src] + [os.path.join(ROOT, 'pyqg_generative_b200', 'csrc', f) for f in ('qg_core.cuh', 'qg_host.hpp')]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(['g++', '-O2', '-fPIC', '-shared', '-std=c++17', '-o', so, src])
    return ctypes.CDLL(so)
