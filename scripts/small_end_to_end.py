"""Small end-to-end pass over every kernel family (seconds): coupled tc steps at 64^2 and 48^2 with the time-averaged budget
diagnostics, a 128^2 cluster-path step, an operator call.  (compute-sanitizer is not available on the GPU pool.)"""
import os, sys, tempfile, pathlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from conftest import write_model_folder
from pyqg_generative_b200.models.cgan_regression import CGANRegression
from pyqg_generative_b200.tools import operators as ops
from pyqg_generative_b200.tools.stochastic_pyqg import EnsembleQGModel, stochastic_QGModel
tmp = pathlib.Path(tempfile.mkdtemp())
rng = np.random.RandomState(0)
for N in (64, 48):
    model = CGANRegression(folder=write_model_folder(tmp, 'gan'), nx=N, precision='tc')
    m = stochastic_QGModel(dict(nx=N, log_level=0, tmax=1e12, tavestart=0., taveint=14400., dt=14400., members=3,
                                parameterization=model, precision='tc', seed=1), 'constant', 1)
    m.set_q(rng.randn(3, 2, N, N) * 1e-6)
    m._step_forward(3)
    d = m.averaged_diagnostics()
    assert np.isfinite(m.q).all() and np.isfinite(d['KEflux']).all()
m = EnsembleQGModel(nx=128, members=2, log_level=0, dt=7200., tavestart=0., taveint=7200.)
m.set_q(rng.randn(2, 2, 128, 128) * 1e-6)
m._step_forward(3)
assert np.isfinite(m.q).all() and np.isfinite(m.budget_sums()['APEflux']).all()
y = ops.Operator1(rng.randn(2, 128, 128), 48)
assert np.isfinite(y).all()
torch.cuda.synchronize()
print('small end-to-end pass done')
