"""Time of one _invert (u, v, psi, spectral psi of every member: what a snapshot needs) and one q setter, per grid.
usage: python scripts/invert_time.py [nx=96] [members=512]      (QGB_GENERIC_STEP=1 selects the run-time-N interpreter at 32 / 48 / 96)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqg_generative_b200 import _lib
from pyqg_generative_b200.tools.stochastic_pyqg import EnsembleQGModel
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 96
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
m = EnsembleQGModel(nx=nx, dt=3600., members=B, log_level=0)
q0 = np.random.RandomState(0).randn(B, 2, nx, nx) * 1e-6
m.set_q(q0)
m._invert()
def timeit(f, n=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
inv = timeit(lambda: _lib.check(m._lib.qgb_invert(m._h, m._stream()), m._h))
print(json.dumps({'nx': nx, 'members': B, 'invert_ms': inv, 'generic': os.environ.get('QGB_GENERIC_STEP')}))
