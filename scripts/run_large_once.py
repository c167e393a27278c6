"""A few free-running steps of the 256^2 x 64 ensemble (configs[1]) for profiling the cluster kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pyqg_generative_b200 import _lib
from pyqg_generative_b200.tools.parameters import EDDY_PARAMS
from pyqg_generative_b200.tools.stochastic_pyqg import EnsembleQGModel
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
members = int(sys.argv[2]) if len(sys.argv) > 2 else 64
p = dict(EDDY_PARAMS.nx(nx)); p.update(members=members, log_level=0, tmax=1e12, tavestart=1e12)
m = EnsembleQGModel(**p)
m.set_q(bench.synthetic_states(members, nx, 7))
_lib.check(m._lib.qgb_step(m._h, 6, m._stream()), m._h)
torch.cuda.synchronize()
print('ok')
