"""Free-running (no closure) step rate of the large-grid (thread-block-cluster) path: configs[1] and neighbours."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pyqg_generative_b200 import _lib
from pyqg_generative_b200.tools.parameters import EDDY_PARAMS
from pyqg_generative_b200.tools.stochastic_pyqg import EnsembleQGModel

for nx, members in ((256, 64), (128, 64), (128, 256), (512, 16)):
    p = dict(EDDY_PARAMS.nx(nx))
    p.update(members=members, log_level=0, tmax=1e12, tavestart=1e12)
    m = EnsembleQGModel(**p)
    m.set_q(bench.synthetic_states(members, nx, 7))
    lib, h, st = m._lib, m._h, m._stream()
    _lib.check(lib.qgb_step(h, 5, st), h)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 20
    e0.record(); _lib.check(lib.qgb_step(h, steps, st), h); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ke = m.diagnostics()[0]
    print(json.dumps({'nx': nx, 'members': members, 'ms_per_step': round(ms, 4), 'member_steps_per_s': round(members / ms * 1e3, 1),
                      'healthy': bool(np.isfinite(ke).all())}))
    del m
