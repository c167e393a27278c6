"""configs[4]: coarse-graining of 256^2 snapshots (Operator1/2 and PV_subgrid_forcing to 64^2), snapshots per second."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqg_generative_b200.tools import operators as ops
from pyqg_generative_b200.tools.parameters import EDDY_PARAMS

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rng = np.random.default_rng(0)
q = torch.as_tensor(rng.standard_normal((B, 2, 256, 256)) * np.array([7e-6, 1e-6])[None, :, None, None]).cuda()
params = dict(EDDY_PARAMS.nx(256))
def timeit(f, reps=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
for name, f in (('Operator1 256->64', lambda: ops.Operator1(q, 64)), ('Operator2 256->64', lambda: ops.Operator2(q, 64)),
                ('PV_subgrid_forcing Operator1 none', lambda: ops.PV_subgrid_forcing(q, 64, ops.Operator1, params, dealias='none')),
                ('PV_subgrid_forcing Operator2 3/2-rule', lambda: ops.PV_subgrid_forcing(q, 64, ops.Operator2, params, dealias='3/2-rule'))):
    t = timeit(f)
    print(json.dumps({'op': name, 'snapshots': B, 'ms_per_call': round(t * 1e3, 3), 'snapshots_per_s': round(B / t, 1)}))
