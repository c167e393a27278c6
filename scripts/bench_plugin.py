"""Latency of the drop-in closure callable (pyqg ``parameterization(m) -> dq``, batch of ONE on the host like the reference's
own use): CGANRegression.__call__ with m.q (2, ny, nx) float64 numpy in, dq float64 numpy out."""
import os, sys, time, tempfile, pathlib, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from conftest import write_model_folder
from pyqg_generative_b200.models.cgan_regression import CGANRegression
from pyqg_generative_b200.tools.stochastic_pyqg import AR1_sampler

class M: pass
tmp = pathlib.Path(tempfile.mkdtemp())
for prec in ('fp32', 'tc'):
    for B, N in ((1, 48), (1, 64), (16, 64)):
        model = CGANRegression(folder=write_model_folder(tmp, 'gan'), nx=N, precision=prec)
        m = M(); m.ny = m.nx = N
        m.q = np.random.RandomState(0).randn(*(((B,) if B > 1 else ()) + (2, N, N))) * 1e-6
        m.sampling_type = 'AR1'; m.noise_sampler = AR1_sampler(1)
        for _ in range(5): y = model(m)
        t0 = time.perf_counter()
        for _ in range(50): y = model(m)
        dt = (time.perf_counter() - t0) / 50
        print(json.dumps({'precision': prec, 'batch': B, 'nx': N, 'ms_per_call': round(dt * 1e3, 3), 'out': list(np.shape(y))}))
