"""Time the spectral step kernel alone (no closure): member-steps/s and fraction of the HBM roofline.
usage: python scripts/spectral_time.py [nx=64] [members=1024] [steps=200]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqg_generative_b200 import _lib
from pyqg_generative_b200.tools.stochastic_pyqg import EnsembleQGModel

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
dt = {32: 14400., 48: 14400., 64: 14400., 96: 7200., 128: 7200., 256: 3600.}.get(nx, 3600.)
m = EnsembleQGModel(nx=nx, dt=dt, members=B, log_level=0, tmax=1e12, tavestart=1e12)
rng = np.random.RandomState(0)
q0 = rng.randn(B, 2, nx, nx) * np.array([7.8e-6, 1.05e-6])[None, :, None, None]
h = np.fft.rfftn(q0, axes=(-2, -1)); h[..., nx // 4:, :] = 0; h[..., :, nx // 4:] = 0
m.set_q(np.fft.irfftn(h, s=(nx, nx), axes=(-2, -1)))
m._step_forward(5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
_lib.check(m._lib.qgb_step(m._h, steps, m._stream()), m._h)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
nk = nx // 2 + 1
byts = 5 * (2 * nx * nk * 16) + (2 * nx * nx * 8)         # bytes_dyn = 5 S_c + S_r
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
print(json.dumps({'nx': nx, 'members': B, 'ms_per_step': ms, 'member_steps_per_s': B / (ms * 1e-3),
                  'hbm_frac': byts * B / (ms * 1e-3) / 1e9 / peaks['hbm_gbs'], 'ke_finite': bool(np.isfinite(m.diagnostics()[0]).all()),
                  's64_off': os.environ.get('QGB_S64_OFF')}))
