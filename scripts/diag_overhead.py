"""Cost of pyqg's time-averaged diagnostics (KEspec, Ensspec and the 13 budget terms sampled every ``taveint``) on top of the free-running
step: member-steps/s with the averaging switched off and on (the reference's runs average over the second half, one sample a day).
usage: python scripts/diag_overhead.py [nx=64] [members=1024] [steps=240]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqg_generative_b200 import _lib
from pyqg_generative_b200.tools.stochastic_pyqg import EnsembleQGModel

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 240
dt = {32: 14400., 48: 14400., 64: 14400., 96: 7200., 128: 7200., 256: 3600.}.get(nx, 3600.)
rng = np.random.RandomState(0)
q0 = rng.randn(B, 2, nx, nx) * np.array([7.8e-6, 1.05e-6])[None, :, None, None]
h = np.fft.rfftn(q0, axes=(-2, -1)); h[..., nx // 4:, :] = 0; h[..., :, nx // 4:] = 0
q0 = np.fft.irfftn(h, s=(nx, nx), axes=(-2, -1))
out = {'nx': nx, 'members': B, 'steps': steps, 'dt': dt}
for name, tave in (('off', 1e12), ('on', 0.0)):
    m = EnsembleQGModel(nx=nx, dt=dt, members=B, log_level=0, tmax=1e12, tavestart=tave, taveint=86400.)
    m.set_q(q0)
    m._step_forward(2 * int(86400. / dt))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(m._lib.qgb_step(m._h, steps, m._stream()), m._h)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out['ms_per_step_' + name] = ms
    out['member_steps_per_s_' + name] = B / (ms * 1e-3)
    del m
out['samples'] = steps * dt / 86400.
out['ms_per_sample'] = (out['ms_per_step_on'] - out['ms_per_step_off']) * steps / out['samples']
print(json.dumps(out))
