"""CPU oracle arm of the long-run statistics check (profiles/r1_long_run.md): the Colab online-simulation setting
(nx=48 eddy, shipped CGAN generator, white latent noise every step) integrated by the oracle (pyqg shim + CPU torch
AndrewCNN) for a few independent members; writes the ensemble KE time series and time-mean KE spectrum.
usage: python scripts/long_run_oracle.py <members> <steps> <out.npz>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from conftest import golden_state_dict
from oracle import cnn_ref, operators_ref as opr, pyqg_shim

members, steps, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
sd, xs, ys = golden_state_dict('weights_gan.npz')
N, dt = 48, 14400.

class Par(pyqg_shim.QParameterization):
    def __init__(self, seed):
        self.rng = np.random.RandomState(seed)
    def __call__(self, m):
        z = self.rng.randn(1, 2, N, N).astype('float32')
        y = cnn_ref.predict_snapshot('gan', [sd], xs, ys, m.q, z)
        return cnn_ref.demean(y)

every = 50
ke = np.zeros((members, steps // every))
spec = np.zeros((2, N, N // 2 + 1)); nspec = 0
for b in range(members):
    np.random.seed(100 + b)
    m = pyqg_shim.QGModel(nx=N, dt=dt, log_level=0, tmax=1e12, tavestart=1e12, q_parameterization=Par(500 + b))
    opr.set_initial_condition(m)
    for s in range(steps):
        m._step_forward()
        if (s + 1) % every == 0:
            m._invert()
            ke[b, (s + 1) // every - 1] = m._calc_ke()
            if s + 1 > steps // 2:
                spec += m.wv2 * np.abs(m.ph) ** 2 / m.M ** 2; nspec += 1
    print('member', b, 'final KE %.3e' % ke[b, -1], flush=True)
np.savez(out, ke=ke, kespec=spec / max(nspec, 1), every=every, steps=steps)
