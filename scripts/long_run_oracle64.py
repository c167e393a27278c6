"""CPU oracle arm of the 10 000-step check at the benchmark grid (north_star: "matching KE time series and spectra over 10k
steps"): nx = 64 eddy configuration, dt = 4 h, no closure -- the setting of the recorded reference log
/root/reference/notebooks/3-2-dealiasing.ipynb:1431-1440 (KE 4.73e-4 at step 5000, 4.98e-4 at step 10 000, CFL 0.19) --
integrated by oracle/pyqg_shim.py for a few members whose initial conditions are drawn by the JAMES set_initial_condition
from ONE seeded stream, so that the GPU test can start from identical states (tests/test_gpu_longrun.py).

Writes tests/golden/long_run_oracle_64.npz: KE(t) of every member every 100 steps, the CFL numbers at the log cadence, the
full state of members 0 and 1 at step 2500 (still pre-chaotic: the trajectories are reproducible to ~1e-12 there), and the
time-mean KE / enstrophy spectra over steps 5000-10 000.
usage: python scripts/long_run_oracle64.py [members=8] [steps=10000]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import operators_ref as opr, pyqg_shim  # noqa: E402

members = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
N, dt, every, seed = 64, 14400., 100, 2024
rng = np.random.RandomState(seed)
ke = np.zeros((members, steps // every))
cfl = np.zeros((members, steps // 1000))
kespec = np.zeros((2, N, N // 2 + 1))
ensspec = np.zeros((2, N, N // 2 + 1))
nspec = 0
q_mid = []
q0_sum = []
for b in range(members):
    m = pyqg_shim.QGModel(nx=N, dt=dt, log_level=0, tmax=1e12, tavestart=1e12)
    opr.set_initial_condition(m, rng)
    q0_sum.append([m.q.sum(), np.abs(m.q).sum()])
    for s in range(1, steps + 1):
        m._step_forward()
        if s % every == 0:
            m._invert()
            ke[b, s // every - 1] = m._calc_ke()
            if s >= steps // 2:
                kespec += m.wv2 * np.abs(m.ph) ** 2 / m.M ** 2
                ensspec += np.abs(m.qh) ** 2 / m.M ** 2
                nspec += 1
        if s % 1000 == 0:
            cfl[b, s // 1000 - 1] = m._calc_cfl()
        if s == 2500 and b < 2:
            q_mid.append(m.q.copy())
    print('member', b, 'KE at 5k / 10k: %.3e %.3e  CFL %.3f' % (ke[b, steps // 2 // every - 1], ke[b, -1], cfl[b, -1]), flush=True)
np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'long_run_oracle_64.npz'), ke=ke, cfl=cfl, kespec=kespec / nspec,
                    ensspec=ensspec / nspec, q_2500=np.stack(q_mid), q0_sum=np.array(q0_sum), every=every, steps=steps,
                    seed=seed, members=members)
