"""GPU arm of the long-run statistics check (profiles/r1_long_run.md): nx=48 eddy + shipped CGAN generator, white latent
noise every step, ensembles in precision fp32 and tc; compared with the CPU oracle ensemble of scripts/long_run_oracle.py.
usage: python scripts/long_run_gpu.py <members> <steps> <oracle.npz>"""
import os, sys, tempfile, pathlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from conftest import write_model_folder
from pyqg_generative_b200 import parallel
from pyqg_generative_b200.models.cgan_regression import CGANRegression
from pyqg_generative_b200.tools.simulate import set_initial_condition
from pyqg_generative_b200.tools.stochastic_pyqg import stochastic_QGModel

members, steps = int(sys.argv[1]), int(sys.argv[2])
ora = np.load(sys.argv[3]) if len(sys.argv) > 3 and os.path.exists(sys.argv[3]) else None
N, dt, every = 48, 14400., 50
tmp = pathlib.Path(tempfile.mkdtemp())
folder = write_model_folder(tmp, 'gan')
res = {}
for prec in ('fp32', 'tc'):
    model = CGANRegression(folder=folder, nx=N, precision=prec)
    m = stochastic_QGModel(dict(nx=N, dt=dt, log_level=0, tmax=1e12, tavestart=steps // 2 * dt, taveint=every * dt,
                                members=members, parameterization=model, precision=prec, seed=7), 'constant', 1)
    set_initial_condition(m, np.random.RandomState(3))
    ke = np.zeros((steps // every, members))
    for i in range(steps // every):
        m._step_forward(every)
        k, cfl, flags = m.diagnostics()
        assert not flags.any(), (prec, i, flags.sum())
        ke[i] = k
    d = m.averaged_diagnostics()
    res[prec] = (ke, d['KEspec'])
    print('%-4s ensemble-mean KE at steps %s: %s   (member std at the end %.2e)' % (
        prec, [every * (i + 1) for i in (9, 39, 79, steps // every - 1) if i < steps // every],
        ['%.3e' % ke[i].mean() for i in (9, 39, 79, steps // every - 1) if i < steps // every], ke[-1].std()))

def iso(spec):
    mm = stochastic_QGModel(dict(nx=N, dt=dt, log_level=0, members=1), 'constant', 1)
    return parallel.calc_ispec(mm.k, mm.l, spec)

half = steps // every // 2
for name, (ke, sp) in res.items():
    print('%-4s time-mean KE (second half) %.4e +- %.1e (standard error over members)' % (
        name, ke[half:].mean(), ke[half:].mean(axis=0).std() / np.sqrt(members)))
k32, s32 = iso(res['fp32'][1][0]); ktc, stc = iso(res['tc'][1][0])
print('upper-layer isotropic KE spectrum, tc vs fp32 ensembles: max relative difference over kappa %.3f' % np.abs(stc / s32 - 1).max())
if ora is not None:
    oke = ora['ke']
    print('oracle (CPU, %d members) time-mean KE (second half) %.4e +- %.1e' % (
        oke.shape[0], oke[:, half:].mean(), oke[:, half:].mean(axis=1).std() / np.sqrt(oke.shape[0])))
    ko, so = iso(ora['kespec'][0])
    print('upper-layer isotropic KE spectrum, fp32 ensemble vs oracle: relative difference per kappa', np.round(s32 / so - 1, 2))
    for i in (9, 19, 29, 39):
        print('  step %5d: KE oracle %.3e  fp32 %.3e  tc %.3e' % (every * (i + 1), oke[:, i].mean(), res['fp32'][0][i].mean(), res['tc'][0][i].mean()))
