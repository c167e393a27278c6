"""Where does the step go: closure evaluation alone, spectral step alone (no closure), whole coupled step."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqg_generative_b200 import _lib, build

def main():
    nx, B, prec, reps = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 'tc', 20
    build.build()
    from oracle import cnn_ref
    from pyqg_generative_b200.models.cgan_regression import CGANRegression
    from pyqg_generative_b200.tools.cnn_tools import ChannelwiseScaler
    from pyqg_generative_b200.tools.stochastic_pyqg import stochastic_QGModel
    sd = cnn_ref.random_state_dict(4, 2, seed=0)
    gan = CGANRegression(folder='/nonexistent', nx=nx, precision=prec)
    gan.G.load_state_dict(sd)
    gan.x_scale, gan.y_scale = ChannelwiseScaler(), ChannelwiseScaler()
    gan.x_scale.std = np.array([7.8e-6, 1.05e-6], 'float32').reshape(1, 2, 1, 1)
    gan.y_scale.std = np.array([1e-11, 1e-12], 'float32').reshape(1, 2, 1, 1)
    m = stochastic_QGModel(dict(nx=nx, log_level=0, tmax=1e12, tavestart=1e12, members=B, parameterization=gan,
                                precision=prec, seed=1), 'constant', 1)
    m0 = stochastic_QGModel(dict(nx=nx, log_level=0, tmax=1e12, tavestart=1e12, members=B), 'constant', 1)
    rng = np.random.default_rng(0)
    q = rng.standard_normal((B, 2, nx, nx)) * 1e-6
    m.set_q(q); m0.set_q(q)
    def t(f):
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    lib = m._lib
    def closure():
        for _ in range(reps): _lib.check(lib.qgb_closure_eval(m._h, m._stream()), m._h)
    def step0():
        _lib.check(lib.qgb_step(m0._h, reps, m0._stream()), m0._h)
    def step():
        _lib.check(lib.qgb_step(m._h, reps, m._stream()), m._h)
    import time
    torch.cuda.synchronize()
    for name, f in (('closure_eval', closure), ('coupled step', step), ('closure_eval', closure)):
        t0 = time.perf_counter(); f(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print('%-20s host enqueue %.3f ms per call, until drained %.3f ms' % (name, (t1 - t0) * 1e3 / reps, (t2 - t0) * 1e3 / reps))
    for name, f in (('closure_eval only', closure), ('spectral step, no closure', step0), ('coupled step', step), ('closure_eval only', closure), ('coupled step', step)):
        print('%-28s %.3f ms' % (name, t(f)))

if __name__ == '__main__':
    main()
