"""Per-layer time of the tensor-core closure (CUDA events around every launch of one layer, qgb_profile_begin/end).
usage: python scripts/layer_times.py [nx] [members] [precision] [reps]"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqg_generative_b200 import _lib, build

def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    prec = sys.argv[3] if len(sys.argv) > 3 else 'tc'
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
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
    rng = np.random.default_rng(0)
    m.set_q(rng.standard_normal((B, 2, nx, nx)) * 1e-6)
    lib, h, st = m._lib, m._h, m._stream()
    _lib.check(lib.qgb_step(h, 3, st), h)
    torch.cuda.synchronize()
    tot = 0.0
    for layer in range(8):
        _lib.check(lib.qgb_profile_begin(h, 0, layer), h)
        _lib.check(lib.qgb_step(h, reps, st), h)
        pms, pl, pim = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
        _lib.check(lib.qgb_profile_end(h, ctypes.byref(pms), ctypes.byref(pl), ctypes.byref(pim)), h)
        per = pms.value / max(pl.value, 1)
        tot += per
        print('layer %d: %.1f us per launch (%d launches, %d images)' % (layer + 1, per * 1e3, pl.value, pim.value // max(pl.value, 1)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _lib.check(lib.qgb_step(h, reps, st), h); e1.record(); torch.cuda.synchronize()
    print('sum of conv layers %.3f ms; whole step %.3f ms' % (tot, e0.elapsed_time(e1) / reps))

if __name__ == '__main__':
    main()
