"""Diagnostic: tensor-core vs fp32 closure at large batch (multi-tile persistent loop) and health over steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cnn_ref
from pyqg_generative_b200.tools.cnn_tools import AndrewCNN
import bench

sd = cnn_ref.random_state_dict(4, 2, seed=0)
for B in (8, 64, 300):
    x = torch.randn(B, 4, 64, 64, generator=torch.Generator().manual_seed(B)).cuda()
    n32 = AndrewCNN(4, 2, precision='fp32'); n32.load_state_dict(sd)
    ntc = AndrewCNN(4, 2, precision='tc'); ntc.load_state_dict(sd)
    y32 = n32(x).cpu().numpy(); ytc = ntc(x).cpu().numpy()
    err = np.abs(ytc - y32).reshape(B, -1).max(1) / np.abs(y32).max()
    print('B=%d  max rel err per image: max %.2e  median %.2e  nonfinite %d  worst images %s' % (
        B, err.max(), np.median(err), (~np.isfinite(ytc)).sum(), np.argsort(err)[-5:]))
    ytc2 = ntc(x).cpu().numpy()
    print('   deterministic across calls:', np.array_equal(ytc, ytc2))

from pyqg_generative_b200.models.cgan_regression import CGANRegression
from pyqg_generative_b200.tools.cnn_tools import ChannelwiseScaler
from pyqg_generative_b200.tools.stochastic_pyqg import stochastic_QGModel
for prec in ('fp32', 'tc'):
    gan = CGANRegression(folder='/nonexistent', nx=64, precision=prec)
    gan.G.load_state_dict(sd)
    gan.x_scale, gan.y_scale = ChannelwiseScaler(), ChannelwiseScaler()
    gan.x_scale.std = np.array(bench.X_STD, 'float32').reshape(1, 2, 1, 1)
    gan.y_scale.std = np.array(bench.Y_STD, 'float32').reshape(1, 2, 1, 1)
    B = 256
    m = stochastic_QGModel(dict(nx=64, dt=14400., log_level=0, tmax=1e12, tavestart=1e12, members=B,
                                parameterization=gan, precision=prec, seed=2024), 'constant', 1)
    m.set_q(bench.synthetic_states(B, 64, 1234))
    for s in range(0, 41, 5):
        ke, cfl, flags = m.diagnostics()
        f = m.PV_forcing if s else None
        print(prec, 'step', m.tc, 'ke mean %.3e max %.3e  cfl max %.3f  flagged %d  |forcing| max %s' % (
            np.nanmean(ke), np.nanmax(ke), np.nanmax(cfl), (flags != 0).sum(), None if f is None else '%.2e' % np.abs(f).max()))
        m._step_forward(5)
