"""Tensor-core closure against the fp32 path for awkward batch sizes and grid shapes (tile tails of the persistent kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cnn_ref
from pyqg_generative_b200.tools.cnn_tools import AndrewCNN
sd = cnn_ref.random_state_dict(4, 2, seed=3)
nets = {}
for prec in ('fp32', 'tc'):
    nets[prec] = AndrewCNN(4, 2, precision=prec); nets[prec].load_state_dict(sd)
worst = 0
for B, ny, nx in ((1, 64, 64), (7, 64, 64), (149, 64, 64), (1000, 64, 64), (1025, 32, 32), (3, 128, 64), (2, 64, 128), (5, 48, 96), (9, 96, 48), (2, 160, 160)):
    x = torch.randn(B, 4, ny, nx, generator=torch.Generator().manual_seed(B)).cuda()
    y32 = nets['fp32'](x).cpu().numpy(); ytc = nets['tc'](x).cpu().numpy()
    e = float(np.sqrt(((ytc - y32) ** 2).sum() / (y32 ** 2).sum()))
    worst = max(worst, e)
    print('B=%4d %3dx%3d  tc vs fp32 rel-L2 %.2e  finite %s' % (B, ny, nx, e, bool(np.isfinite(ytc).all())))
    assert e < 1e-3 and np.isfinite(ytc).all()
print('ok, worst %.2e' % worst)
