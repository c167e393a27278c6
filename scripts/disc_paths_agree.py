"""The discriminator's two GEMM paths (tensor-core split-TF32, FFMA under QGB_DISC_GEMM=ffma) on an odd configuration: nx = 48 (the
shipped models' grid: 3 x 3 last layer), batch 5.  Prints the gradients' relative difference per tensor; run once per path:
    python scripts/disc_paths_agree.py save /tmp/a.npz ; QGB_DISC_GEMM=ffma python scripts/disc_paths_agree.py cmp /tmp/a.npz"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqg_generative_b200.models.cgan_regression import CGANRegression, CGANTrainer
mode, path = sys.argv[1], sys.argv[2]
B, nx = 5, 48
rng = np.random.RandomState(3)
x = rng.randn(B, 2, nx, nx).astype('float32'); y = rng.randn(B, 2, nx, nx).astype('float32')
z1 = rng.randn(B, 2, nx, nx).astype('float32'); z2 = rng.randn(B, 2, nx, nx).astype('float32'); eps = rng.rand(B).astype('float32')
torch.manual_seed(9)                                  # (the constructor's weights_init draws the generator; D below)
net = CGANRegression(folder='/nonexistent', nx=nx)
net.D.load_state_dict({k: v * 2.5 for k, v in net.D.state_dict().items()})
tr = CGANTrainer(net, nx, nx, max_batch=8)
losses = tr.step(x, y, 0.0, 0.0, True, z1=z1, z2=z2, eps=eps, coin=0, update=False)
g = {('D/' + k): v for k, v in tr.D.last_grads().items()}
g.update({('G/' + k): v for k, v in tr.G.last_grads().items()})
print(os.environ.get('QGB_DISC_GEMM', 'tc'), losses)
if mode == 'save':
    np.savez(path, **g)
else:
    ref = np.load(path)
    worst = max(np.abs(g[k] - ref[k]).max() / np.abs(ref[k]).max() for k in g)
    print('worst relative difference between the two paths: %.2e' % worst)
