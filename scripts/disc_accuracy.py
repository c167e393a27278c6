import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from oracle import train_ref
from pyqg_generative_b200.tools.cnn_tools import DCGAN_discriminator, weights_init
nx = 64
torch.manual_seed(9)
rng = np.random.RandomState(1)
xin = rng.randn(8, 6, nx, nx).astype('float32')
for ndf in (8, 64):
    D = DCGAN_discriminator(6, ndf=ndf, nx=nx); weights_init(D)
    sd = {k: v * 2.5 for k, v in D.state_dict().items()}
    D.load_state_dict(sd)
    ref = train_ref.Disc({k: v.numpy() for k, v in sd.items()}, nx).double()(torch.as_tensor(xin).double()).detach().numpy().reshape(-1)
    ref32 = train_ref.Disc({k: v.numpy() for k, v in sd.items()}, nx)(torch.as_tensor(xin)).detach().numpy().reshape(-1)
    out = D(torch.as_tensor(xin)).numpy().reshape(-1)
    print(os.environ.get('QGB_DISC_GEMM', 'tc'), 'ndf', ndf, 'ours vs f64 %.2e' % (np.abs(out - ref).max() / np.abs(ref).max()), ' torch fp32 vs f64 %.2e' % (np.abs(ref32 - ref).max() / np.abs(ref).max()))
