"""A few iterations of the CVAE / CGAN trainers at the shipped sizes (batch 64, 64^2) for ncu launch lists.
usage: adv_once.py cvae|cgan [iterations]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
which = sys.argv[1] if len(sys.argv) > 1 else 'cgan'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nx, B = 64, 64
rng = np.random.RandomState(0)
x = torch.as_tensor(rng.randn(B, 2, nx, nx).astype('float32')).cuda()
y = torch.as_tensor(rng.randn(B, 2, nx, nx).astype('float32')).cuda()
if which == 'cvae':
    from pyqg_generative_b200.models.cvae_regression import CVAERegression, CVAETrainer
    tr = CVAETrainer(CVAERegression(folder='/nonexistent'), nx, nx, max_batch=B)
    for _ in range(iters):
        print(tr.step(x, y, 2e-4))
else:
    from pyqg_generative_b200.models.cgan_regression import CGANRegression, CGANTrainer
    tr = CGANTrainer(CGANRegression(folder='/nonexistent', nx=nx), nx, nx, max_batch=B)
    for i in range(iters):
        print(tr.step(x, y, 2e-4, 2e-4, i == 0))
tr.close()
