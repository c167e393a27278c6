"""A few training steps of the shipped 2 -> 128 -> ... -> 2 network (batch 64, 64^2) for ncu launch lists of the trainer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyqg_generative_b200.tools.cnn_tools import AndrewCNN, Trainer
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rng = np.random.RandomState(0)
x = rng.randn(batch, 2, nx, nx).astype('float32'); y = rng.randn(batch, 2, nx, nx).astype('float32')
tr = Trainer(AndrewCNN(2, 2), nx, nx, max_batch=batch)
for _ in range(steps):
    print(tr.step(x, y, 1e-3))
tr.close()
