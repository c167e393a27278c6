import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyqg_generative_b200.tools import operators as ops
from oracle import operators_ref as opr
x = np.random.RandomState(1).randn(2, 48, 48)
for n, N in ((48, 32), (48, 64), (48, 72), (48, 96), (48, 144)):
    try:
        y = ops.fft_interpolate(x, n, N)
        print(n, N, 'ok', np.abs(y - opr.fft_interpolate(x, n, N)).max())
    except Exception as e:
        print(n, N, 'FAIL', str(e)[:150])
