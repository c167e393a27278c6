"""Training-step throughput of the device-side trainer (qgb_train_step) next to the library baseline on the same GPU: torch eager
(cuDNN convolutions, autograd, torch.optim.Adam) on the same network, batch and data -- the reference's own training path
(tools/cnn_tools.py:645-700) when it runs on a GPU.  Prints one JSON line per configuration."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import cnn_ref, train_ref                      # measurement script: the oracle module supplies the torch restatement
from pyqg_generative_b200.tools.cnn_tools import AndrewCNN, Trainer

HIDDEN = [128, 64, 32, 32, 32, 32, 32]
MAC = 267072                                               # MAC per pixel of the 2 -> ... -> 2 network


def main():
    for nx, batch in ((64, 64), (48, 64), (96, 32)):
        sd = cnn_ref.random_state_dict(2, 2, seed=0)
        rng = np.random.RandomState(0)
        x = rng.randn(batch, 2, nx, nx).astype('float32')
        y = rng.randn(batch, 2, nx, nx).astype('float32')
        net = AndrewCNN(2, 2, hidden_channels=HIDDEN)
        net.load_state_dict(sd)
        tr = Trainer(net, nx, nx, max_batch=batch)
        for _ in range(3):
            tr.step(x, y, 1e-3)
        torch.cuda.synchronize()
        n = 20
        t0 = time.perf_counter()
        for _ in range(n):
            loss = tr.step(x, y, 1e-3)
        torch.cuda.synchronize()
        ours = (time.perf_counter() - t0) / n
        launches = tr.launch_count()
        tr.close()
        out = {'nx': nx, 'batch': batch, 'ours_ms_per_step': ours * 1e3, 'ours_images_per_s': batch / ours, 'loss': loss,
               'ours_tflops_fwd_bwd': 3 * 2 * MAC * nx * nx * batch / ours / 1e12, 'kernels_per_step': launches / (n + 3)}
        for name, tf32 in (('torch_tf32', True), ('torch_fp32', False)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            ref = train_ref.Net({k: v.numpy() for k, v in sd.items()}).cuda().train()
            opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
            xd, yd = torch.as_tensor(x), torch.as_tensor(y)

            def step():
                opt.zero_grad()
                l = ref.compute_loss(xd.cuda(non_blocking=True), yd.cuda(non_blocking=True))['loss']
                l.backward()
                opt.step()
                return l.item()
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n):
                step()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n
            out[name + '_ms_per_step'] = dt * 1e3
            out[name + '_images_per_s'] = batch / dt
        print(json.dumps(out))


if __name__ == '__main__':
    main()
