"""Iteration time of the device-side CVAE (qgb_train_cvae_step) and CGAN (qgb_train_cgan_step) trainers at the shipped sizes
(AndrewCNN 4 -> 128 -> 64 -> 32 x 5 -> 2, encoder 4 -> ... -> 4, DCGAN discriminator ndf = 64; 64 images of 64 x 64) next to the
library baseline on the same GPU: torch eager (cuDNN, autograd incl. the double backward of the gradient penalty, torch.optim.Adam)
running the restated loops of oracle/train_ref.py -- what the reference's train_CVAE / train_CGAN run on a GPU.
The CGAN figure is the mean over 5 iterations, one of which updates the generator (i % 5 == 0).  One JSON line per configuration."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import cnn_ref, train_ref
from pyqg_generative_b200.models.cgan_regression import CGANRegression, CGANTrainer
from pyqg_generative_b200.models.cvae_regression import CVAERegression, CVAETrainer


def timed(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


def main():
    nx, B = 64, 64
    rng = np.random.RandomState(0)
    x = rng.randn(B, 2, nx, nx).astype('float32')
    y = rng.randn(B, 2, nx, nx).astype('float32')
    xd, yd = torch.as_tensor(x).cuda(), torch.as_tensor(y).cuda()
    # ---- CVAE
    vae = CVAERegression(folder='/nonexistent')
    vae.decoder.load_state_dict(cnn_ref.random_state_dict(4, 2, seed=0))
    enc_sd = cnn_ref.random_state_dict(4, 4, seed=1)
    vae.encoder.load_state_dict(enc_sd)
    tr = CVAETrainer(vae, nx, nx, max_batch=B)
    ours = timed(lambda: tr.step(xd, yd, 2e-4), 10)
    launches = (tr.enc.launch_count() + tr.dec.launch_count()) / 13
    tr.close()
    out = {'trainer': 'cvae', 'nx': nx, 'batch': B, 'ours_ms_per_step': ours * 1e3, 'kernels_per_step': launches}
    for name, tf32 in (('torch_tf32', True), ('torch_fp32', False)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        enc = train_ref.Net({k: v.numpy() for k, v in enc_sd.items()}).cuda().train()
        dec = train_ref.Net({k: v.numpy() for k, v in cnn_ref.random_state_dict(4, 2, seed=0).items()}).cuda().train()
        opt = torch.optim.Adam(list(enc.parameters()) + list(dec.parameters()), lr=2e-4)

        def step():
            opt.zero_grad()
            l = train_ref.cvae_losses(enc, dec, xd, yd, torch.randn_like(xd))
            l['loss'].backward()
            opt.step()
            return l['loss'].item()
        out[name + '_ms_per_step'] = timed(step, 10) * 1e3
    print(json.dumps(out))
    # ---- CGAN
    gan = CGANRegression(folder='/nonexistent', nx=nx)
    g_sd = cnn_ref.random_state_dict(4, 2, seed=0)
    gan.G.load_state_dict(g_sd)
    d_sd = {k: v.clone() for k, v in gan.D.state_dict().items()}
    tr = CGANTrainer(gan, nx, nx, max_batch=B)
    it = [0]

    def ours_step():
        tr.step(xd, yd, 2e-4, 2e-4, it[0] % 5 == 0)
        it[0] += 1
    ours = timed(ours_step, 10, warm=5)
    launches = (tr.G.launch_count() + tr.D.launch_count()) / 15
    tr.close()
    out = {'trainer': 'cgan', 'nx': nx, 'batch': B, 'ours_ms_per_iteration': ours * 1e3, 'kernels_per_iteration': launches}
    for name, tf32 in (('torch_tf32', True), ('torch_fp32', False)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        G = train_ref.Net({k: v.numpy() for k, v in g_sd.items()}).cuda().train()
        D = train_ref.Disc({k: v.numpy() for k, v in d_sd.items()}, nx).cuda().train()
        optD = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.5, 0.999))
        optG = torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.5, 0.999))
        it = [0]

        def step():
            z1, z2 = torch.randn_like(xd), torch.randn_like(xd)
            eps = torch.rand(B, 1, 1, 1, device='cuda')
            train_ref.cgan_iteration(G, D, optD, optG, xd, yd, z1, z2, eps, int(np.random.randint(0, 2)), it[0] % 5 == 0)
            it[0] += 1
        out[name + '_ms_per_iteration'] = timed(step, 10, warm=5) * 1e3
    print(json.dumps(out))


if __name__ == '__main__':
    main()
