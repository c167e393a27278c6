"""CGANRegression closure: pyqg_generative/models/cgan_regression.py on libqgb200.

Kept: constructor kwargs of ``model_args.json`` (regression, nx, generator, folder, div, hidden_channels), file
formats (G.pt, D.pt, x_scale.json, y_scale.json), and ``generate`` :133-137, ``generate_latent_noise`` :154-155,
``predict_snapshot`` :157-162, ``predict_mean_snapshot`` :164-171, ``predict`` :173-189, ``predict_ensemble`` :191-195.
Training: the discriminator (:57), ``fit`` :66-87, ``save_model`` :89-107 and ``train_CGAN`` :222-344 -- every iteration
of the WGAN-GP loop (two generator passes, four discriminator passes, gradient penalty with its exact second-order term,
Adam on D, and on every 5th iteration the generator update through D) is ONE call of ``qgb_train_cgan_step``.
"""
import ctypes
import os
from os.path import exists
from time import time

import numpy as np
import torch

from .. import _lib
from ..tools.cnn_tools import AndrewCNN, AverageLoss, DCGAN_discriminator, DiscState, Trainer, apply_function, extract, \
    minibatch, multistep_lr, prepare_PV_data, save_model_args, weights_init, write_log
from ._cnn_closure import CNNClosure, batched_mean_var, make_dataset
from .cvae_regression import _to_device, evaluate_prediction, loss_to_log

LAMBDA_DRIFT = 1e-3     # (applied inside qgb_train_cgan_step)
LAMBDA_GP = 10


class CGANRegression(CNNClosure):
    closure_kind = _lib.CLOSURE_GAN

    def __init__(self, regression='None', nx=64, generator='Andrew', folder='model', div=False,
                 hidden_channels=[128, 64, 32, 32, 32, 32, 32], precision='fp32'):
        self.folder = folder
        self.n_latent = 2
        self.regression, self.generator, self.nx, self.div = regression, generator, nx, div
        self.hidden_channels = hidden_channels
        if generator != 'Andrew':
            raise ValueError('generator not implemented')
        if regression != 'None':
            raise NotImplementedError("regression != 'None' (residual mean network) is not on the accelerated path")
        self.G = AndrewCNN(2 + self.n_latent, 2, div=div, hidden_channels=hidden_channels, precision=precision)
        self._D = None                      # built on first use: inference never needs the discriminator
        self.G.apply(weights_init)
        self.load_GAN(folder)

    @property
    def D(self):
        """DCGAN_discriminator(n_in + 2 n_out, bn='None', nx) (:57; 'minibatch discrimination': x and two forcings)."""
        if self._D is None:
            self._D = DCGAN_discriminator(2 + 2 * 2, bn='None', nx=self.nx)
            self._D.apply(weights_init)
            if exists('%s/D.pt' % self.folder):
                self._D.load_state_dict(torch.load('%s/D.pt' % self.folder, map_location='cpu'))
        return self._D

    @D.setter
    def D(self, value):
        self._D = value

    def _nets(self):
        return [self.G]

    def load_GAN(self, folder):
        if exists('%s/G.pt' % folder):
            self._load_state(self.G, '%s/G.pt' % folder)
            self._read_scales(folder)
            return True
        return False

    def fit(self, ds_train, ds_test, num_epochs=200, num_epochs_regression=50, batch_size=64, learning_rate=2e-4, nruns=5):
        """cgan_regression.py:66-87."""
        os.makedirs(self.folder, exist_ok=True)
        X_train, Y_train, X_test, Y_test, self.x_scale, self.y_scale = prepare_PV_data(ds_train, ds_test)
        self.save_model(*train_CGAN(self, ds_train, ds_test, X_train, Y_train, num_epochs, batch_size, learning_rate, nruns))

    def save_model(self, optim_loss, log_train, log_test):
        """cgan_regression.py:89-107: stats.nc, G.pt, D.pt, scalers, model_args.json."""
        os.makedirs(self.folder, exist_ok=True)
        stats, epoch = loss_to_log(optim_loss, log_train, log_test, name='loss')
        write_log(stats, '%s/stats.nc' % self.folder)
        print('Optimal epoch is ', epoch)
        print('The Last epoch is used for prediction')
        torch.save(self.G.state_dict(), '%s/G.pt' % self.folder)
        torch.save(self.D.state_dict(), '%s/D.pt' % self.folder)
        self.x_scale.write('x_scale.json', folder=self.folder)
        self.y_scale.write('y_scale.json', folder=self.folder)
        save_model_args('CGANRegression', folder=self.folder, regression=self.regression, nx=self.nx,
                        generator=self.generator, div=self.div, hidden_channels=self.hidden_channels)

    def generate(self, x, z=None):
        if z is None:
            z = torch.randn((x.shape[0], self.n_latent, x.shape[2], x.shape[3]), device=x.device)
        return self.G(torch.cat([x, z], dim=1))

    def generate_mean_var(self, x, M):
        return batched_mean_var(self.generate, x, M)

    def generate_ensemble(self, x, M):
        return torch.stack([self.generate(x) for _ in range(M)], dim=0)

    def generate_latent_noise(self, ny, nx):
        return np.random.randn(1, self.n_latent, ny, nx).astype('float32')

    def predict_snapshot(self, m, noise):
        X, single = self._normalized_q(m)
        noise = np.asarray(noise, dtype='float32').reshape(X.shape)
        Y = apply_function(self.G, X, noise, fun=self.generate)
        return self._denorm64(Y, single)

    def predict_mean_snapshot(self, m, M=100):
        X, single = self._normalized_q(m)
        acc = np.zeros_like(X)
        for _ in range(M):
            acc += apply_function(self.G, X, fun=self.generate)
        return self._denorm64(acc / M, single)

    def predict(self, ds, M=1000):
        X = self.x_scale.normalize(extract(ds, 'q').astype('float32'))
        Y, mean, var = apply_function(self.G, X, fun=self.generate_mean_var, M=M)
        shape = self._shape_of(ds)
        return make_dataset(
            q_forcing_advection=self.y_scale.denormalize(Y).reshape(shape),
            q_forcing_advection_mean=self.y_scale.denormalize(mean).reshape(shape),
            q_forcing_advection_var=self.y_scale.denormalize_var(var).reshape(shape))

    def predict_ensemble(self, ds, M=1000):
        X = self.x_scale.normalize(extract(ds, 'q').astype('float32'))
        chunks = [apply_function(self.G, X[i:i + 64], fun=self.generate_ensemble, M=M).reshape((M, -1) + X.shape[1:])
                  for i in range(0, len(X), 64)]
        Y = np.concatenate(chunks, axis=1)
        return self.y_scale.denormalize(Y).reshape((-1,) + self._shape_of(ds))


class CGANTrainer(object):
    """Device state of the generator / discriminator pair; ``step`` = one iteration of the loop at cgan_regression.py:256-292."""
    LOSS_KEYS = ('D_loss', 'D_grad', 'D_drift', 'G_loss')

    def __init__(self, net, ny, nx, max_batch=64, device=None):
        if ny != nx:
            raise ValueError('the discriminator expects square images')
        self.net = net
        self.G = Trainer(net.G, ny, nx, max_batch=max_batch, device=device)
        self.G.set_adam(0.5, 0.999)                                   # optim.Adam(..., betas=(0.5, 0.999)) :246-247
        self.D = DiscState(net.D, max_batch=max_batch, device=device)
        self._lib = _lib.load()
        self.g_loss = float('nan')

    def step(self, x, y, lr_d, lr_g, g_step, z1=None, z2=None, eps=None, coin=None, update=True):
        """x, y: (B, 2, nx, nx) float32.  z1, z2: latent noise of the two generator passes (default torch.randn on the device,
        ``generate`` :133-137); eps: (B,) uniform numbers and coin in {0, 1} of ``gradient_penalty`` :176-178 (default
        torch.rand / np.random.randint like the reference).  ``g_step``: also update the generator (every 5th iteration, :277).
        ``update=False``: gradients only (both networks), no optimizer step."""
        dev = torch.device('cuda:%d' % self.G.device)
        dv = lambda a: _to_device(a, dev)
        xd, yd = dv(x), dv(y)
        if xd.shape != yd.shape or xd.dim() != 4 or xd.shape[1] != 2:
            raise ValueError('expected x, y of shape (B, 2, ny, nx), got %s and %s' % (tuple(xd.shape), tuple(yd.shape)))
        B = xd.shape[0]
        z1d = torch.randn(xd.shape, device=dev) if z1 is None else dv(z1).reshape(xd.shape)
        z2d = torch.randn(xd.shape, device=dev) if z2 is None else dv(z2).reshape(xd.shape)
        e = np.ascontiguousarray(torch.rand(B).numpy() if eps is None else np.asarray(eps, dtype='float32').reshape(B))
        c = int(np.random.randint(0, 2, 1)[0]) if coin is None else int(coin)
        out = (ctypes.c_double * 4)()
        g_mode = (2 if update else 1) if g_step else 0
        _lib.check_train(self._lib.qgb_train_cgan_step(
            self.G._h, self.D._h, xd.data_ptr(), yd.data_ptr(), z1d.data_ptr(), z2d.data_ptr(), e.ctypes.data, c, B, 1,
            float(lr_d), float(lr_g), 1 if update else 0, g_mode, out, torch.cuda.current_stream(dev).cuda_stream), self.G._h)
        self.G.steps += 2                          # two training-mode generator passes per iteration (num_batches_tracked)
        if g_step:
            self.g_loss = float(out[3])
        return dict(D_loss=float(out[0]), D_grad=float(out[1]), D_drift=float(out[2]), G_loss=self.g_loss)

    def sync(self):
        self.G.sync_to(self.net.G)
        self.D.sync_to(self.net.D)

    def close(self):
        self.G.close()
        self.D.close()


def train_CGAN(net, ds_train, ds_test, X_train, Y_train, num_epochs, batch_size, learning_rate, nruns=5, evaluate=True,
               noise=None):
    """cgan_regression.py:222-344: WGAN-GP with drift penalty, Adam(betas 0.5, 0.999) + MultiStepLR(gamma 0.5) on both networks,
    the generator updated on every 5th minibatch of an epoch; epoch means of D_loss / D_grad / D_drift / G_loss in
    ``optim_loss`` and the offline scores of nruns random train / test runs after every epoch (``evaluate=False`` skips them).
    ``noise``: optional object with ``z(shape)``, ``eps(B)`` and ``coin()`` supplying the random draws (parity tests)."""
    X_train, Y_train = np.asarray(X_train), np.asarray(Y_train)
    print('Training starts on device %s, number of samples %d' % (torch.cuda.get_device_name(0), len(X_train)))
    tr = CGANTrainer(net, X_train.shape[2], X_train.shape[3], max_batch=batch_size)
    optim_loss, log_train, log_test = {}, [], []
    t_s = time()
    for epoch in range(num_epochs):
        t_e = time()
        lr = multistep_lr(learning_rate, num_epochs, epoch, gamma=0.5)
        logger = AverageLoss(optim_loss)
        for i, (x, y) in enumerate(minibatch(X_train, Y_train, batch_size=batch_size)):
            kw = {}
            if noise is not None:          # the reference's order of draws: z1, z2 (generate), eps, coin (gradient_penalty)
                kw = dict(z1=noise.z(tuple(x.shape)), z2=noise.z(tuple(x.shape)), eps=noise.eps(len(x)), coin=noise.coin())
            logger.accumulate(optim_loss, tr.step(x, y, lr, lr, i % 5 == 0, **kw), len(x))
        logger.average(optim_loss)
        tr.sync()
        if evaluate:
            log_train.append(evaluate_prediction(net, ds_train, nruns))
            log_test.append(evaluate_prediction(net, ds_test, nruns))
        t = time()
        msg = '[%d/%d] [%.2f/%.2f] D_loss: %.2f G_loss: %.2f' % (epoch + 1, num_epochs, t - t_e,
                                                                   (t - t_s) * (num_epochs / (epoch + 1) - 1),
                                                                   optim_loss['D_loss'][-1], optim_loss['G_loss'][-1])
        if evaluate:
            msg += ' L2_mean: [%.3f,%.3f] L2_total: [%.3f,%.3f] L2_res: [%.3f,%.3f]' % (
                log_train[-1]['L2_mean'], log_test[-1]['L2_mean'], log_train[-1]['L2_total'], log_test[-1]['L2_total'],
                log_train[-1]['L2_residual'], log_test[-1]['L2_residual'])
        print(msg)
    tr.close()
    return optim_loss, log_train, log_test
