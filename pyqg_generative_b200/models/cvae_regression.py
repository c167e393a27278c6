"""CVAERegression closure: pyqg_generative/models/cvae_regression.py on libqgb200.

The decoder is the same AndrewCNN(4 -> 2) as the GAN generator (:45); ``generate`` :114-118,
``generate_latent_noise`` :128-129, ``predict_snapshot`` :131-136, ``predict_mean_snapshot`` :138-145,
``predict`` :147-163.  Training: the encoder AndrewCNN(4 -> 4) (:48), ``fit`` :53-70, ``save_model`` :72-91 and
``train_CVAE`` :250-320 -- every iteration (forward of both networks with batch statistics, ELBO of ``compute_loss``
:177-230, backward through decoder, reparameterisation and encoder, Adam on both) is ONE call of ``qgb_train_cvae_step``.
"""
import ctypes
import os
from os.path import exists
from time import time

import numpy as np
import torch

from .. import _lib
from ..tools.cnn_tools import AndrewCNN, AverageLoss, Trainer, apply_function, extract, minibatch, multistep_lr, \
    prepare_PV_data, save_model_args, write_log
from ..tools.computational_tools import subgrid_scores
from ._cnn_closure import CNNClosure, batched_mean_var, make_dataset

LOSS_KEYS = ('loss', 'loss_recon', 'loss_KL', 'MSE', 'var_latent', 'var_aggr')


class CVAERegression(CNNClosure):
    closure_kind = _lib.CLOSURE_VAE

    def __init__(self, regression='None', decoder_var='adaptive', folder='model', div=False,
                 hidden_channels=[128, 64, 32, 32, 32, 32, 32], precision='fp32'):
        self.folder = folder
        self.n_latent = 2
        self.regression, self.decoder_var, self.div = regression, decoder_var, div
        self.hidden_channels = hidden_channels
        if regression != 'None':
            raise NotImplementedError("regression != 'None' (residual mean network) is not on the accelerated path")
        self.decoder = AndrewCNN(2 + self.n_latent, 2, div=div, hidden_channels=hidden_channels, precision=precision)
        self.encoder = AndrewCNN(2 + 2, 2 * self.n_latent, precision=precision)        # (x, y) -> (mu, logvar), :48
        self.load_model(folder)

    def _nets(self):
        return [self.decoder]

    def load_model(self, folder):
        if exists('%s/decoder.pt' % folder):
            self._load_state(self.decoder, '%s/decoder.pt' % folder)
            if exists('%s/encoder.pt' % folder):
                self._load_state(self.encoder, '%s/encoder.pt' % folder)
            self._read_scales(folder)
            return True
        return False

    def fit(self, ds_train, ds_test, num_epochs=200, num_epochs_regression=50, batch_size=64, learning_rate=2e-4, nruns=5):
        """cvae_regression.py:53-70."""
        os.makedirs(self.folder, exist_ok=True)
        X_train, Y_train, X_test, Y_test, self.x_scale, self.y_scale = prepare_PV_data(ds_train, ds_test)
        self.save_model(*train_CVAE(self, ds_train, ds_test, X_train, Y_train, num_epochs, batch_size, learning_rate, nruns))

    def save_model(self, optim_loss, log_train, log_test):
        """cvae_regression.py:72-91: stats.nc (epoch means of the losses and the offline scores), encoder.pt, decoder.pt,
        scalers and model_args.json."""
        os.makedirs(self.folder, exist_ok=True)
        stats, epoch = loss_to_log(optim_loss, log_train, log_test)
        write_log(stats, '%s/stats.nc' % self.folder)
        print('Optimal epoch:', epoch)
        print('The Last epoch is used for prediction')
        torch.save(self.encoder.state_dict(), '%s/encoder.pt' % self.folder)
        torch.save(self.decoder.state_dict(), '%s/decoder.pt' % self.folder)
        self.x_scale.write('x_scale.json', folder=self.folder)
        self.y_scale.write('y_scale.json', folder=self.folder)
        save_model_args('CVAERegression', folder=self.folder, regression=self.regression, div=self.div,
                        decoder_var=self.decoder_var, hidden_channels=self.hidden_channels)

    def generate(self, x, z=None):
        if z is None:
            z = torch.randn((x.shape[0], self.n_latent, x.shape[2], x.shape[3]), device=x.device)
        return self.decoder(torch.cat([x, z], dim=1))

    def generate_mean_var(self, x, M):
        return batched_mean_var(self.generate, x, M)

    def generate_latent_noise(self, ny, nx):
        return np.random.randn(1, self.n_latent, ny, nx).astype('float32')

    def predict_snapshot(self, m, noise):
        X, single = self._normalized_q(m)
        noise = np.asarray(noise, dtype='float32').reshape(X.shape)
        Y = apply_function(self.decoder, X, noise, fun=self.generate)
        return self._denorm64(Y, single)

    def predict_mean_snapshot(self, m, M=100):
        X, single = self._normalized_q(m)
        acc = np.zeros_like(X)
        for _ in range(M):
            acc += apply_function(self.decoder, X, fun=self.generate)
        return self._denorm64(acc / M, single)

    def predict(self, ds, M=1000):
        X = self.x_scale.normalize(extract(ds, 'q').astype('float32'))
        Y, mean, var = apply_function(self.decoder, X, fun=self.generate_mean_var, M=M)
        shape = self._shape_of(ds)
        return make_dataset(
            q_forcing_advection=self.y_scale.denormalize(Y).reshape(shape),
            q_forcing_advection_mean=self.y_scale.denormalize(mean).reshape(shape),
            q_forcing_advection_var=self.y_scale.denormalize_var(var).reshape(shape))


def evaluate_prediction(net, ds, nruns=None, M=16):
    """cvae_regression.py:232-243 (= cgan_regression.py:197-208): offline scores of ``net.predict`` on ``nruns`` random runs."""
    nrun = np.shape(ds['q'])[0]
    idx = np.arange(nrun)
    if nruns is not None and nruns < len(idx):
        idx = np.random.choice(idx, nruns, replace=False)
    sub = {k: np.asarray(getattr(ds[k], 'values', ds[k]))[idx] for k in ('q', 'q_forcing_advection')}
    preds = net.predict(sub, M=M)
    scores = subgrid_scores(sub['q_forcing_advection'], _values(preds['q_forcing_advection_mean']),
                            _values(preds['q_forcing_advection']))
    return {k: scores[k] for k in ('L2_mean', 'L2_total', 'L2_residual', 'var_ratio')}


def _to_device(a, dev):
    """numpy array / CPU or CUDA tensor -> contiguous float32 tensor on ``dev``."""
    t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))
    return t.to(device=dev, dtype=torch.float32).contiguous()


def _values(v):
    return np.asarray(getattr(v, 'values', v))


def loss_to_log(optim_loss, log_train, log_test, name='L2_loss'):
    """cvae_regression.py:245-255 / cgan_regression.py:210-220 without xarray, with the variables of the shipped
    Google-Colab/{VAE,GAN}/stats.nc: per-epoch series of the optimisation losses, the train scores, the test scores (suffix
    ``_test``), ``var_ratio`` (epoch, lev) -- the reference's ``ds.update`` leaves the TEST values there, the rename only covers the
    L2 scores --, ``name`` = L2_total_test + L2_residual_test and the scalar ``Epoch_opt`` = the epoch (1-based) where it is smallest."""
    out = {k: list(v) for k, v in optim_loss.items()}
    for key in ('L2_mean', 'L2_total', 'L2_residual'):
        out[key] = [float(l[key]) for l in log_train]
        out[key + '_test'] = [float(l[key]) for l in log_test]
    out['var_ratio'] = np.array([np.asarray(l['var_ratio'], dtype=np.float64).reshape(2) for l in log_test])
    out[name] = [a + b for a, b in zip(out['L2_total_test'], out['L2_residual_test'])]
    epoch_opt = int(np.argmin(out[name])) + 1
    out['Epoch_opt'] = float(epoch_opt)
    return out, epoch_opt


class CVAETrainer(object):
    """Device state of the encoder / decoder pair; ``step`` = one iteration of the loop at cvae_regression.py:283-289."""

    def __init__(self, net, ny, nx, max_batch=64, device=None):
        self.net = net
        self.enc = Trainer(net.encoder, ny, nx, max_batch=max_batch, device=device)
        self.dec = Trainer(net.decoder, ny, nx, max_batch=max_batch, device=device)
        self._lib = _lib.load()
        dv = net.decoder_var
        self.decoder_var = -1.0 if dv == 'adaptive' else (1.0 if dv == 'fixed' else float(dv))

    def step(self, x, y, lr, eps=None, update=True):
        """x, y: (B, 2, ny, nx) float32 (numpy or torch); eps: the reparameterisation draw (default torch.randn on the device).
        Returns the six losses of ``compute_loss``."""
        dev = torch.device('cuda:%d' % self.enc.device)
        xd, yd = _to_device(x, dev), _to_device(y, dev)
        if xd.shape != yd.shape or xd.dim() != 4 or xd.shape[1] != 2:
            raise ValueError('expected x, y of shape (B, 2, ny, nx), got %s and %s' % (tuple(xd.shape), tuple(yd.shape)))
        ed = torch.randn(xd.shape, device=dev) if eps is None else \
            _to_device(eps, dev)
        out = (ctypes.c_double * 6)()
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check_train(self._lib.qgb_train_cvae_step(self.enc._h, self.dec._h, xd.data_ptr(), yd.data_ptr(), ed.data_ptr(),
                                                       xd.shape[0], 1, float(lr), self.decoder_var, 1 if update else 0, out,
                                                       stream), self.enc._h)
        if update:
            self.enc.steps += 1
            self.dec.steps += 1
        return dict(zip(LOSS_KEYS, [float(v) for v in out]))

    def sync(self):
        self.enc.sync_to(self.net.encoder)
        self.dec.sync_to(self.net.decoder)

    def close(self):
        self.enc.close()
        self.dec.close()


def train_CVAE(net, ds_train, ds_test, X_train, Y_train, num_epochs, batch_size, learning_rate, nruns=5, evaluate=True,
               noise=None):
    """cvae_regression.py:250-320: Adam over chain(encoder, decoder) + MultiStepLR(gamma 0.1), shuffled minibatches, epoch means
    of the six losses in ``optim_loss`` and the offline scores of nruns random train / test runs after every epoch
    (``evaluate=False`` skips them: the tests time the optimisation alone).  ``noise``: optional callable ``shape -> array``
    supplying the reparameterisation draws (parity tests); default = torch.randn on the device."""
    X_train, Y_train = np.asarray(X_train), np.asarray(Y_train)
    print('Training starts on device %s, number of samples %d' % (torch.cuda.get_device_name(0), len(X_train)))
    tr = CVAETrainer(net, X_train.shape[2], X_train.shape[3], max_batch=batch_size)
    optim_loss, log_train, log_test = {}, [], []
    t_s = time()
    for epoch in range(num_epochs):
        t_e = time()
        lr = multistep_lr(learning_rate, num_epochs, epoch)
        logger = AverageLoss(optim_loss)
        for x, y in minibatch(X_train, Y_train, batch_size=batch_size):
            logger.accumulate(optim_loss, tr.step(x, y, lr, eps=None if noise is None else noise(tuple(x.shape))), len(x))
        logger.average(optim_loss)
        tr.sync()                                   # the per-epoch scores use the current weights in eval mode
        if evaluate:
            log_train.append(evaluate_prediction(net, ds_train, nruns))
            log_test.append(evaluate_prediction(net, ds_test, nruns))
        t = time()
        msg = '[%d/%d] [%.2f/%.2f] MSE/KL: [%.3f, %.3f] Var: [%.3f,%.3f]' % (
            epoch + 1, num_epochs, t - t_e, (t - t_s) * (num_epochs / (epoch + 1) - 1), optim_loss['MSE'][-1],
            optim_loss['loss_KL'][-1], optim_loss['var_latent'][-1], optim_loss['var_aggr'][-1])
        if evaluate:
            msg += ' L2_mean: [%.3f,%.3f] L2_total: [%.3f,%.3f] L2_res: [%.3f,%.3f]' % (
                log_train[-1]['L2_mean'], log_test[-1]['L2_mean'], log_train[-1]['L2_total'], log_test[-1]['L2_total'],
                log_train[-1]['L2_residual'], log_test[-1]['L2_residual'])
        print(msg)
    tr.close()
    return optim_loss, log_train, log_test
