"""CGANRegression closure: inference surface of pyqg_generative/models/cgan_regression.py on libqgb200.

Kept: constructor kwargs of ``model_args.json`` (regression, nx, generator, folder, div, hidden_channels), file
formats (G.pt, x_scale.json, y_scale.json), and ``generate`` :133-137, ``generate_latent_noise`` :154-155,
``predict_snapshot`` :157-162, ``predict_mean_snapshot`` :164-171, ``predict`` :173-189, ``predict_ensemble`` :191-195.
Out of scope: the discriminator and WGAN-GP training (:197-344); ``D.pt`` is therefore not required.
"""
import os
from os.path import exists

import numpy as np
import torch

from .. import _lib
from ..tools.cnn_tools import AndrewCNN, apply_function, extract
from ._cnn_closure import CNNClosure, batched_mean_var, make_dataset


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
        self.load_GAN(folder)

    def _nets(self):
        return [self.G]

    def load_GAN(self, folder):
        if exists('%s/G.pt' % folder):
            self._load_state(self.G, '%s/G.pt' % folder)
            self._read_scales(folder)
            return True
        return False

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
