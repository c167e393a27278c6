"""CVAERegression closure: inference surface of pyqg_generative/models/cvae_regression.py on libqgb200.

The decoder is the same AndrewCNN(4 -> 2) as the GAN generator (:45); ``generate`` :114-118,
``generate_latent_noise`` :128-129, ``predict_snapshot`` :131-136, ``predict_mean_snapshot`` :138-145,
``predict`` :147-163.  The encoder and ELBO training (:47-49, :165-320) are out of scope; ``encoder.pt`` is ignored.
"""
from os.path import exists

import numpy as np
import torch

from .. import _lib
from ..tools.cnn_tools import AndrewCNN, apply_function, extract
from ._cnn_closure import CNNClosure, batched_mean_var, make_dataset


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
        self.load_model(folder)

    def _nets(self):
        return [self.decoder]

    def load_model(self, folder):
        if exists('%s/decoder.pt' % folder):
            self._load_state(self.decoder, '%s/decoder.pt' % folder)
            self._read_scales(folder)
            return True
        return False

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
