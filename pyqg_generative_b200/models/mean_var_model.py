"""MeanVarModel (GZ) closure: inference surface of pyqg_generative/models/mean_var_model.py on libqgb200.

``net_mean`` + ``VarCNN`` (softplus head, :14-17); ``generate_latent_noise`` :102-103 (float64, (2,ny,nx));
``predict_snapshot = y_std * (mean + noise * sqrt(var))`` :105-109; ``predict_mean_snapshot`` :111-115;
``predict`` :117-134.  ``fit`` (:41-66) is out of scope.
"""
from os.path import exists

import numpy as np

from .. import _lib
from ..tools.cnn_tools import AndrewCNN, apply_function, extract
from ._cnn_closure import CNNClosure, make_dataset


class VarCNN(AndrewCNN):
    """AndrewCNN followed by softplus (positive variance)."""

    def forward(self, x, softplus=True, precision=None):
        return super().forward(x, softplus=True, precision=precision)

    __call__ = forward


class MeanVarModel(CNNClosure):
    closure_kind = _lib.CLOSURE_GZ

    def __init__(self, folder='model', hidden_channels=[128, 64, 32, 32, 32, 32, 32], precision='fp32'):
        self.folder = folder
        self.hidden_channels = hidden_channels
        self.net_mean = AndrewCNN(2, 2, hidden_channels=hidden_channels, precision=precision)
        self.net_var = VarCNN(2, 2, hidden_channels=hidden_channels, precision=precision)
        self.load_mean(folder)
        self.load_var(folder)

    def _nets(self):
        return [self.net_mean, self.net_var]

    def load_mean(self, folder):
        if exists('%s/net_mean.pt' % folder):
            self._load_state(self.net_mean, '%s/net_mean.pt' % folder)
            self._read_scales(folder)
            return True
        return False

    def load_var(self, folder):
        if exists('%s/net_var.pt' % folder):
            self._load_state(self.net_var, '%s/net_var.pt' % folder)
            return True
        return False

    def generate_latent_noise(self, ny, nx):
        return np.random.randn(2, ny, nx)

    def predict_snapshot(self, m, noise):
        X, single = self._normalized_q(m)
        noise = np.asarray(noise, dtype='float64').reshape(X.shape)
        Y = apply_function(self.net_mean, X) + noise * (apply_function(self.net_var, X)) ** 0.5
        return self._denorm64(Y, single)

    def predict_mean_snapshot(self, m, M=100):
        X, single = self._normalized_q(m)
        return self._denorm64(apply_function(self.net_mean, X), single)

    def predict(self, ds, M=1000):
        X = self.x_scale.normalize(extract(ds, 'q').astype('float32'))
        shape = self._shape_of(ds)
        mean = self.y_scale.denormalize(apply_function(self.net_mean, X)).reshape(shape)
        var = self.y_scale.denormalize_var(apply_function(self.net_var, X)).reshape(shape)
        Y = mean + np.sqrt(var) * np.random.randn(*var.shape)
        return make_dataset(q_forcing_advection=Y, q_forcing_advection_mean=mean, q_forcing_advection_var=var)
