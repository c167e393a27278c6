"""OLSModel closure (deterministic CNN): inference surface of pyqg_generative/models/ols_model.py on libqgb200.

``generate_latent_noise`` returns 0 (:68-69), ``predict_snapshot`` :71-75, ``predict`` :77-97; ``fit`` is out of scope.
"""
from os.path import exists

import numpy as np

from .. import _lib
from ..tools.cnn_tools import AndrewCNN, apply_function, extract
from ._cnn_closure import CNNClosure, make_dataset


class OLSModel(CNNClosure):
    closure_kind = _lib.CLOSURE_OLS

    def __init__(self, div=False, batch_norm=True, bias=True, final_activation='None',
                 hidden_channels=[128, 64, 32, 32, 32, 32, 32], folder='model', precision='fp32'):
        self.folder = folder
        self.div, self.batch_norm, self.bias, self.final_activation = div, batch_norm, bias, final_activation
        self.hidden_channels = hidden_channels
        self.net = AndrewCNN(2, 2, div=div, batch_norm=batch_norm, bias=bias, final_activation=final_activation,
                             hidden_channels=hidden_channels, precision=precision)
        self.load_model(folder)

    def _nets(self):
        return [self.net]

    def load_model(self, folder):
        if exists('%s/net.pt' % folder):
            self._load_state(self.net, '%s/net.pt' % folder)
            self._read_scales(folder)
            return True
        return False

    def generate_latent_noise(self, ny, nx):
        return 0

    def predict_snapshot(self, m, noise=None):
        X, single = self._normalized_q(m)
        return self._denorm64(apply_function(self.net, X), single)

    def predict_mean_snapshot(self, m, M=100):
        return self.predict_snapshot(m)

    def predict(self, ds, M=1000):
        X = self.x_scale.normalize(extract(ds, 'q').astype('float32'))
        Y = self.y_scale.denormalize(apply_function(self.net, X)).reshape(self._shape_of(ds))
        return make_dataset(q_forcing_advection=Y, q_forcing_advection_mean=Y, q_forcing_advection_var=Y * 0)
