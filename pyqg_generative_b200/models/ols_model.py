"""OLSModel closure (deterministic CNN): inference surface of pyqg_generative/models/ols_model.py on libqgb200.

``generate_latent_noise`` returns 0 (:68-69), ``predict_snapshot`` :71-75, ``predict`` :77-97; ``fit`` :36-46 and
``save_model`` :48-57 on the device-side trainer (tools/cnn_tools.py ``train``).
"""
import os
from os.path import exists

import numpy as np
import torch

from .. import _lib
from ..tools.cnn_tools import AndrewCNN, apply_function, extract, prepare_PV_data, save_model_args, train, write_log
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

    def fit(self, ds_train, ds_test, num_epochs=50, batch_size=64, learning_rate=0.001):
        """ols_model.py:36-46."""
        X_train, Y_train, X_test, Y_test, self.x_scale, self.y_scale = prepare_PV_data(ds_train, ds_test)
        train(self.net, X_train, Y_train, X_test, Y_test, num_epochs, batch_size, learning_rate)
        self.save_model()

    def save_model(self):
        """ols_model.py:48-57."""
        os.makedirs(self.folder, exist_ok=True)
        torch.save(self.net.state_dict(), '%s/net.pt' % self.folder)
        self.x_scale.write('x_scale.json', folder=self.folder)
        self.y_scale.write('y_scale.json', folder=self.folder)
        save_model_args('OLSModel', folder=self.folder, div=self.div, batch_norm=self.batch_norm, bias=self.bias,
                        final_activation=self.final_activation, hidden_channels=self.hidden_channels)
        write_log(self.net.log_dict, '%s/stats.nc' % self.folder)

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
