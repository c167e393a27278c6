"""MeanVarModel (GZ) closure: inference surface of pyqg_generative/models/mean_var_model.py on libqgb200.

``net_mean`` + ``VarCNN`` (softplus head, :14-17); ``generate_latent_noise`` :102-103 (float64, (2,ny,nx));
``predict_snapshot = y_std * (mean + noise * sqrt(var))`` :105-109; ``predict_mean_snapshot`` :111-115;
``predict`` :117-134; ``fit`` :41-66 (two-stage: the mean network on the forcing, then the softplus network on the squared
residuals) and ``save_model`` :68-81 on the device-side trainer (tools/cnn_tools.py ``train``).
"""
import os
from os.path import exists

import numpy as np
import torch

from .. import _lib
from ..tools.cnn_tools import AndrewCNN, apply_function, extract, prepare_PV_data, save_model_args, train, write_log
from ._cnn_closure import CNNClosure, make_dataset


class VarCNN(AndrewCNN):
    """AndrewCNN followed by softplus (positive variance)."""
    softplus_output = True

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

    def fit(self, ds_train, ds_test, num_epochs=50, batch_size=64, learning_rate=0.001):
        """mean_var_model.py:41-66."""
        os.makedirs(self.folder, exist_ok=True)
        X_train, Y_train, X_test, Y_test, self.x_scale, self.y_scale = prepare_PV_data(ds_train, ds_test)
        if self.load_mean(self.folder):
            print('Net mean is loaded instead of training')
        else:
            train(self.net_mean, X_train, Y_train, X_test, Y_test, num_epochs, batch_size, learning_rate)
        rsq_train = (Y_train - apply_function(self.net_mean, X_train)) ** 2
        rsq_test = (Y_test - apply_function(self.net_mean, X_test)) ** 2
        train(self.net_var, X_train, rsq_train, X_test, rsq_test, num_epochs, batch_size, learning_rate)
        self.save_model()

    def save_model(self):
        """mean_var_model.py:68-81."""
        os.makedirs(self.folder, exist_ok=True)
        torch.save(self.net_mean.state_dict(), '%s/net_mean.pt' % self.folder)
        torch.save(self.net_var.state_dict(), '%s/net_var.pt' % self.folder)
        self.x_scale.write('x_scale.json', folder=self.folder)
        self.y_scale.write('y_scale.json', folder=self.folder)
        save_model_args('MeanVarModel', folder=self.folder, hidden_channels=self.hidden_channels)
        if hasattr(self.net_mean, 'log_dict'):          # nothing to save if the mean network was read from file
            write_log(self.net_mean.log_dict, '%s/stats_mean.nc' % self.folder)
        write_log(self.net_var.log_dict, '%s/stats_var.nc' % self.folder)

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
