"""Shared plumbing of the four CNN closures (file loading, scaling, batched host <-> device marshalling)."""
import os

import numpy as np
import torch

from ..tools.cnn_tools import AndrewCNN, ChannelwiseScaler, apply_function, extract
from .parameterization import DeviceClosure


def make_dataset(**arrays):
    """xarray.Dataset with dims (run,time,lev,y,x) when xarray is installed, else a plain dict of numpy arrays."""
    try:
        import xarray as xr
        dims = ['run', 'time', 'lev', 'y', 'x']
        return xr.Dataset({k: xr.DataArray(v, dims=dims[-v.ndim:]) for k, v in arrays.items()})
    except ImportError:
        return dict(arrays)


class CNNClosure(DeviceClosure):
    def _load_state(self, net, path):
        sd = torch.load(path, map_location='cpu')
        net.load_state_dict(sd)

    def _read_scales(self, folder):
        self.x_scale = ChannelwiseScaler().read('x_scale.json', folder)
        self.y_scale = ChannelwiseScaler().read('y_scale.json', folder)

    @staticmethod
    def _q_of(m):
        """m.q as (B,2,ny,nx) + whether the caller passed an un-batched model."""
        q = np.asarray(m.q)
        return (q[None], True) if q.ndim == 3 else (q, False)

    def _normalized_q(self, m):
        q, single = self._q_of(m)
        return self.x_scale.normalize(q.astype('float32')), single

    def _denorm64(self, Y, single):
        out = self.y_scale.denormalize(Y)
        out = out[0] if single else out
        return out.astype('float64')

    def _shape_of(self, ds):
        q = ds['q']
        return tuple(getattr(q, 'shape'))
