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


def batched_mean_var(generate, x, M, images_per_forward=1024):
    """``(y_first, mean, var)`` of M generator samples per input, like the reference's
    ``y = torch.stack([generate(x) for _ in range(M)]); y[0], y.mean(0), y.var(0)`` (models/cgan_regression.py:164-166,
    cvae_regression.py:138-140), but with the M noise realisations folded into the batch axis (up to ``images_per_forward``
    images per forward pass, running sums in float64) so that the tensor-core kernels see full persistent grids instead of
    1000 launches of 64 images, and no (M, B, 2, ny, nx) tensor is materialised."""
    import torch
    B = x.shape[0]
    rep = max(1, min(int(M), images_per_forward // max(B, 1)))
    first = None
    s = torch.zeros(x.shape[0:1] + (2,) + x.shape[2:], dtype=torch.float64, device=x.device)
    s2 = torch.zeros_like(s)
    done = 0
    while done < M:
        r = min(rep, M - done)
        y = generate(x.repeat((r, 1, 1, 1))).reshape((r, B) + tuple(s.shape[1:]))
        if first is None:
            first = y[0].clone()
        yd = y.double()
        s += yd.sum(dim=0)
        s2 += (yd * yd).sum(dim=0)
        done += r
    mean = s / M
    var = (s2 - M * mean * mean) / max(M - 1, 1)           # unbiased, like torch.var
    return first, mean.float(), var.clamp_min(0).float()
