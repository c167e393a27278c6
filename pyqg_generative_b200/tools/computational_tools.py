"""Offline scores of a stochastic subgrid model: ``subgrid_scores`` of pyqg_generative/tools/computational_tools.py:39-83 and the
power ``spectrum`` it uses (tools/spectral_tools.py:7-101, ``spectrum(time=slice(None, None))``), on plain numpy arrays of shape
(run, time, lev, y, x).  The training loops evaluate them once per epoch (models/cgan_regression.py:197-208,
models/cvae_regression.py:232-243); they are host post-processing, off the step path.
"""
import numpy as np

from .spectral_tools import calc_ispec


class _Grid(object):
    """Wavenumber grid of ``pyqg.QGModel(nx=n)`` (L = 1e6), the five attributes ``calc_ispec`` reads."""

    def __init__(self, n, L=1e6):
        self.dk = self.dl = 2 * np.pi / L
        self.kk = self.dk * np.arange(0, n // 2 + 1)
        self.ll = self.dl * np.append(np.arange(0, n // 2), np.arange(-n // 2, 0))
        k, l = np.meshgrid(self.kk, self.ll)
        self.wv = np.sqrt(k ** 2 + l ** 2)


def power_spectrum(x):
    """spectrum(type='power', averaging=False, truncate=False, time=slice(None, None))(x): (lev, k) isotropic power spectral
    density of the run- and time-mean |rfft2(x) / M|^2 of each layer."""
    x = np.asarray(x, dtype=np.float64)
    M = x.shape[-1] * x.shape[-2]
    af2 = (np.abs(np.fft.rfftn(x, axes=(-2, -1)) / M) ** 2).mean(axis=(0, 1))
    grid = _Grid(x.shape[-1])
    return np.stack([calc_ispec(grid, af2[z], averaging=False, truncate=False)[1] for z in range(af2.shape[0])])


def subgrid_scores(true, mean, gen):
    """computational_tools.py:39-83: dict with R2_mean, L2_mean (mean prediction against the truth, per layer then averaged),
    R2_total, L2_total (power spectrum of the generated forcing against the true one), R2_residual, L2_residual (spectra of the
    residuals gen - mean and true - mean) and var_ratio (lev,)."""
    true, mean, gen = (np.asarray(a, dtype=np.float64) for a in (true, mean, gen))

    def R2(x, x_true, axes):
        return float((1 - ((x - x_true) ** 2).mean(axis=axes) / x_true.var(axis=axes)).mean())

    def L2(x, x_true, axes):
        return float(((((x - x_true) ** 2).mean(axis=axes) / (x_true ** 2).mean(axis=axes)) ** 0.5).mean())

    field_axes = (0, 1, 3, 4)
    out = dict(R2_mean=R2(mean, true, field_axes), L2_mean=L2(mean, true, field_axes))
    sp_true, sp_gen = power_spectrum(true), power_spectrum(gen)
    out['R2_total'], out['L2_total'] = R2(sp_gen, sp_true, (1,)), L2(sp_gen, sp_true, (1,))
    sp_true_res, sp_gen_res = power_spectrum(true - mean), power_spectrum(gen - mean)
    out['R2_residual'], out['L2_residual'] = R2(sp_gen_res, sp_true_res, (1,)), L2(sp_gen_res, sp_true_res, (1,))
    out['var_ratio'] = ((gen - mean) ** 2).mean(axis=field_axes) / ((true - mean) ** 2).mean(axis=field_axes)
    return out
