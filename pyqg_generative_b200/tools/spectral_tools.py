"""Isotropic spectra of half-plane (rfft2-layout) spectral densities: ``calc_ispec`` of
pyqg_generative/tools/spectral_tools.py:103-180, the function the reference's comparison code applies to the time-averaged
pyqg diagnostics (KEspec, Ensspec, KEflux, ..., tools/comparison_tools.py:152-157, 234-247).  Host numpy: it post-processes
the (all-reduced) device accumulators at output cadence and is not on the step path.
"""
import numpy as np


class GridView(object):
    """The five grid attributes ``calc_ispec`` reads from a pyqg model (ll, kk, dl, dk, wv), for callers that only have
    the wavenumber arrays (``parallel.calc_ispec``)."""

    def __init__(self, k, l):
        k, l = np.asarray(k, dtype=np.float64), np.asarray(l, dtype=np.float64)
        self.kk, self.ll = k[0], l[:, 0]
        self.dk, self.dl = self.kk[1] - self.kk[0], self.ll[1] - self.ll[0]
        self.wv = np.sqrt(k ** 2 + l ** 2)


def calc_ispec(model, _var_dens, averaging=True, truncate=True, nd_wavenumber=False, nfactor=1):
    """Isotropic spectrum of ``_var_dens`` = |rfft2(signal)|^2 / M^2 on the grid of ``model`` (any object with ll, kk, dl,
    dk, wv: an ``EnsembleQGModel``, a pyqg model, ``GridView``).  Returns ``(kr, phr)``: bin centres and spectral density,
    normalised so that ``signal.var() == phr.sum() * (kr[1] - kr[0])`` (summation mode).

    Semantics follow the reference exactly (they are what its published spectra were computed with):
      * the self-conjugate columns k = 0 and k = N/2 are halved, every bin is doubled at the end (Hermitian half plane);
      * bins are [kr, kr + dkr) with LEFT edges ``arange(kmin, kmax - dkr, dkr)``, kmin = min(dk, dl),
        dkr = nfactor * sqrt(dk^2 + dl^2), kmax = the inscribed circle (``truncate``) or the corner of the spectral box;
      * ``averaging``: mean over the shell (closed upper edge, ``<=``) times the shell's half circumference in grid units,
        pi * (kr + dkr/2) / (dk dl); an empty shell gives 0.  Otherwise: sum over the half-open shell / dkr (Parseval holds);
      * ``nd_wavenumber``: wavenumbers in units of kmin, density rescaled to keep the integral.
    """
    dens = np.array(_var_dens, dtype=np.float64, copy=True)
    dens[..., 0] *= 0.5
    dens[..., -1] *= 0.5
    kmax = min(np.abs(model.ll).max(), np.abs(model.kk).max()) if truncate \
        else np.hypot(np.abs(model.ll).max(), np.abs(model.kk).max())
    kmin = min(model.dk, model.dl)
    dkr = np.sqrt(model.dk ** 2 + model.dl ** 2) * nfactor
    left = np.arange(kmin, kmax - dkr, dkr)
    wv = np.asarray(model.wv)
    phr = np.zeros(left.size)
    for i, lo in enumerate(left):
        if averaging:
            shell = (wv >= lo) & (wv <= lo + dkr)
            if shell.any():
                phr[i] = dens[shell].mean() * (lo + dkr / 2) * np.pi / (model.dk * model.dl)
        else:
            shell = (wv >= lo) & (wv < lo + dkr)
            phr[i] = dens[shell].sum() / dkr
    phr *= 2
    kr = left + dkr / 2
    if nd_wavenumber:
        kr, phr = kr / kmin, phr * kmin
    return kr, phr
