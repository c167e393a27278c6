"""CPU ORACLE (test infrastructure, NOT product code) -- numpy restatement of the coarse-graining path.

Follows /root/reference/pyqg_generative/tools/operators.py:
  cut_off :117-132, model_filter :92-99, gauss_filter :84-90, Operator1/2/4/5 :204-217,
  fft_interpolate :134-190, divergence :241-247, advect :249-268, apply_operator_to_model :219-236,
  PV_subgrid_forcing :283-287;  and tools/simulate.py:147-168 (set_initial_condition).
Written against plain (nlev, ny, nx) / (ny, nx) float64 numpy arrays (the ``array_format`` decorator's numpy branch,
operators.py:44-49).  Pinned by tests/golden/operators_*.npz, produced by the unmodified reference functions.
"""
import numpy as np

from .pyqg_shim import QGModel


def _grid(n, L=1e6):
    dk = 2 * np.pi / L
    ll = dk * np.append(np.arange(0., n / 2), np.arange(-n / 2, 0.))
    kk = dk * np.arange(0., n // 2 + 1)
    k, l = np.meshgrid(kk, ll)
    return k, l, L / n


def _per_layer(f):
    def wrapper(X, nc=None):
        X = np.asarray(X)
        if X.ndim == 2:
            return f(X, nc)
        if X.ndim == 3:
            return np.stack([f(x, nc) for x in X])
        raise ValueError('numpy array should be 2 or 3 dimensional')
    return wrapper


@_per_layer
def cut_off(X, nc):
    if nc % 2 != 0:
        raise ValueError('nc must be even')
    ratio = X.shape[0] / nc
    n = nc // 2
    Xf = np.fft.rfftn(X)
    trunc = np.vstack((Xf[:n, :n + 1], Xf[-n:, :n + 1])) / ratio ** 2
    trunc[n, 0] = 0          # FILTER_2h_HARMONICS (operators.py:8,126-130)
    trunc[:, n] = 0
    return np.fft.irfftn(trunc)


@_per_layer
def model_filter(X, nc=None):
    n = X.shape[0]
    k, l, dx = _grid(n)
    wvx = np.sqrt((k * dx) ** 2 + (l * dx) ** 2)
    filtr = np.exp(-23.6 * (wvx - 0.65 * np.pi) ** 4)
    filtr[wvx <= 0.65 * np.pi] = 1.
    return np.fft.irfftn(np.fft.rfftn(X) * filtr)


@_per_layer
def gauss_filter(X, nc):
    n = X.shape[0]
    ratio = n / nc
    k, l, dx = _grid(n)
    return np.fft.irfftn(np.fft.rfftn(X) * np.exp(-(k ** 2 + l ** 2) * (ratio * dx) ** 2 / 24))


def Operator1(X, nc):
    return model_filter(cut_off(X, nc))


def Operator2(X, nc):
    return gauss_filter(cut_off(X, nc), nc // 2)


def Operator4(X, nc):
    return model_filter(Operator2(X, nc))


def Operator5(X, nc):
    return cut_off(X, nc)


def fft_interpolate(x, n, N, truncate_2h=True):
    x = np.asarray(x)
    if x.shape[-2] != n or x.shape[-1] != n:
        raise ValueError('Input variable must be n*n points')
    if n % 2 != 0 or N % 2 != 0:
        raise ValueError('Grid sizes (n,N) must be even')
    nn = min(n // 2, N // 2)
    xf = np.fft.rfftn(x, axes=(-2, -1))
    Xf = np.zeros(x.shape[:-2] + (N, N // 2 + 1), dtype='complex128')
    if truncate_2h:
        xf[..., nn, 0] = 0
    Xf[..., :nn, :nn + 1] = xf[..., :nn, :nn + 1]
    Xf[..., -nn:, :nn + 1] = xf[..., -nn:, :nn + 1]
    if truncate_2h:
        Xf[..., nn, 0] = 0
        Xf[..., :, nn] = 0
    return np.fft.irfftn(Xf, axes=(-2, -1)) * (N / n) ** 2


def divergence(fx, fy):
    n = fx.shape[-1]
    k, l, _ = _grid(n)
    ddx = lambda x: np.fft.irfftn(np.fft.rfftn(x, axes=(-2, -1)) * 1j * k, axes=(-2, -1))
    ddy = lambda x: np.fft.irfftn(np.fft.rfftn(x, axes=(-2, -1)) * 1j * l, axes=(-2, -1))
    return ddx(fx) + ddy(fy)


def advect(var, u, v, dealias='none'):
    if dealias == 'none':
        return divergence(var * u, var * v)
    if dealias == '2/3-rule':
        n = u.shape[-1]
        k, l, dx = _grid(n)
        wvx = np.sqrt((k * dx) ** 2 + (l * dx) ** 2)
        with np.errstate(over='ignore', under='ignore'):
            filtr = np.exp(-1e20 * (wvx - 0.65 * np.pi) ** 4)
        filtr[wvx <= 0.65 * np.pi] = 1.
        f = lambda x: np.fft.irfftn(np.fft.rfftn(x, axes=(-2, -1)) * filtr, axes=(-2, -1))
        _var, _u, _v = f(var), f(u), f(v)
        return f(divergence(_var * _u, _var * _v))
    if dealias == '3/2-rule':
        n = u.shape[-1]
        N = int((n * 3) // 2)
        _var, _u, _v = (fft_interpolate(a, n, N) for a in (var, u, v))
        return divergence(fft_interpolate(_var * _u, N, n), fft_interpolate(_var * _v, N, n))
    raise ValueError('dealias should be none or 2/3-rule or 3/2-rule')


def apply_operator_to_model(q, nc, operator, pyqg_params):
    qf = operator(np.asarray(q, dtype='float64'), nc)
    params = dict(pyqg_params)
    params.update(dict(nx=qf.shape[1], log_level=0))
    m = QGModel(**params)
    m.q = qf
    m._invert()
    m._calc_derived_fields()
    return m


def PV_subgrid_forcing(q, nc, operator, pyqg_params, dealias='none'):
    m = apply_operator_to_model(q, 1, lambda x, nc: x, pyqg_params)
    mf = apply_operator_to_model(q, nc, operator, pyqg_params)
    forcing = advect(mf.q, mf.u, mf.v, dealias) - operator(advect(m.q, m.u, m.v, dealias), nc)
    return forcing, mf, m


def set_initial_condition(m, rng=np.random):
    """tools/simulate.py:147-168 (JAMES-paper IC).  ``rng`` must provide ``rand`` (np.random or a RandomState)."""
    q2d = 1e-7 * rng.rand(m.ny, m.nx)
    q2d -= q2d.mean(axis=(-2, -1), keepdims=True)
    q2d *= np.sqrt(m.nx * m.ny / 64 ** 2)
    q1d = 1e-6 * (np.ones((m.ny, 1)) * rng.rand(1, m.nx))
    q1d -= q1d.mean(axis=(-2, -1), keepdims=True)
    q1d *= np.sqrt(m.nx / 64)
    noise = q1d + q2d
    Xf = np.fft.rfftn(noise)
    noise = np.fft.irfftn(Xf * (m.wv < np.pi / (m.L / 32)))
    m.set_q1q2(noise, 0 * m.x)
    m._invert()
