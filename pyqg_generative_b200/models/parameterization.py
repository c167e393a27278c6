"""Closure base classes: the pyqg ``QParameterization`` algebra and the reference ``Parameterization.__call__``.

* ``QParameterization`` / ``WeightedParameterization`` / ``CompositeParameterization`` restate
  pyqg 0.7.2 ``pyqg/parameterizations.py`` (``weight * model`` and ``a + b`` are used at
  pyqg_generative/tools/simulate.py:242,244,259).
* ``Parameterization.__call__(m)`` follows pyqg_generative/models/parameterization.py:23-34.
* ``DeviceClosure`` adds what the reference does not have: the closure can be *attached* to an
  ``EnsembleQGModel`` so that sampling, CNN forward, denormalisation and injection run inside the engine.
"""
import ctypes

import numpy as np

from .. import _lib


class QParameterization(object):
    parameterization_type = 'q_parameterization'

    def __call__(self, m):
        raise NotImplementedError

    def __add__(self, other):
        return CompositeParameterization(self, other)

    def __mul__(self, constant):
        return WeightedParameterization(self, constant)

    __rmul__ = __mul__


class CompositeParameterization(QParameterization):
    def __init__(self, *params):
        self.params = params

    def __call__(self, m):
        return np.sum([np.array(p(m)) for p in self.params], axis=0)


class WeightedParameterization(QParameterization):
    def __init__(self, param, weight):
        self.param = param
        self.weight = weight

    def __call__(self, m):
        inner = self.param
        while isinstance(inner, WeightedParameterization):
            inner = inner.param
        if getattr(m, '_closure', None) is inner and inner is not None:
            # attached device closure: the engine already applies the (product of the) weights folded in at attach time
            return inner(m)
        return np.array(self.param(m)) * self.weight


class Parameterization(QParameterization):
    """Reference ``models/parameterization.py``: subclasses provide generate_latent_noise / predict_snapshot /
    predict_mean_snapshot / predict."""

    def generate_latent_noise(self, ny, nx):
        raise NotImplementedError

    def predict_snapshot(self, m, noise):
        raise NotImplementedError

    def predict_mean_snapshot(self, m, M=100):
        raise NotImplementedError

    def predict(self, ds, M=1000):
        raise NotImplementedError

    def __call__(self, m):
        if getattr(m, '_closure', None) is self:
            return m.closure_eval()          # everything happens on the device
        demean = lambda x: x - x.mean(axis=(-2, -1), keepdims=True)
        if m.sampling_type == 'deterministic':
            m.PV_forcing = demean(self.predict_mean_snapshot(m))
        else:
            nb = np.shape(m.q)[0] if np.ndim(m.q) == 4 else 0     # batched host model: independent noise per member

            def latent_noise():
                if not nb:
                    return self.generate_latent_noise(m.ny, m.nx)
                return np.concatenate([np.asarray(self.generate_latent_noise(m.ny, m.nx)).reshape((1, -1, m.ny, m.nx))
                                       for _ in range(nb)])
            if m.noise_sampler.update(latent_noise):
                m.PV_forcing = demean(self.predict_snapshot(m, m.noise_sampler.noise))
        return m.PV_forcing


class DeviceClosure(Parameterization):
    """A closure whose networks can be loaded into a libqgb200 engine handle."""

    closure_kind = _lib.CLOSURE_NONE
    n_mean = 100

    def _nets(self):
        """list of AndrewCNN objects in engine order (net 0, net 1)."""
        raise NotImplementedError

    def _attach(self, model, weight=1.0, precision='fp32'):
        lib = _lib.load()
        for i, net in enumerate(self._nets()):
            arr, keep = net.c_layers()
            _lib.check(lib.qgb_cnn_load(model._h, self.closure_kind, i, len(arr), arr), model._h)
        xs = (ctypes.c_float * 2)(*[float(v) for v in np.asarray(self.x_scale.std).reshape(-1)[:2]])
        ys = (ctypes.c_float * 2)(*[float(v) for v in np.asarray(self.y_scale.std).reshape(-1)[:2]])
        prec = _lib.PRECISIONS[precision]
        _lib.check(lib.qgb_closure_config(model._h, xs, ys, float(weight), prec), model._h)
