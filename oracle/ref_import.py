"""CPU ORACLE helper (test infrastructure, NOT product code).

Imports the reference's own in-tree Python modules *unmodified* from /root/reference on top of the
pyqg shim (``oracle/pyqg_shim.py``), following SURVEY.md Appendix E.  /root/reference only exists in
the build container, so this module is used by ``tests/golden/make_golden.py`` (fixture generation)
and by container-only cross-checks; nothing that runs on the GPU box may depend on it.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('QGB_REFERENCE_ROOT', '/root/reference')


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'pyqg_generative'))


def install_stubs():
    """Register stand-ins for pyqg / xarray / gcm_filters in ``sys.modules`` (idempotent)."""
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import pyqg_shim

    if 'pyqg' not in sys.modules or not hasattr(sys.modules['pyqg'], '_qgb_shim'):
        pyqg = types.ModuleType('pyqg')
        pyqg._qgb_shim = True
        pyqg.QGModel = pyqg_shim.QGModel
        pyqg.Parameterization = pyqg_shim.Parameterization
        pyqg.QParameterization = pyqg_shim.QParameterization
        pyqg.UVParameterization = pyqg_shim.UVParameterization
        params = types.ModuleType('pyqg.parameterizations')
        for name in ('Parameterization', 'QParameterization', 'UVParameterization',
                     'CompositeParameterization', 'WeightedParameterization'):
            setattr(params, name, getattr(pyqg_shim, name))
        pyqg.parameterizations = params
        sys.modules['pyqg'] = pyqg
        sys.modules['pyqg.parameterizations'] = params

    if 'xarray' not in sys.modules:
        try:
            import xarray  # noqa: F401
        except Exception:
            xr = types.ModuleType('xarray')

            class DataArray(object):
                def __init__(self, data=None, attrs=None, **kw):
                    self.values = data
                    self.attrs = attrs or {}

            class Dataset(dict):
                pass

            xr.DataArray = DataArray
            xr.Dataset = Dataset
            sys.modules['xarray'] = xr

    if 'gcm_filters' not in sys.modules:
        try:
            import gcm_filters  # noqa: F401
        except Exception:
            sys.modules['gcm_filters'] = types.ModuleType('gcm_filters')

    if 'pyqg_parameterization_benchmarks' not in sys.modules:
        ppb = types.ModuleType('pyqg_parameterization_benchmarks')
        utils = types.ModuleType('pyqg_parameterization_benchmarks.utils')
        utils.FeatureExtractor = object
        ppb.utils = utils
        sys.modules['pyqg_parameterization_benchmarks'] = ppb
        sys.modules['pyqg_parameterization_benchmarks.utils'] = utils


def _give_specs():
    """torch's lazy imports probe optional packages with importlib.util.find_spec, which rejects modules without __spec__."""
    from importlib.machinery import ModuleSpec
    for name in ('pyqg', 'pyqg.parameterizations', 'xarray', 'gcm_filters', 'pyqg_parameterization_benchmarks',
                 'pyqg_parameterization_benchmarks.utils'):
        mod = sys.modules.get(name)
        if mod is not None and getattr(mod, '__spec__', None) is None:
            mod.__spec__ = ModuleSpec(name, None)


def import_reference():
    """Return the reference's ``pyqg_generative`` package imported unmodified (container only)."""
    if not reference_available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)
    install_stubs()
    _give_specs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import pyqg_generative  # noqa: F401
    import pyqg_generative.tools.cnn_tools  # noqa: F401
    import pyqg_generative.tools.operators  # noqa: F401
    import pyqg_generative.tools.stochastic_pyqg  # noqa: F401
    import pyqg_generative.tools.parameters  # noqa: F401
    import pyqg_generative.models.parameterization  # noqa: F401
    import pyqg_generative.models.cvae_regression  # noqa: F401
    import pyqg_generative.models.mean_var_model  # noqa: F401
    import pyqg_generative.models.ols_model  # noqa: F401
    import pyqg_generative.models.cgan_regression  # noqa: F401
    return sys.modules['pyqg_generative']
