"""Snapshot datasets in the layout the reference's analysis code reads.

Reference: ``drop_vars(m.to_dataset())`` / ``concat_in_time`` (pyqg_generative/tools/simulate.py:16-60) and the per-member
``<n>.nc`` files that ``dataset_statistics`` opens with ``xr.open_mfdataset(combine='nested', concat_dim='run')``
(tools/comparison_tools.py:196-204).  xarray and netCDF4 are not dependencies of this package: files are written as
NetCDF-3 (64-bit offset) with ``scipy.io.netcdf_file``, which xarray opens natively.

Layout of one file (= one run, like the reference) :
  dims       time, lev (2), y, x, l (= ny), k (= nx/2+1)
  coords     time [days, attrs units='days'], lev [1, 2], x, y [m], l, k [rad/m]
  float32    q, u, v, psi (time, lev, y, x) [+ q_forcing_advection in forcing datasets];  Ubg, Qy (lev)
  float32    time-averaged spectral diagnostics of the LAST snapshot: KEspec, Ensspec (lev, l, k); KEflux, APEflux,
             APEgenspec, KEfrictionspec, entspec, paramspec, paramspec_KEflux, paramspec_APEflux, ENSflux, ENSgenspec,
             ENSfrictionspec, Dissspec, ENSDissspec, ENSparamspec (l, k); EKE (lev); attribute EKEdiss
  attrs      pyqg_params (str of the dict, :144), pyqg:<name> physical parameters like pyqg's to_dataset
``write_netcdf`` keeps a leading ``run`` dimension instead (one file for the whole local ensemble).
"""
import os

import numpy as np

PHYSICAL = ('q', 'u', 'v', 'psi')
FORCING = ('q_forcing_advection',)      # forcing datasets (generate_subgrid_forcing, tools/simulate.py:62-106) carry S next to q, u, v, psi
LAYERED_SPECTRA = ('KEspec', 'Ensspec')
PLANE_SPECTRA = ('KEflux', 'APEflux', 'APEgenspec', 'KEfrictionspec', 'entspec', 'paramspec', 'paramspec_KEflux',
                 'paramspec_APEflux', 'ENSflux', 'ENSgenspec', 'ENSfrictionspec', 'Dissspec', 'ENSDissspec', 'ENSparamspec')
_ATTR_KEYS = ('nx', 'ny', 'L', 'W', 'dt', 'rek', 'filterfac', 'beta', 'rd', 'delta', 'H1', 'U1', 'U2', 'tavestart', 'taveint')


def model_coords(m):
    """Coordinates and constant fields pyqg's ``to_dataset`` attaches (x, y, l, k, lev, Ubg, Qy) + 'pyqg:' attributes."""
    x = (np.arange(m.nx) + 0.5) * m.L / m.nx
    y = (np.arange(m.ny) + 0.5) * m.W / m.ny
    coords = dict(x=x, y=y, l=np.asarray(m.ll, dtype=np.float64), k=np.asarray(m.kk, dtype=np.float64),
                  lev=np.array([1, 2], dtype=np.int32), Ubg=np.asarray(m.Ubg, 'float32'), Qy=np.asarray(m.Qy, 'float32'))
    attrs = {'pyqg:%s' % k: float(getattr(m, k)) for k in _ATTR_KEYS if hasattr(m, k)}
    return coords, attrs


def _nc_var(f, name, data, dims, attrs=None):
    data = np.asarray(data)
    v = f.createVariable(name, data.dtype.newbyteorder('>').char if data.dtype.kind == 'i' else data.dtype.char, dims)
    v[:] = data
    for k, val in (attrs or {}).items():
        setattr(v, k, val)
    return v


def _write(path, ds, run_index=None):
    """ds: dict from run_simulation (arrays with a leading run axis).  run_index=None keeps the run dimension."""
    from scipy.io import netcdf_file
    q = np.asarray(ds['q'])
    nrun, ntime, nlev, ny, nx = q.shape
    with netcdf_file(path, 'w', version=2) as f:
        lead = ()
        if run_index is None:
            f.createDimension('run', nrun)
            lead = ('run',)
        for name, n in (('time', ntime), ('lev', nlev), ('y', ny), ('x', nx), ('l', ny), ('k', nx // 2 + 1)):
            f.createDimension(name, n)
        c = ds.get('coords', {})
        _nc_var(f, 'time', np.asarray(ds['time'], dtype=np.float64), ('time',), {'units': 'days', 'long_name': 'time'})
        _nc_var(f, 'lev', np.asarray(c.get('lev', np.arange(1, nlev + 1)), dtype=np.int32), ('lev',))
        for name in ('x', 'y', 'l', 'k'):
            if name in c:
                _nc_var(f, name, np.asarray(c[name], dtype=np.float64), (name,))
        for name in ('Ubg', 'Qy'):
            if name in c:
                _nc_var(f, name, np.asarray(c[name], dtype=np.float32), ('lev',))
        for name in PHYSICAL + FORCING:
            if name not in ds:
                if name in PHYSICAL:
                    raise KeyError(name)
                continue
            a = np.asarray(ds[name], dtype=np.float32)
            _nc_var(f, name, a if run_index is None else a[run_index], lead + ('time', 'lev', 'y', 'x'))
        for name in LAYERED_SPECTRA:
            if name in ds:
                _nc_var(f, name, np.asarray(ds[name], dtype=np.float32), ('lev', 'l', 'k'))
        for name in PLANE_SPECTRA:
            if name in ds:
                _nc_var(f, name, np.asarray(ds[name], dtype=np.float32), ('l', 'k'))
        if 'EKE' in ds:            # pyqg's scalar diagnostics (time averages like the spectra)
            _nc_var(f, 'EKE', np.asarray(ds['EKE'], dtype=np.float32), ('lev',))
        if 'EKEdiss' in ds:        # (scipy's NetCDF-3 writer has no scalar variables: the time average goes to an attribute)
            setattr(f, 'EKEdiss', float(ds['EKEdiss']))
        for k, v in ds.get('attrs', {}).items():
            setattr(f, k, v if isinstance(v, (int, float)) else str(v))


def write_netcdf(ds, path):
    """Whole local ensemble in one file, dims (run, time, lev, y, x)."""
    _write(path, ds, None)
    return path


def write_runs(ds, folder, first=0):
    """One ``<first + i>.nc`` per member, without a run dimension: the file layout of scripts/run_parameterized.py /
    tools/simulate.py:249-263 that ``dataset_statistics('folder/*.nc')`` concatenates along ``run``."""
    os.makedirs(folder, exist_ok=True)
    paths = []
    for i in range(np.asarray(ds['q']).shape[0]):
        paths.append(os.path.join(folder, '%d.nc' % (first + i)))
        _write(paths[-1], ds, i)
    return paths


def write_forecast(fc, path):
    """Ensemble forecast file of the reference's ``--forecast`` mode (tools/simulate.py:254-292): q, u, v, psi of run 0 and
    their ensemble means ``*_mean``, dims (time, lev, y, x), float32."""
    from scipy.io import netcdf_file
    ntime, nlev, ny, nx = np.asarray(fc['q']).shape
    with netcdf_file(path, 'w', version=2) as f:
        for name, n in (('time', ntime), ('lev', nlev), ('y', ny), ('x', nx)):
            f.createDimension(name, n)
        _nc_var(f, 'time', np.asarray(fc['time'], dtype=np.float64), ('time',), {'units': 'days', 'long_name': 'time'})
        _nc_var(f, 'lev', np.arange(1, nlev + 1, dtype=np.int32), ('lev',))
        for name in PHYSICAL:
            _nc_var(f, name, np.asarray(fc[name], dtype=np.float32), ('time', 'lev', 'y', 'x'))
            _nc_var(f, name + '_mean', np.asarray(fc[name + '_mean'], dtype=np.float32), ('time', 'lev', 'y', 'x'))
        for k, v in fc.get('attrs', {}).items():
            setattr(f, k, v if isinstance(v, (int, float)) else str(v))
    return path


def read_netcdf(path):
    """File -> dict of numpy arrays (+ 'attrs', 'dims'); for users without xarray and for the tests."""
    from scipy.io import netcdf_file
    out = {}
    with netcdf_file(path, 'r', mmap=False) as f:
        out['dims'] = {k: v for k, v in f.dimensions.items()}
        out['var_dims'] = {k: tuple(v.dimensions) for k, v in f.variables.items()}
        out['var_attrs'] = {k: dict(v._attributes) for k, v in f.variables.items()}
        for k, v in f.variables.items():
            a = np.array(v[:])
            out[k] = a.astype(a.dtype.newbyteorder('='))        # NetCDF-3 is big-endian on disk
        out['attrs'] = {k: (v.decode() if isinstance(v, bytes) else v) for k, v in f._attributes.items()}
    return out
