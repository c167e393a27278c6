"""Driver entry points of pyqg_generative/tools/simulate.py for the batched GPU engine.

``run_simulation(pyqg_params, parameterization, q_init, sampling_freq)`` (:108-145) and
``set_initial_condition(m)`` (:147-168) keep the reference signatures; ``pyqg_params`` may additionally carry
``members`` (ensemble size on this GPU), ``member_offset``, ``device``, ``precision`` and ``seed``.
Snapshots are returned as a dict of numpy arrays with the variable names ``drop_vars(m.to_dataset())`` keeps (:16-60), a
leading ``run`` axis and float32 precision; ``tools/dataset.py`` writes them as NetCDF files in the reference's layout
(one ``<n>.nc`` per run) and ``to_xarray`` wraps them when xarray is importable.
"""
import json
import os

import numpy as np

from .dataset import model_coords, read_netcdf, write_forecast, write_runs
from .parameters import ANDREW_1000_STEPS, DAY
from .stochastic_pyqg import EnsembleQGModel, stochastic_QGModel


def initial_condition_fields(nx, L=1e6, members=1, rng=None):
    """Upper-layer PV of the JAMES-paper initial condition (:147-166) for ``members`` members, (members, nx, nx) float64.
    Members are drawn one after the other in the reference's order (rand(ny, nx), then rand(1, nx)), so a seeded ``rng``
    gives exactly what ``members`` successive reference calls would."""
    rng = np.random if rng is None else rng
    N = int(nx)
    dk = 2. * np.pi / L
    k, l = np.meshgrid(dk * np.arange(0., N // 2 + 1), dk * np.append(np.arange(0., N / 2), np.arange(-N / 2, 0.)))
    keep = np.sqrt(k ** 2 + l ** 2) < np.pi / (L / 32)
    out = np.empty((int(members), N, N))
    for b in range(int(members)):
        q2d = 1e-7 * rng.rand(N, N)
        q2d -= q2d.mean(axis=(-2, -1), keepdims=True)
        q2d *= np.sqrt(N * N / 64 ** 2)
        q1d = 1e-6 * (np.ones((N, 1)) * rng.rand(1, N))
        q1d -= q1d.mean(axis=(-2, -1), keepdims=True)
        q1d *= np.sqrt(N / 64)
        out[b] = np.fft.irfftn(np.fft.rfftn(q1d + q2d) * keep)
    return out


def set_initial_condition(m, rng=None):
    """JAMES-paper initial condition (:147-168), drawn independently for every member; lower layer at rest."""
    noise = initial_condition_fields(m.nx, m.L, m.members, rng)
    m.set_q(np.stack([noise, np.zeros_like(noise)], axis=1))
    m._invert()


def snapshot(m, out=None):
    """The physical-space variables ``drop_vars(m.to_dataset())`` keeps (:16-36), float32, shape (run,lev,y,x).  ``out``: dict of
    preallocated (run,lev,y,x) float32 arrays the fields are downloaded into (``run_simulation`` hands out slices of the final
    (time,run,lev,y,x) arrays, so a long ensemble run is not copied a second time when the snapshots are joined)."""
    m._invert()
    if hasattr(m, 'real32') and not getattr(m, 'squeeze', False):      # converted on the device: half the PCIe bytes
        o = out or {}
        return dict(q=m.real32('q', o.get('q')), u=m.real32('u', o.get('u')), v=m.real32('v', o.get('v')), psi=m.real32('p', o.get('psi')),
                    time=np.float64(m.t / 86400.))
    return dict(q=np.asarray(m.q, 'float32'), u=np.asarray(m.u, 'float32'), v=np.asarray(m.v, 'float32'),
                psi=np.asarray(m.p, 'float32'), time=np.float64(m.t / 86400.))


SPECTRAL_VARS = ('KEspec', 'Ensspec', 'KEflux', 'APEflux', 'APEgenspec', 'KEfrictionspec', 'entspec', 'paramspec',
                 'paramspec_KEflux', 'paramspec_APEflux', 'ENSflux', 'ENSgenspec', 'ENSfrictionspec', 'Dissspec', 'ENSDissspec',
                 'ENSparamspec')


def concat_in_time(snaps):
    out = {k: np.stack([s[k] for s in snaps], axis=1 if np.ndim(snaps[0][k]) else 0) for k in ('q', 'u', 'v', 'psi')}
    out['time'] = np.array([s['time'] for s in snaps])
    return out


def to_xarray(d, attrs=None):
    """dict from ``run_simulation`` -> xarray.Dataset with dims (run,time,lev,y,x) if xarray is installed."""
    try:
        import xarray as xr
    except ImportError:
        return d
    ds = xr.Dataset({k: xr.DataArray(d[k], dims=['run', 'time', 'lev', 'y', 'x']) for k in ('q', 'u', 'v', 'psi')})
    ds['time'] = xr.DataArray(d['time'], dims=['time'], attrs={'units': 'days'})
    for k in SPECTRAL_VARS:
        if k in d:
            ds[k] = xr.DataArray(d[k], dims=['lev', 'l', 'k'] if np.ndim(d[k]) == 3 else ['l', 'k'])
    ds.attrs.update(attrs or {})
    return ds


def run_simulation(pyqg_params, parameterization=None, q_init=None, sampling_freq=ANDREW_1000_STEPS, rng=None):
    """Reference :108-145.  ``parameterization`` = dict(self=closure, sampling='AR1'|'constant'|'deterministic', nsteps=int)."""
    pyqg_params = dict(pyqg_params)
    pyqg_params['tmax'] = float(pyqg_params['tmax'])
    if parameterization is None:
        m = EnsembleQGModel(**pyqg_params)
    else:
        params = dict(pyqg_params)
        params['parameterization'] = parameterization['self']
        m = stochastic_QGModel(params, parameterization['sampling'], parameterization['nsteps'])
        m.squeeze = False
    set_initial_condition(m, rng)
    snaps = []
    if q_init is not None:
        m.set_q(np.asarray(q_init, dtype='float64'))
        m._invert()
        snaps.append(snapshot(m))
    # the number of snapshots is known in advance: download every snapshot straight into its slot of (time,run,lev,y,x) arrays and
    # hand out (run,time,lev,y,x) views of them (np.stack of per-snapshot arrays copied a 1024-member run a second time)
    tsnapints = int(np.ceil(sampling_freq / m.dt))
    n_expected = len(snaps) + int(np.ceil((m.tmax - m.t) / m.dt)) // tsnapints
    store = None
    if n_expected > 0 and not getattr(m, 'squeeze', False) and hasattr(m, 'real32'):
        store = {k: np.empty((n_expected, m.members, 2, m.ny, m.nx), dtype=np.float32) for k in ('q', 'u', 'v', 'psi')}
        for i, sn in enumerate(snaps):
            for k in store:
                store[k][i] = sn[k]
    for t in m.run_with_snapshots(tsnapint=sampling_freq):
        i = len(snaps)
        slot = {k: a[i] for k, a in store.items()} if store is not None and i < n_expected else None
        snaps.append(snapshot(m, slot))
    if store is not None and len(snaps) == n_expected:
        ds = {k: np.moveaxis(a, 0, 1) for k, a in store.items()}
        ds['time'] = np.array([sn['time'] for sn in snaps])
    else:
        ds = concat_in_time(snaps)
    ds.update(m.averaged_diagnostics())      # KEspec, Ensspec, KEflux, APEflux, APEgenspec, KEfrictionspec, paramspec*, entspec
    coords, attrs = model_coords(m)
    ds['coords'] = coords
    ds['attrs'] = dict(attrs, pyqg_params=str(pyqg_params))
    ds['model'] = m
    return ds


def run_forecast(pyqg_params, parameterization, q_init, n_ens, sampling_freq=DAY, rng=None):
    """Ensemble forecast of the reference's ``--forecast`` mode (:254-292): ``n_ens`` runs from the SAME initial condition
    ``q_init`` (2, ny, nx), differing only in the stochastic closure, sampled every ``sampling_freq`` (1 day).  The reference
    calls ``run_simulation`` n_ens times; here the runs are the members of one batched integration.  Returns the fields of
    run 0 and the ensemble means, like the file the reference writes: {q, u, v, psi, q_mean, ..., time}."""
    params = dict(pyqg_params, members=int(n_ens))
    q0 = np.broadcast_to(np.asarray(q_init, dtype='float64'), (int(n_ens), 2) + np.shape(q_init)[-2:])
    ds = run_simulation(params, parameterization, q0, sampling_freq, rng)
    out = {'time': ds['time'], 'attrs': ds['attrs'], 'coords': ds['coords']}
    for var in ('q', 'u', 'v', 'psi'):
        out[var] = ds[var][0]
        out[var + '_mean'] = ds[var].astype('float64').mean(axis=0).astype('float32')
    return out


def generate_subgrid_forcing(Nc, pyqg_params, sampling_freq=ANDREW_1000_STEPS, operators=None, dealias='3/2-rule', rng=None):
    """Reference :62-106: run a hi-res ensemble and coarse-grain every ``sampling_freq`` seconds.  Defaults follow the
    reference ([Operator2, Operator5], '3/2-rule', keys '<Operator>-<nc>-dealias'); the published datasets used
    ``operators=[Operator1, Operator2], dealias='none'`` (scripts/train_parameterizations.py:28), keys '<Operator>-<nc>'.
    Returns {key: dict(q_forcing_advection, q, u, v, psi, time)} with float32 arrays (run,time,lev,y,x)."""
    from . import operators as ops
    if operators is None:
        operators = [ops.Operator2, ops.Operator5]
    suffix = '' if dealias == 'none' else '-dealias'
    pyqg_params = dict(pyqg_params)
    pyqg_params['tmax'] = float(pyqg_params['tmax'])
    m = EnsembleQGModel(**pyqg_params)
    set_initial_condition(m, rng)
    out = {}
    phys = {k: pyqg_params[k] for k in ('rek', 'delta', 'beta', 'rd', 'H1', 'U1', 'U2', 'L', 'filterfac') if k in pyqg_params}
    for t in m.run_with_snapshots(tsnapint=sampling_freq):
        qdns = m.device_q()
        for op in operators:
            for nc in Nc:
                forcing, mf = ops.PV_subgrid_forcing(qdns, nc, op, phys, dealias, return_fields=True)
                rec = dict(q_forcing_advection=forcing, q=mf['q'], u=mf['u'], v=mf['v'], psi=mf['psi'])
                rec = {k: np.asarray(v, 'float32') for k, v in rec.items()}
                rec['time'] = m.t / 86400.
                out.setdefault('%s-%d%s' % (op.__name__, nc, suffix), []).append(rec)
    for key, recs in out.items():
        d = {k: np.stack([r[k] for r in recs], axis=1) for k in ('q_forcing_advection', 'q', 'u', 'v', 'psi')}
        d['time'] = np.array([r['time'] for r in recs])
        nc = d['q'].shape[-1]
        d['coords'] = dict(x=(np.arange(nc) + 0.5) * m.L / nc, y=(np.arange(nc) + 0.5) * m.L / nc, lev=np.array([1, 2], dtype=np.int32))
        d['attrs'] = {'pyqg_params': str(pyqg_params)}
        out[key] = d
    return out


def _load_model(folder='model'):
    from ..models.cgan_regression import CGANRegression
    from ..models.cvae_regression import CVAERegression
    from ..models.mean_var_model import MeanVarModel
    from ..models.ols_model import OLSModel
    classes = dict(CGANRegression=CGANRegression, CVAERegression=CVAERegression, MeanVarModel=MeanVarModel,
                   OLSModel=OLSModel)
    with open(os.path.join(folder, 'model_args.json')) as f:
        model_args = json.load(f)
    name = model_args.pop('model')
    if name not in classes:
        raise ValueError('model %s is not on the accelerated path' % name)
    model_args.setdefault('folder', folder)
    return classes[name](**model_args)


def main(argv=None):
    """CLI with the flags of the reference ``simulate.py`` (:175-189) that lie on the accelerated path."""
    import argparse
    import ast
    p = argparse.ArgumentParser()
    p.add_argument('--pyqg_params', type=str, default=str({}))
    p.add_argument('--ensemble_member', type=int, default=0)
    p.add_argument('--members', type=int, default=1, help='ensemble members integrated together on this GPU')
    p.add_argument('--forcing', type=str, default='no')
    p.add_argument('--sampling_freq', type=int, default=ANDREW_1000_STEPS)
    p.add_argument('--reference', type=str, default='no')
    p.add_argument('--parameterization', type=str, default='no')
    p.add_argument('--subfolder', type=str, default='')
    p.add_argument('--sampling', type=str, default='AR1')
    p.add_argument('--nsteps', type=int, default=1)
    p.add_argument('--model_weight', type=float, default=1.0)
    p.add_argument('--model_folder', type=str, default='model')
    p.add_argument('--precision', type=str, default='fp32')
    p.add_argument('--forecast', type=str, default='no')
    p.add_argument('--initial_condition', type=str, default='no',
                   help="dict(path=, selector=dict(run=, time=), operator='Operator1'|..., n_ens=, number=) like the reference")
    args = p.parse_args(argv)
    params = dict(ast.literal_eval(args.pyqg_params))
    params.setdefault('members', args.members)
    params.setdefault('member_offset', args.ensemble_member)
    if args.subfolder:
        os.makedirs(args.subfolder, exist_ok=True)

    def save_runs(ds, folder):          # <ensemble_member + i>.nc like the reference (:249-263)
        write_runs(ds, folder or '.', first=args.ensemble_member)

    if args.forcing == 'yes':
        out = generate_subgrid_forcing([32, 48, 64, 96, 128], params, args.sampling_freq)
        for key, ds in out.items():         # <Operator>-<nc>[-dealias]/<member>.nc, what xr.open_mfdataset('<key>/*.nc') reads (:192-199)
            write_runs(ds, os.path.join(args.subfolder, key) if args.subfolder else key, first=args.ensemble_member)
    if args.reference == 'yes':
        save_runs(run_simulation(params, sampling_freq=args.sampling_freq), args.subfolder)
    if args.parameterization == 'yes':
        params['precision'] = args.precision
        model = args.model_weight * _load_model(args.model_folder)
        par = dict(self=model, sampling=args.sampling, nsteps=args.nsteps)
        save_runs(run_simulation(params, par, sampling_freq=args.sampling_freq), args.subfolder)


    if args.forecast == 'yes':
        from . import operators as ops
        ic = dict(ast.literal_eval(args.initial_condition))
        src = read_netcdf(ic['path'] + str(ic['selector']['run']) + '.nc')
        q_init = np.asarray(src['q'][ic['selector']['time']], dtype='float64')
        if ic.get('operator') in ('Operator1', 'Operator2', 'Operator5') and q_init.shape[-1] != params['nx']:
            q_init = getattr(ops, ic['operator'])(q_init, params['nx'])          # coarse-grain the hi-res snapshot (:274-278)
        par = None
        if os.path.exists(os.path.join(args.model_folder, 'model_args.json')):
            params['precision'] = args.precision
            par = dict(self=args.model_weight * _load_model(args.model_folder), sampling=args.sampling, nsteps=args.nsteps)
        params.pop('members', None)
        fc = run_forecast(params, par, q_init, ic['n_ens'], 1 * DAY)
        write_forecast(fc, os.path.join(args.subfolder or '.', '%s.nc' % ic['number']))


if __name__ == '__main__':
    main()
