"""run_simulation -> NetCDF files in the reference's layout on the GPU engine."""
import numpy as np
import pytest

from oracle import pyqg_shim

pytestmark = pytest.mark.gpu


def test_run_simulation_writes_reference_style_files(tmp_path):
    from pyqg_generative_b200.tools import dataset, simulate
    N, dt = 32, 14400.
    params = dict(nx=N, dt=dt, tmax=24 * dt, tavestart=8 * dt, taveint=4 * dt, members=3, log_level=0)
    ds = simulate.run_simulation(params, sampling_freq=8 * dt, rng=np.random.RandomState(1))
    assert ds['q'].shape == (3, 3, 2, N, N) and ds['q'].dtype == np.float32
    assert np.allclose(ds['time'], np.array([8, 16, 24]) * dt / 86400.)
    # time-averaged spectral diagnostics come with the dataset (pyqg to_dataset + concat_in_time: from the last snapshot)
    for k in ('KEspec', 'Ensspec'):
        assert ds[k].shape == (2, N, N // 2 + 1)
    for k in ('KEflux', 'APEflux', 'APEgenspec', 'KEfrictionspec', 'entspec', 'paramspec', 'ENSflux', 'ENSgenspec',
              'ENSfrictionspec', 'Dissspec', 'ENSDissspec', 'ENSparamspec'):
        assert ds[k].shape == (N, N // 2 + 1)
    assert ds['EKE'].shape == (2,) and ds['EKE'][0] > ds['EKE'][1] > 0 and ds['EKEdiss'] > 0
    paths = dataset.write_runs(ds, str(tmp_path / 'eddy'), first=0)
    d = dataset.read_netcdf(paths[2])
    assert np.array_equal(d['q'], ds['q'][2]) and np.array_equal(d['psi'], ds['psi'][2])
    assert d['var_dims']['KEspec'] == ('lev', 'l', 'k') and np.allclose(d['KEspec'], ds['KEspec'].astype('float32'))
    assert d['var_dims']['Dissspec'] == ('l', 'k') and np.allclose(d['Dissspec'], ds['Dissspec'].astype('float32'))
    assert d['var_dims']['EKE'] == ('lev',) and np.allclose(d['EKE'], ds['EKE'].astype('float32'))
    assert np.isclose(d['attrs']['EKEdiss'], ds['EKEdiss'], rtol=1e-6)
    o = pyqg_shim.QGModel(nx=N, dt=dt, log_level=0)
    assert np.allclose(d['k'], o.kk) and np.allclose(d['l'], o.ll) and np.allclose(d['Qy'], o.Qy.astype('float32'))
    assert d['attrs']['pyqg_params'] == str(dict(params, tmax=float(params['tmax'])))
    # last snapshot: velocities are the inversion of the stored q (float32 round trip of the fp64 fields)
    o.q = ds['model'].q[2]
    o._invert()
    assert np.abs(d['u'][-1] - o.u).max() < 1e-6 * np.abs(o.u).max()


def test_cli_writes_one_file_per_member(tmp_path):
    from pyqg_generative_b200.tools import dataset, simulate
    sub = str(tmp_path / 'ref')
    simulate.main(['--reference=yes', '--members=2', '--ensemble_member=4', '--subfolder=' + sub, '--sampling_freq=57600',
                   "--pyqg_params={'nx': 32, 'dt': 14400.0, 'tmax': 115200.0, 'log_level': 0}"])
    import os
    assert sorted(os.listdir(sub)) == ['4.nc', '5.nc']
    d = dataset.read_netcdf(os.path.join(sub, '5.nc'))
    assert d['q'].shape == (2, 2, 32, 32) and np.isfinite(d['q']).all()


def test_forecast_mode_runs_the_ensemble_in_one_batch(tmp_path):
    """tools/simulate.py --forecast (reference :254-292): members share the initial condition read from a run file."""
    import os
    from pyqg_generative_b200.tools import dataset, simulate
    N, dt = 32, 14400.
    ref = simulate.run_simulation(dict(nx=N, dt=dt, tmax=12 * dt, members=2, log_level=0), sampling_freq=6 * dt,
                                  rng=np.random.RandomState(2))
    dataset.write_runs(ref, str(tmp_path / 'ref'), first=0)
    ic = dict(path=str(tmp_path / 'ref') + '/', selector=dict(run=1, time=-1), operator='Operator1', n_ens=3, number=7)
    simulate.main(['--forecast=yes', '--subfolder=' + str(tmp_path / 'fc'), '--model_folder=' + str(tmp_path / 'nomodel'),
                   '--initial_condition=' + str(ic), "--pyqg_params={'nx': 32, 'dt': 14400.0, 'tmax': 172800.0, 'log_level': 0}"])
    d = dataset.read_netcdf(os.path.join(str(tmp_path / 'fc'), '7.nc'))
    assert d['q'].shape == (3, 2, N, N) and np.allclose(d['time'], [0., 1., 2.])            # IC + 2 daily snapshots
    assert np.array_equal(d['q'][0], ref['q'][1, -1])                                        # starts from the selected snapshot
    # without a stochastic closure every member follows the same trajectory: run 0 equals the ensemble mean
    assert np.abs(d['q'] - d['q_mean']).max() <= 1e-6 * np.abs(d['q']).max()
    o = pyqg_shim.QGModel(nx=N, dt=dt, log_level=0)
    o.q = ref['q'][1, -1].astype('float64')
    for _ in range(6):
        o._step_forward()
    assert np.abs(d['q'][1] - o.q).max() < 1e-6 * np.abs(o.q).max()
