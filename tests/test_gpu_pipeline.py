"""The reference's whole workflow on this engine, end to end, at toy sizes: a hi-res ensemble is integrated and coarse-grained into a
forcing dataset (``generate_subgrid_forcing``, tools/simulate.py:62-106), a CGAN closure is fitted to it (``CGANRegression.fit``,
models/cgan_regression.py:66-107), the saved folder is loaded the way the CLI does (``_load_model``, tools/simulate.py:236-244) and
drives a coarse stochastic ensemble online (``run_simulation`` :108-145).  No numbers from the reference here -- every stage has its own
parity test; this one holds the hand-overs between them (array layouts, file formats, scalers, precision modes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_forcing_dataset_to_trained_closure_to_online_run(tmp_path):
    from pyqg_generative_b200.models.cgan_regression import CGANRegression
    from pyqg_generative_b200.tools import operators as ops
    from pyqg_generative_b200.tools.simulate import _load_model, generate_subgrid_forcing, run_simulation
    DAY = 86400.0
    # 1. forcing dataset: 6 hi-res members at 128^2, 40 days spin-up data sampled every 5 days, Operator2 to 32^2
    hires = dict(nx=128, dt=7200.0, tmax=40 * DAY, tavestart=1e12, members=6, log_level=0)
    data = generate_subgrid_forcing([32], hires, sampling_freq=5 * DAY, operators=[ops.Operator2], dealias='none',
                                    rng=np.random.RandomState(0))
    ds = data['Operator2-32']
    assert ds['q'].shape == (6, 8, 2, 32, 32) and ds['q_forcing_advection'].shape == ds['q'].shape
    assert np.isfinite(ds['q_forcing_advection']).all() and ds['q_forcing_advection'].std() > 0
    train = {k: ds[k][:4] for k in ('q', 'q_forcing_advection')}
    test = {k: ds[k][4:] for k in ('q', 'q_forcing_advection')}
    # 2. fit the shipped generator / discriminator architectures (the tensor-core path serves this generator layout)
    folder = str(tmp_path / 'gan')
    model = CGANRegression(folder=folder, nx=32)
    np.random.seed(0)
    model.fit(train, test, num_epochs=2, batch_size=8, learning_rate=2e-4, nruns=2)
    # 3. the saved folder loads like the CLI loads it, and drives an online ensemble in fp32 and in tensor-core precision
    loaded = _load_model(folder)
    assert type(loaded).__name__ == 'CGANRegression' and loaded.hidden_channels == [128, 64, 32, 32, 32, 32, 32]
    lowres = dict(nx=32, dt=14400.0, tmax=20 * DAY, tavestart=1e12, members=4, log_level=0)
    states = {}
    for prec in ('fp32', 'tc'):
        loaded = _load_model(folder)
        out = run_simulation(dict(lowres, precision=prec), dict(self=loaded, sampling='AR1', nsteps=1), sampling_freq=10 * DAY,
                             rng=np.random.RandomState(1))
        assert out['q'].shape == (4, 2, 2, 32, 32) and np.isfinite(out['q']).all()
        m = out['model']
        assert m.tc == 120 and np.isfinite(m.PV_forcing).all() and np.abs(m.PV_forcing).max() > 0
        states[prec] = out['q']
    # same seeds, same noise: the two precisions track each other over the 120 coupled steps
    assert np.abs(states['tc'] - states['fp32']).max() < 1e-2 * np.abs(states['fp32']).max()


def test_train_model_cli_reads_the_forcing_files_the_simulate_cli_writes(tmp_path):
    """tools/train_model.py (reference :11-54) on files in the layout ``simulate.py --forcing yes`` writes (one <member>.nc per run)."""
    from pyqg_generative_b200.tools import train_model
    from pyqg_generative_b200.tools.dataset import write_runs
    from pyqg_generative_b200.models.ols_model import OLSModel
    rng = np.random.RandomState(3)
    q = (rng.randn(5, 3, 2, 16, 16) * np.array([7e-6, 1e-6])[None, None, :, None, None]).astype('float32')
    ds = dict(q=q, u=q, v=q, psi=q, q_forcing_advection=(1e-6 * (np.roll(q, 1, axis=-1) - q)).astype('float32'),
              time=np.arange(3.0), coords=dict(x=np.arange(16.0), y=np.arange(16.0), lev=np.array([1, 2], dtype=np.int32)), attrs={})
    folder = str(tmp_path / 'Operator2-16')
    paths = write_runs(ds, folder)
    assert len(paths) == 5
    runs = train_model.load_runs(folder + '/*.nc')
    assert runs['q'].shape == (5, 3, 2, 16, 16) and np.array_equal(runs['q_forcing_advection'], ds['q_forcing_advection'])
    out = str(tmp_path / 'model')
    model = train_model.main(['--model', 'OLSModel', '--model_args', "dict(folder=%r, hidden_channels=[16, 8])" % out,
                              '--fit_args', 'dict(num_epochs=2, batch_size=4)', '--nruns', '4', '--train_path', folder + '/*.nc'])
    assert isinstance(model, OLSModel) and len(model.net.log_dict['loss']) == 2
    for f in ('net.pt', 'x_scale.json', 'y_scale.json', 'model_args.json', 'stats.nc'):
        assert (tmp_path / 'model' / f).exists(), f
    with pytest.raises(ValueError, match='not on the accelerated path'):
        train_model.main(['--model', 'CVAEBottleneck', '--train_path', folder + '/*.nc'])
