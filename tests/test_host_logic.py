"""Host-side logic that needs no GPU: configuration tables, samplers, weight packing (BatchNorm folding), sharding."""
import os
import numpy as np
import pytest
import torch

from conftest import golden, golden_state_dict
from oracle import cnn_ref


def test_parameters_match_reference_tables():
    from pyqg_generative_b200.tools.parameters import EDDY_PARAMS, JET_PARAMS, ANDREW_1000_STEPS, YEAR
    assert EDDY_PARAMS.nx(256)['dt'] == 3600 and EDDY_PARAMS.nx(96)['dt'] == 7200 and EDDY_PARAMS.nx(48)['dt'] == 14400
    assert EDDY_PARAMS['nx'] == 64 and 'dt' in EDDY_PARAMS and EDDY_PARAMS['tmax'] == 10 * YEAR
    assert JET_PARAMS['rek'] == 7e-08 and JET_PARAMS['delta'] == 0.1 and JET_PARAMS['beta'] == 1e-11
    assert ANDREW_1000_STEPS == 3600000
    p = EDDY_PARAMS.nx(48)._update({'tmax': 1.0})
    assert p['tmax'] == 1.0 and EDDY_PARAMS['tmax'] != 1.0


def test_samplers_reproduce_reference_sequences():
    from pyqg_generative_b200.tools.stochastic_pyqg import AR1_sampler, constant_sampler
    g = golden('samplers.npz')
    for n in (1, 4, -1):
        s, xi = AR1_sampler(n), list(g['ar1_%d_xi' % n])
        for i in range(6):
            assert s.update(lambda: xi[i]) is True
            assert np.array_equal(s.noise, g['ar1_%d' % n][i])
    for n in (1, 3):
        s, rng = constant_sampler(n), np.random.RandomState(9)
        flags = []
        for i in range(8):
            flags.append(s.update(lambda: rng.randn(3)))
            assert np.array_equal(s.noise, g['const_%d' % n][i])
        assert flags == list(g['const_%d_flags' % n])


def test_unknown_sampling_type_raises_like_reference():
    from pyqg_generative_b200.tools.stochastic_pyqg import stochastic_QGModel
    with pytest.raises(ValueError, match='Unknown sampling type'):
        stochastic_QGModel(dict(nx=64), 'bogus', 1)


def test_state_dict_loading_and_batchnorm_folding():
    from pyqg_generative_b200.tools.cnn_tools import AndrewCNN
    sd, _, _ = golden_state_dict('weights_gan.npz')
    net = AndrewCNN(4, 2)
    assert net.load_state_dict(sd) == '<All keys matched successfully>'
    assert sorted(net.state_dict()) == sorted(sd)
    layers = net.layers()
    assert [(L['cin'], L['cout'], L['ksize']) for L in layers] == \
        [(4, 128, 5), (128, 64, 5), (64, 32, 3), (32, 32, 3), (32, 32, 3), (32, 32, 3), (32, 32, 3), (32, 2, 3)]
    # the folded per-layer arithmetic reproduces the oracle network on CPU (pure torch check of the packing)
    x = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(0))
    y = x
    import torch.nn.functional as F
    for L in layers:
        p = L['ksize'] // 2
        y = F.conv2d(F.pad(y, (p, p, p, p), mode='circular'), torch.as_tensor(L['weight']), torch.as_tensor(L['bias']))
        if L['relu_bn']:
            y = F.relu(y) * torch.as_tensor(L['bn_scale'])[None, :, None, None] + torch.as_tensor(L['bn_shift'])[None, :, None, None]
    ref = cnn_ref.andrew_cnn_forward(sd, x)
    assert (y - ref).abs().max() <= 2e-5 * ref.abs().max()
    with pytest.raises(RuntimeError):
        net.load_state_dict({'conv.0.weight': torch.zeros(1)})
    with pytest.raises(NotImplementedError):
        AndrewCNN(4, 2, div=True)


def test_scaler_reads_reference_json(tmp_path):
    from conftest import write_model_folder
    from pyqg_generative_b200.tools.cnn_tools import ChannelwiseScaler
    folder = write_model_folder(tmp_path, 'gan')
    s = ChannelwiseScaler().read('x_scale.json', folder)
    assert s.std.shape == (1, 2, 1, 1) and s.std.dtype == np.float32
    assert np.allclose(s.std.ravel(), [7.784383e-06, 1.0471941e-06], rtol=1e-6)
    X = np.ones((3, 2, 4, 4), 'float32')
    assert np.allclose(s.denormalize(s.normalize(X)), X)
    s.write('copy.json', folder)
    assert np.array_equal(ChannelwiseScaler().read('copy.json', folder).std, s.std)


def test_pyqg_parameterization_algebra():
    from pyqg_generative_b200.models.parameterization import QParameterization, WeightedParameterization

    class P(QParameterization):
        def __call__(self, m):
            return np.ones((2, 4, 4))
    w = 0.5 * P()
    assert isinstance(w, WeightedParameterization) and w.parameterization_type == 'q_parameterization'
    assert np.all(w(None) == 0.5) and np.all((P() + w)(None) == 1.5)


def test_generator_not_implemented_raises_like_reference(tmp_path):
    from pyqg_generative_b200.models.cgan_regression import CGANRegression
    with pytest.raises(ValueError, match='generator not implemented'):
        CGANRegression(generator='DeepInversion', folder=str(tmp_path))


def test_member_sharding_is_a_partition():
    from pyqg_generative_b200.parallel import shard_members
    for total in (1, 7, 64, 1024):
        for world in (1, 2, 3, 8):
            blocks = [shard_members(total, r, world) for r in range(world)]
            assert sum(c for c, _ in blocks) == total
            assert all(blocks[r][1] + blocks[r][0] == blocks[r + 1][1] for r in range(world - 1))
            assert max(c for c, _ in blocks) - min(c for c, _ in blocks) <= 1


def test_calc_ispec_matches_reference_golden():
    """tools/spectral_tools.calc_ispec (and the array-argument adapter parallel.calc_ispec) against the outputs of the
    unmodified reference function (pyqg_generative/tools/spectral_tools.py:103-180; tests/golden/make_golden.py
    ispec_fixture) for every option combination: bit-for-bit."""
    from pyqg_generative_b200 import parallel
    from pyqg_generative_b200.tools.spectral_tools import calc_ispec
    from oracle import pyqg_shim
    g = golden('ispec.npz')
    for n in (48, 64):
        m = pyqg_shim.QGModel(nx=n, log_level=0)
        spec = g['spec_%d' % n]
        for avg in (True, False):
            for trunc in (True, False):
                for nd in (False, True):
                    for nf in (1, 2):
                        tag = '%d_%d%d%d%d' % (n, avg, trunc, nd, nf)
                        kr, ph = calc_ispec(m, spec, averaging=avg, truncate=trunc, nd_wavenumber=nd, nfactor=nf)
                        assert np.array_equal(kr, g['kr_' + tag]) and np.array_equal(ph, g['ph_' + tag]), tag
                        kr2, ph2 = parallel.calc_ispec(m.k, m.l, spec, averaging=avg, truncate=trunc, nd_wavenumber=nd, nfactor=nf)
                        assert np.array_equal(kr2, kr) and np.array_equal(ph2, ph), tag
    # Parseval in summation mode (the normalisation the reference documents): signal.var() == phr.sum() * dkr
    m = pyqg_shim.QGModel(nx=64, log_level=0)
    x = np.random.RandomState(3).randn(64, 64)
    xh = np.fft.rfftn(x) * (m.wv < 30 * m.dk) * (m.wv >= m.dk)     # inside the truncation circle, no mean
    x = np.fft.irfftn(xh)
    kr, ph = calc_ispec(m, np.abs(xh) ** 2 / m.M ** 2, averaging=False, truncate=True)
    assert abs(ph.sum() * (kr[1] - kr[0]) - x.var()) < 1e-12 * x.var()


def test_initial_condition_matches_reference_golden():
    """tools/simulate.initial_condition_fields against the unmodified reference set_initial_condition
    (pyqg_generative/tools/simulate.py:147-168) under the same seeded host RNG: two successive members per grid size."""
    from pyqg_generative_b200.tools.simulate import initial_condition_fields
    g = golden('initial_condition.npz')
    for n in (48, 64, 96):
        np.random.seed(1000 + n)
        q1 = initial_condition_fields(n, 1e6, members=2)
        ref = g['q_%d' % n]                     # (member, lev, y, x)
        assert np.array_equal(ref[:, 1], np.zeros_like(ref[:, 1]))
        assert np.abs(q1 - ref[:, 0]).max() <= 1e-15 * np.abs(ref[:, 0]).max(), n
        rs = np.random.RandomState(1000 + n)     # an explicit generator gives the same stream
        assert np.array_equal(initial_condition_fields(n, 1e6, members=2, rng=rs), q1)


def test_netcdf_writer_layout_matches_reference_files(tmp_path):
    """tools/dataset.py: one <n>.nc per run with the variables / dims / dtypes / units that
    drop_vars + concat_in_time produce (pyqg_generative/tools/simulate.py:16-60,138-145)."""
    from pyqg_generative_b200.tools import dataset
    rng = np.random.RandomState(0)
    R, T, N = 3, 4, 16
    ds = {k: rng.randn(R, T, 2, N, N).astype('float32') for k in dataset.PHYSICAL}
    ds['time'] = np.arange(1, T + 1) * 41.6666666667
    ds['KEspec'] = rng.rand(2, N, N // 2 + 1)
    ds['KEflux'] = rng.randn(N, N // 2 + 1)
    ds['paramspec'] = rng.randn(N, N // 2 + 1)
    ds['coords'] = dict(x=(np.arange(N) + .5) * 1e6 / N, y=(np.arange(N) + .5) * 1e6 / N, lev=np.array([1, 2]),
                        l=np.fft.fftfreq(N, 1. / N) * 2 * np.pi / 1e6, k=np.arange(N // 2 + 1) * 2 * np.pi / 1e6,
                        Ubg=np.array([0.025, 0.]), Qy=np.array([1e-10, 2e-11]))
    ds['attrs'] = {'pyqg_params': str(dict(nx=N, dt=3600.)), 'pyqg:beta': 1.5e-11}
    paths = dataset.write_runs(ds, str(tmp_path / 'runs'), first=5)
    assert [os.path.basename(p) for p in paths] == ['5.nc', '6.nc', '7.nc']
    d = dataset.read_netcdf(paths[1])
    assert d['dims'] == {'time': T, 'lev': 2, 'y': N, 'x': N, 'l': N, 'k': N // 2 + 1}
    for k in dataset.PHYSICAL:
        assert d[k].dtype == np.float32 and d['var_dims'][k] == ('time', 'lev', 'y', 'x')
        assert np.array_equal(d[k], ds[k][1])
    assert d['var_attrs']['time']['units'] == b'days' and np.allclose(d['time'], ds['time'])
    assert d['var_dims']['KEspec'] == ('lev', 'l', 'k') and d['var_dims']['KEflux'] == ('l', 'k')
    assert np.allclose(d['KEspec'], ds['KEspec'].astype('float32')) and d['KEspec'].dtype == np.float32
    assert d['attrs']['pyqg_params'] == str(dict(nx=N, dt=3600.)) and abs(d['attrs']['pyqg:beta'] - 1.5e-11) < 1e-25
    assert np.allclose(d['Ubg'], [0.025, 0.]) and np.allclose(d['x'], ds['coords']['x'])
    one = dataset.read_netcdf(dataset.write_netcdf(ds, str(tmp_path / 'all.nc')))
    assert one['var_dims']['q'] == ('run', 'time', 'lev', 'y', 'x') and np.array_equal(one['q'], ds['q'])


def test_training_logs_have_the_layout_of_the_shipped_stats_files(tmp_path):
    """write_log / loss_to_log against the layout of Google-Colab/{GAN,VAE}/stats.nc and GZ/stats_var.nc (tests/golden/stats_layout.json,
    read from the reference's files by make_golden.py): same container, variables, dimensions, dtypes and coordinates."""
    import json
    from scipy.io import netcdf_file
    from conftest import GOLDEN
    from pyqg_generative_b200.tools.cnn_tools import write_log
    from pyqg_generative_b200.models.cvae_regression import loss_to_log
    layout = json.load(open(os.path.join(GOLDEN, 'stats_layout.json')))
    scores = [dict(L2_mean=1.0, L2_total=0.5 - 0.1 * i, L2_residual=0.4, var_ratio=[0.9, 0.8]) for i in range(3)]
    cases = {
        'VAE': loss_to_log({k: [1., 2., 3.] for k in ('loss', 'loss_KL', 'var_aggr', 'MSE', 'loss_recon', 'var_latent')}, scores, scores)[0],
        'GAN': loss_to_log({k: [1., 2., 3.] for k in ('D_loss', 'D_drift', 'D_grad', 'G_loss')}, scores, scores, name='loss')[0],
        'GZ_var': {'loss': [1., 2., 3.], 'loss_test': [2., 3., 4.]},
    }
    for tag, log in cases.items():
        path = str(tmp_path / (tag + '.nc'))
        write_log(log, path)
        assert list(open(path, 'rb').read(4)) == layout[tag]['magic'], tag
        with netcdf_file(path, 'r', mmap=False) as nc:
            ours = {k: dict(dims=list(v.dimensions), dtype=v.data.dtype.str) for k, v in nc.variables.items()}
            assert ours == layout[tag]['variables'], (tag, ours)
            assert int(nc.variables['epoch'][0]) == layout[tag]['epoch_first']
            if layout[tag]['lev']:
                assert [int(x) for x in nc.variables['lev'][:]] == layout[tag]['lev']
    assert cases['VAE']['Epoch_opt'] == 3.0
