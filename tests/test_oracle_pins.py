"""The oracle against the reference: committed golden vectors (produced by the UNMODIFIED reference, see
tests/golden/make_golden.py) and the known-answer material recorded in the reference notebooks."""
import numpy as np
import pytest
import torch

from conftest import golden, golden_state_dict
from oracle import cnn_ref, operators_ref as opr, pyqg_shim


@pytest.mark.parametrize('kind,files,zkey,key', [
    ('gan', ['weights_gan.npz'], 'z32', 'gan_snapshot'),
    ('vae', ['weights_vae.npz'], 'z32', 'vae_snapshot'),
    ('gz', ['weights_gz_mean.npz', 'weights_gz_var.npz'], 'z64', 'gz_snapshot')])
def test_predict_snapshot_restatement_matches_reference(kind, files, zkey, key):
    c = golden('closure_48.npz')
    nets = [golden_state_dict(f)[0] for f in files]
    _, xs, ys = golden_state_dict(files[0])
    y = cnn_ref.predict_snapshot(kind, nets, xs, ys, c['q'].astype('float64'), c[zkey])
    assert np.abs(y - c[key]).max() <= 1e-6 * np.abs(c[key]).max()


def test_generate_restatement_matches_reference():
    c = golden('closure_48.npz')
    sd, _, _ = golden_state_dict('weights_gan.npz')
    x = torch.cat([torch.as_tensor(c['generate_x']), torch.as_tensor(c['generate_z'])], dim=1)
    y = cnn_ref.andrew_cnn_forward(sd, x).numpy()
    assert np.abs(y - c['gan_generate']).max() <= 1e-5 * np.abs(c['gan_generate']).max()


def test_call_demeans_and_uses_sampler_noise():
    c = golden('closure_48.npz')
    sd, xs, ys = golden_state_dict('weights_vae.npz')
    y = cnn_ref.predict_snapshot('vae', [sd], xs, ys, c['q'].astype('float64'), c['vae_call_noise'])
    assert np.abs(cnn_ref.demean(y) - c['vae_call']).max() <= 1e-6 * np.abs(c['vae_call']).max()
    assert np.abs(c['vae_call'].mean(axis=(1, 2))).max() < 1e-20


def test_operator_restatements_match_reference():
    o = golden('operators_128.npz')
    q = o['q'].astype('float64')
    for nc in (32, 48, 64):
        for name in ('Operator1', 'Operator2', 'Operator5', 'cut_off'):
            ref = o['%s_%d' % (name, nc)]
            assert np.abs(getattr(opr, name)(q, nc) - ref).max() <= 1e-14 * np.abs(ref).max(), (name, nc)
    assert np.abs(opr.fft_interpolate(o['interp_in'], 48, 72) - o['interp_48_72']).max() < 1e-13
    assert np.abs(opr.fft_interpolate(o['interp_in'], 48, 32) - o['interp_48_32']).max() < 1e-13
    for opn in ('Operator1', 'Operator2', 'Operator5'):
        for de, tag in (('none', 'none'), ('3/2-rule', '32')):
            f, mf, m = opr.PV_subgrid_forcing(q, 64, getattr(opr, opn), {}, de)
            ref = o['S_%s_%s' % (opn, tag)]
            assert np.abs(f - ref).max() <= 1e-12 * np.abs(ref).max(), (opn, tag)


def test_operator4_and_two_thirds_rule_restatements_match_reference():
    q = golden('operators_128.npz')['q'].astype('float64')
    o = golden('operators_128_more.npz')
    for nc in (32, 48, 64):
        ref = o['Operator4_%d' % nc]
        assert np.abs(opr.Operator4(q, nc) - ref).max() <= 1e-14 * np.abs(ref).max(), nc
    for opn in ('Operator1', 'Operator2', 'Operator4', 'Operator5'):
        f, mf, m = opr.PV_subgrid_forcing(q, 64, getattr(opr, opn), {}, '2/3-rule')
        ref = o['S_%s_23' % opn]
        assert np.abs(f - ref).max() <= 1e-12 * np.abs(ref).max(), opn
    for de, tag in (('none', 'none'), ('3/2-rule', '32')):
        f, mf, m = opr.PV_subgrid_forcing(q, 64, opr.Operator4, {}, de)
        ref = o['S_Operator4_%s' % tag]
        assert np.abs(f - ref).max() <= 1e-12 * np.abs(ref).max(), tag


def test_samplers_match_reference():
    g = golden('samplers.npz')
    for n in (1, 4, -1):
        s = cnn_ref.AR1Sampler(n)
        xi = list(g['ar1_%d_xi' % n])
        for i in range(6):
            s.update(lambda: xi[i])
            assert np.array_equal(s.noise, g['ar1_%d' % n][i])
    for n in (1, 3):
        s = cnn_ref.ConstantSampler(n)
        rng = np.random.RandomState(9)
        flags = [s.update(lambda: rng.randn(3)) for _ in range(8)]
        assert flags == list(g['const_%d_flags' % n])


# ---- known answers recorded in the reference notebooks -------------------------------------------
def test_q_setter_roundtrip_is_exact():
    # notebooks/3-2-dealiasing.ipynb:88 : m.q = q; m._invert(); ||q - m.q|| = 0.0
    m = pyqg_shim.QGModel(nx=64, log_level=0)
    q = np.random.RandomState(0).randn(2, 64, 64)
    m.q = q
    m._invert()
    assert np.abs(q - m.q).max() == 0.0


def test_cut_off_equals_fft_interpolate():
    # notebooks/3-2-dealiasing.ipynb:586 : ||cut_off(x,16) - fft_interpolate(x,64,16)|| = 0.0
    x = np.random.RandomState(1).randn(64, 64)
    assert np.abs(opr.cut_off(x, 16) - opr.fft_interpolate(x, 64, 16)).max() < 1e-15


def test_interpolation_of_a_resolved_wave_is_exact():
    # notebooks/3-2-dealiasing.ipynb:486 : error 7.8e-15 for cos(x) sin(y) interpolated 16 -> 24
    def field(n):
        x = 2 * np.pi * (np.arange(n) + 0.5) / n
        return np.cos(x)[None, :] * np.sin(x)[:, None]
    # cell-centred grids of different n are shifted relative to each other; use node-centred sampling instead
    def field0(n):
        x = 2 * np.pi * np.arange(n) / n
        return np.cos(x)[None, :] * np.sin(x)[:, None]
    assert np.abs(opr.fft_interpolate(field0(16), 16, 24) - field0(24)).max() < 1e-14


def test_initial_cfl_matches_recorded_logs():
    # notebooks/3-2-dealiasing.ipynb:1431 (64^2, dt=14400: CFL 0.023); online-simulations.ipynb:318 (48^2, dt=7200: 0.009)
    for nx, dt, cfl in ((64, 14400., 0.023), (48, 7200., 0.009)):
        np.random.seed(0)
        m = pyqg_shim.QGModel(nx=nx, dt=dt, log_level=0)
        opr.set_initial_condition(m)
        assert abs(m._calc_cfl() - cfl) < 1e-3


def test_advection_conserves_mean_and_tendency_is_hermitian():
    # the Jacobian integrates to zero (notebook cells 11-13 print ~1e-17 conservation residuals)
    np.random.seed(1)
    m = pyqg_shim.QGModel(nx=64, log_level=0, beta=0.0, rek=0.0, U1=0.0)
    m.q = np.random.randn(2, 64, 64) * 1e-5
    m._invert()
    m._do_advection()
    assert np.abs(m.dqhdt[:, 0, 0]).max() < 1e-18 * 64 * 64


def test_growth_curve_matches_recorded_log():
    # notebooks/3-2-dealiasing.ipynb:1431-1433: KE grows ~x6 per 1000 steps in the linear phase, CFL stays 0.023
    np.random.seed(0)
    m = pyqg_shim.QGModel(nx=64, dt=14400., log_level=1, tmax=1e12)
    opr.set_initial_condition(m)
    for _ in range(2000):
        m._step_forward()
    (_, _, ke1, cfl1), (_, _, ke2, _) = m.log
    assert 2e-7 < ke1 < 4e-6 and abs(cfl1 - 0.023) < 2e-3
    assert 4.0 < ke2 / ke1 < 9.0


def _band_limited_model(nx=64, dt=10., param=None, seed=0):
    rng = np.random.RandomState(seed)
    m = pyqg_shim.QGModel(nx=nx, dt=dt, log_level=0, tavestart=1e20, q_parameterization=param)
    qh = m.fft(rng.randn(2, nx, nx) * np.array([7e-6, 1e-6])[:, None, None])
    qh[:, np.sqrt(m.wv2) / (2 * np.pi / m.L) > nx / 3 - 1] = 0          # quadratic terms alias-free
    m.q = m.ifft(qh)
    m._invert()
    return m


def _full_plane_sum(x):
    w = np.full(x.shape, 2.0)
    w[..., 0] = w[..., -1] = 1.0
    return (x * w).sum()


def test_flux_diagnostics_conserve_energy():
    """pyqg is absent, so the restated diagnostics (oracle/pyqg_shim.py diagnostic_fields) are pinned by identities the true
    definitions satisfy: the Jacobian terms only redistribute energy, sum_k KEflux = sum_k APEflux = 0."""
    m = _band_limited_model()
    d = m.diagnostic_fields()
    scale = _full_plane_sum(np.abs(d['KEflux']))
    assert abs(_full_plane_sum(d['KEflux'])) < 1e-12 * scale and abs(_full_plane_sum(d['APEflux'])) < 1e-12 * scale


def test_spectral_energy_budget_closes():
    """d/dt of the modal energy equals KEflux + APEflux + APEgenspec + KEfrictionspec + paramspec (signs and factors of
    every term), and paramspec splits exactly into its KE and APE parts."""
    rng = np.random.RandomState(3)
    dq = rng.randn(2, 64, 64) * np.array([1e-12, 2e-13])[:, None, None]
    dq -= dq.mean(axis=(1, 2), keepdims=True)

    class Par(pyqg_shim.QParameterization):
        def __call__(self, mm):
            return dq
    m = _band_limited_model(param=Par())

    def energy(mm):
        mm._invert()
        ph = mm.ph
        return 0.5 * (mm.del1 * mm.wv2 * np.abs(ph[0]) ** 2 + mm.del2 * mm.wv2 * np.abs(ph[1]) ** 2
                      + mm.rd ** -2 * mm.del1 * mm.del2 * np.abs(ph[0] - ph[1]) ** 2) / mm.M ** 2
    e0 = energy(m)
    m._do_advection(); m._do_friction(); m._do_q_subgrid_parameterization()
    d = m.diagnostic_fields()
    rhs = d['KEflux'] + d['APEflux'] + d['APEgenspec'] + d['KEfrictionspec'] + d['paramspec']
    m._forward_timestep()
    lhs = (energy(m) - e0) / m.dt
    assert np.abs(lhs - rhs).max() < 2e-4 * np.abs(rhs).max()
    assert np.abs(d['paramspec'] - d['paramspec_KEflux'] - d['paramspec_APEflux']).max() < 1e-13 * np.abs(d['paramspec']).max()
    # each term matters for the closure of the budget (guards against a silently dropped term)
    for k in ('KEflux', 'APEflux', 'APEgenspec', 'KEfrictionspec', 'paramspec'):
        assert np.abs(lhs - (rhs - d[k])).max() > 1e-3 * np.abs(rhs).max(), k


def _filtered_band_model(kmax, param=None, dt=10., seed=5, nx=64):
    """White noise truncated at |kappa| <= kmax (in units of 2 pi / L): kmax > 0.65 pi / dx puts content under the filter."""
    rng = np.random.RandomState(seed)
    m = pyqg_shim.QGModel(nx=nx, dt=dt, log_level=0, tavestart=1e20, q_parameterization=param)
    qh = m.fft(rng.randn(2, nx, nx) * np.array([7e-6, 1e-6])[:, None, None])
    qh[:, np.sqrt(m.wv2) / (2 * np.pi / m.L) > kmax] = 0
    m.q = m.ifft(qh)
    return m


def _forcing_param(seed=3):
    rng = np.random.RandomState(seed)
    dq = rng.randn(2, 64, 64) * np.array([1e-12, 2e-13])[:, None, None]
    dq -= dq.mean(axis=(1, 2), keepdims=True)

    class Par(pyqg_shim.QParameterization):
        def __call__(self, mm):
            return dq
    return Par()


def test_modal_enstrophy_budget_closes_exactly():
    """Enstrophy Z(k) = sum_z del_z |qh_z|^2 / (2 M^2) is quadratic, so Z(q + d) - Z(q) - Z(d) is the projection of the
    increment d on the state: it must equal dt (ENSflux + ENSgenspec + ENSfrictionspec + ENSparamspec + ENSDissspec) to rounding,
    with every term necessary -- signs, layer weights and the 1/dt of the dissipation spectrum included.  The state has content
    under the exponential filter and aliases freely: the identity is algebraic."""
    m = _filtered_band_model(26., param=_forcing_param())
    hr = np.array([m.del1, m.del2])[:, None, None]
    Z = lambda qh: 0.5 * (hr * np.abs(qh) ** 2).sum(axis=0) / m.M ** 2
    ok = m.filtr > 0.5
    for step in range(3):                                   # Euler, AB2, AB3 coefficients of the dissipation spectrum
        m._invert(); m._do_advection(); m._do_friction(); m._do_q_subgrid_parameterization()
        d = m.diagnostic_fields()
        qh0, ph0 = m.qh.copy(), m.ph.copy()
        m._forward_timestep()
        # what the filter removed, recovered from the new state alone: qh1 = filtr X  =>  D = qh1 - qh1 / filtr
        D = np.where(ok, m.qh - m.qh / np.where(ok, m.filtr, 1.0), 0.0)
        ens_d = (hr * np.real(np.conj(qh0) * D)).sum(axis=0) / m.dt / m.M ** 2
        e_d = -(hr * np.real(np.conj(ph0) * D)).sum(axis=0) / m.dt / m.M ** 2
        assert np.abs(d['ENSDissspec'] - ens_d)[ok].max() < 1e-9 * np.abs(ens_d).max() > 0, step
        assert np.abs(d['Dissspec'] - e_d)[ok].max() < 1e-9 * np.abs(e_d).max() > 0, step
        if step > 0:
            continue                                        # (a multistep increment mixes tendencies of earlier states)
        lhs = (Z(m.qh) - Z(qh0) - Z(m.qh - qh0)) / m.dt
        terms = ('ENSflux', 'ENSgenspec', 'ENSfrictionspec', 'ENSparamspec', 'ENSDissspec')
        rhs = sum(d[k] for k in terms)
        scale = np.abs(rhs).max()
        assert np.abs(lhs - rhs).max() < 1e-9 * scale
        for k in terms:                                     # every term is needed: dropping it leaves exactly that term
            resid = np.abs(lhs - (rhs - d[k]))
            assert np.abs(d[k]).max() > 0 and np.abs(resid - np.abs(d[k])).max() < 1e-9 * scale, k
        assert (m.filtr < 0.9).any() and np.abs(d['ENSDissspec']).max() > 0.1 * scale     # the filter really acted


def test_energy_budget_with_filter_dissipation():
    """Same construction for the modal energy: with alias-free quadratic terms (|kappa| <= nx / 3, which still reaches under the
    filter edge 0.65 pi / dx) the first-order change equals KEflux + APEflux + APEgenspec + KEfrictionspec + paramspec +
    Dissspec, and Dissspec carries the budget where the filter acts."""
    m = _filtered_band_model(64 / 3., param=_forcing_param())
    d1, d2, F = m.del1, m.del2, m.rd ** -2 * m.del1 * m.del2

    def energy(qh):
        ph = np.einsum('ij...,j...->i...', m.a, qh)
        return 0.5 * (d1 * m.wv2 * np.abs(ph[0]) ** 2 + d2 * m.wv2 * np.abs(ph[1]) ** 2 + F * np.abs(ph[0] - ph[1]) ** 2) / m.M ** 2
    m._invert(); m._do_advection(); m._do_friction(); m._do_q_subgrid_parameterization()
    d = m.diagnostic_fields()
    qh0 = m.qh.copy()
    m._forward_timestep()
    lhs = (energy(m.qh) - energy(qh0) - energy(m.qh - qh0)) / m.dt
    rhs = d['KEflux'] + d['APEflux'] + d['APEgenspec'] + d['KEfrictionspec'] + d['paramspec'] + d['Dissspec']
    assert np.abs(lhs - rhs).max() < 2e-4 * np.abs(rhs).max()
    ring = (m.filtr < 1.0) & (np.abs(qh0[0]) > 0)
    assert ring.any()
    assert np.abs(lhs - (rhs - d['Dissspec']))[ring].max() > 0.5 * np.abs(d['Dissspec'])[ring].max() > 0
    assert np.abs(lhs - rhs)[ring].max() < 1e-3 * np.abs(d['Dissspec'])[ring].max()


def test_eke_scalars_follow_from_kespec():
    """EKE = 0.5 (u^2 + v^2).mean() per layer and EKEdiss = del2 rek (u_2^2 + v_2^2).mean() are the half-plane sums of KEspec
    (Parseval): the engine derives them from the (averaged) KEspec instead of carrying two more accumulators."""
    m = _band_limited_model()
    d = m.diagnostic_fields()
    eke = 0.5 * np.array([_full_plane_sum(d['KEspec'][z]) for z in range(2)])
    assert np.abs(eke - d['EKE']).max() < 1e-12 * d['EKE'].max()
    assert abs(m.del2 * m.rek * 2 * eke[1] - d['EKEdiss']) < 1e-12 * d['EKEdiss']


def _golden_sd(g, prefix):
    return {k[len(prefix) + 1:]: g[k] for k in g.files if k.startswith(prefix + '/')}


def test_training_oracle_reproduces_the_reference():
    """oracle/train_ref.py against the unmodified reference's compute_loss / autograd / cnn_tools.train (training.npz)."""
    from oracle import train_ref
    g = golden('training.npz')
    for tag, target in (('mean', g['grad_y']), ('var', g['grad_y'] ** 2)):
        loss, grads, after = train_ref.loss_and_grads(_golden_sd(g, tag + '_init'), g['grad_x'], target, softplus=tag == 'var')
        assert abs(loss - float(g[tag + '_loss'])) < 1e-6 * abs(float(g[tag + '_loss']))
        for k, v in _golden_sd(g, tag + '_grad').items():
            assert np.abs(grads[k] - v).max() <= 1e-6 * max(np.abs(v).max(), 1e-30), (tag, k)
        for k, v in _golden_sd(g, tag + '_after').items():
            assert np.abs(after[k] - v).max() <= 1e-6 * np.abs(v).max(), (tag, k)
    np.random.seed(0)
    final, log = train_ref.train(_golden_sd(g, 'run_init'), g['X_train'], g['Y_train'], g['X_test'], g['Y_test'], 4, 8, 1e-3)
    assert np.allclose(log['loss'], g['run_loss'], rtol=1e-5) and np.allclose(log['loss_test'], g['run_loss_test'], rtol=1e-5)
    for k, v in _golden_sd(g, 'run_final').items():
        assert np.abs(final[k].astype('float64') - v).max() <= 1e-5 * max(np.abs(v).max(), 1e-30), k


def test_cvae_oracle_reproduces_the_reference():
    """oracle/train_ref.py:cvae_losses against CVAERegression.compute_loss + autograd of the unmodified reference."""
    import torch
    from oracle import train_ref
    g = golden('training_cvae.npz')
    keys = ('loss', 'loss_recon', 'loss_KL', 'MSE', 'var_latent', 'var_aggr')
    for tag, dv in (('adaptive', 'adaptive'), ('fixed01', 0.1)):
        enc, dec = train_ref.Net(_golden_sd(g, tag + '_enc_init')), train_ref.Net(_golden_sd(g, tag + '_dec_init'))
        enc.train(); dec.train()
        losses = train_ref.cvae_losses(enc, dec, torch.as_tensor(g['grad_x']), torch.as_tensor(g['grad_y']),
                                       torch.as_tensor(g[tag + '_eps']), dv)
        losses['loss'].backward()
        for k, r in zip(keys, g[tag + '_losses']):
            assert abs(float(losses[k]) - r) <= 1e-6 * abs(r), (tag, k)
        for name, net in (('enc', enc), ('dec', dec)):
            for k, p in net.named_parameters():
                v = g['%s_%s_grad/%s' % (tag, name, k)]
                assert np.abs(p.grad.numpy() - v).max() <= 1e-5 * max(np.abs(v).max(), 1e-30), (tag, name, k)


def test_cgan_oracle_reproduces_the_reference():
    """oracle/train_ref.py:cgan_iteration against the first iteration of the unmodified reference's train_CGAN (the gradients
    its two Adam optimizers saw) and the discriminator's forward."""
    import torch
    from oracle import train_ref
    g = golden('training_cgan.npz')
    rng = np.random.RandomState(31)                       # tests/golden/make_golden.py:cgan_data
    X = rng.randn(24, 2, 64, 64).astype('float32')
    Y = (0.5 * np.roll(X, 1, axis=-1) - 0.25 * np.roll(X, 2, axis=-2) + 0.3 * rng.randn(24, 2, 64, 64)).astype('float32')
    np.random.seed(0)
    order = np.arange(24)
    np.random.shuffle(order)
    idx = order[:4]
    rz, re = np.random.RandomState(77), np.random.RandomState(78)
    z1, z2 = (torch.as_tensor(rz.randn(4, 2, 64, 64).astype('float32')) for _ in range(2))
    eps = torch.as_tensor(re.rand(4).astype('float32')).reshape(4, 1, 1, 1)
    coin = int(np.random.randint(0, 2, 1)[0])
    G, D = train_ref.Net(_golden_sd(g, 'G_init')), train_ref.Disc(_golden_sd(g, 'D_init'))
    G.train(); D.train()
    xin = torch.as_tensor(np.random.RandomState(5).randn(3, 6, 64, 64).astype('float32'))
    assert np.abs(D(xin).detach().numpy().reshape(-1) - g['D_forward']).max() < 1e-6
    optD = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.5, 0.999))
    recorded = {}
    real_step = optD.step

    def step():
        recorded.update({k: p.grad.numpy().copy() for k, p in D.net.named_parameters()})
        real_step()
    optD.step = step
    train_ref.cgan_iteration(G, D, optD, None, torch.as_tensor(X[idx]), torch.as_tensor(Y[idx]), z1, z2, eps, coin, True)
    for k, v in _golden_sd(g, 'D_grad0').items():
        assert np.abs(recorded[k] - v).max() <= 1e-5 * np.abs(v).max(), k
    for k, p in G.named_parameters():
        v = g['G_grad0/' + k]
        assert np.abs(p.grad.numpy() - v).max() <= 1e-4 * np.abs(v).max(), k
