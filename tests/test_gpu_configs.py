"""Coupled-step parity for the configurations BASELINE.json names beyond the 64^2 eddy + CGAN headline: configs[3] = jet
configuration (rek 7e-8, delta 0.1, beta 1e-11; tools/parameters.py:37) with the CVAE and GZ closures at 64^2, 48^2 and 96^2 in
the tensor-core precision they are benchmarked in, against the oracle (pyqg shim step + fp32 CPU AndrewCNN) from identical
states and injected latent noise.  Tolerances: closure output <= 1e-3 relative (L2) for the tensor-core path; spectral state
<= 1e-10 relative per step once the oracle is fed the SAME forcing (this separates the fp64 step from the CNN precision)."""
import numpy as np
import pytest

from conftest import golden, golden_state_dict, write_model_folder
from oracle import cnn_ref, pyqg_shim

pytestmark = pytest.mark.gpu
JET = dict(rek=7e-08, delta=0.1, beta=1e-11)
TC_TOL = 1e-3


def rel_l2(a, b):
    a, b = np.asarray(a, 'float64'), np.asarray(b, 'float64')
    return np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum())


class _Slot(object):
    parameterization_type = 'q_parameterization'
    dq = None

    def __call__(self, m):
        return self.dq


def developed_state(n, members, seed):
    """Gaussian random fields with the shipped x_scale stds and a red spectrum truncated at the filter cut-off."""
    rng = np.random.RandomState(seed)
    m = pyqg_shim.QGModel(nx=n, log_level=0)
    out = np.empty((members, 2, n, n))
    for z, std in enumerate((7.784383342368528e-06, 1.0471941322975908e-06)):
        h = np.fft.rfftn(rng.randn(members, n, n), axes=(-2, -1))
        amp = np.where(m.wv > 0, (m.wv / m.dk + 1.0) ** -1.5, 0.0) * (m.wv * m.dx <= 0.65 * np.pi)
        f = np.fft.irfftn(h * amp, s=(n, n), axes=(-2, -1))
        out[:, z] = f / f.std(axis=(-2, -1), keepdims=True) * std
    return out


@pytest.mark.parametrize('kind,nx', [('vae', 64), ('gz', 64), ('vae', 96), ('gz', 96), ('vae', 48), ('gz', 48)])
def test_jet_coupled_steps_in_tensor_core_precision_match_oracle(tmp_path, kind, nx):
    from pyqg_generative_b200.models.cvae_regression import CVAERegression
    from pyqg_generative_b200.models.mean_var_model import MeanVarModel
    from pyqg_generative_b200.tools.stochastic_pyqg import stochastic_QGModel
    dt = 14400.0 if nx <= 64 else 7200.0
    B, nsteps = 3, 3
    if kind == 'vae':
        model = CVAERegression(folder=write_model_folder(tmp_path, 'vae'), precision='tc')
        nets = [golden_state_dict('weights_vae.npz')[0]]
        _, xs, ys = golden_state_dict('weights_vae.npz')
        zdtype = 'float32'
    else:
        model = MeanVarModel(folder=write_model_folder(tmp_path, 'gz'), precision='tc')
        nets = [golden_state_dict('weights_gz_mean.npz')[0], golden_state_dict('weights_gz_var.npz')[0]]
        _, xs, ys = golden_state_dict('weights_gz_mean.npz')
        zdtype = 'float64'
    m = stochastic_QGModel(dict(nx=nx, dt=dt, log_level=0, tmax=1e12, tavestart=1e12, members=B, parameterization=model,
                                precision='tc', **JET), 'AR1', 1)
    q0 = developed_state(nx, B, 40 + nx)
    m.set_q(q0)
    # two oracle ensembles: ``free`` uses its own fp32 CNN forcing (end-to-end comparison), ``fed`` is given the engine's
    # forcing every step (isolates the fp64 spectral step)
    free, fed = [], []
    for b in range(B):
        for lst in (free, fed):
            o = pyqg_shim.QGModel(nx=nx, dt=dt, log_level=0, parameterization=_Slot(), **JET)
            o.q = q0[b]
            lst.append(o)
    rng = np.random.RandomState(7)
    for step in range(nsteps):
        z = rng.randn(B, 2, nx, nx).astype(zdtype)
        m.set_latent(z)
        m._step_forward()
        f_gpu = m.PV_forcing
        q_free_in = np.stack([o.q for o in free])
        dq = cnn_ref.demean(cnn_ref.predict_snapshot(kind, nets, xs, ys, q_free_in, z))
        err_f = rel_l2(f_gpu, dq)
        for b in range(B):
            free[b].q_parameterization.dq = dq[b]
            free[b]._step_forward()
            fed[b].q_parameterization.dq = f_gpu[b]
            fed[b]._step_forward()
        q = m.q
        err_fed = max(np.abs(q[b] - fed[b].q).max() / np.abs(fed[b].q).max() for b in range(B))
        err_free = max(np.abs(q[b] - free[b].q).max() / np.abs(free[b].q).max() for b in range(B))
        print('%s %d step %d: forcing rel-L2 %.2e, q vs oracle(fed the same forcing) %.2e, q vs oracle(free) %.2e'
              % (kind, nx, step, err_f, err_fed, err_free))
        assert err_f < TC_TOL, (step, err_f)
        assert err_fed < 1e-10, (step, err_fed)
        assert err_free < 2e-5, (step, err_free)          # dt * forcing / q ~ 1.4e-2 per step, times the forcing error (measured 2.6e-6)


def test_auto_precision_picks_the_fast_plan_only_where_it_meets_the_tolerance(tmp_path):
    """precision='auto' (qgb_closure_precision): the shipped GAN generator runs layer 2 in a single pass (measured 4-5e-4),
    the shipped VAE decoder (1.3e-3 single-pass) keeps the two-pass plan; both stay within 1e-3 of the fp32 reference."""
    from pyqg_generative_b200.models.cgan_regression import CGANRegression
    from pyqg_generative_b200.models.cvae_regression import CVAERegression
    from pyqg_generative_b200.tools.stochastic_pyqg import stochastic_QGModel
    c = golden('closure_48.npz')
    expect = {}
    for name, cls, kw in (('gan', CGANRegression, dict(nx=48)), ('vae', CVAERegression, {})):
        model = cls(folder=write_model_folder(tmp_path, name), precision='auto', **kw)
        m = stochastic_QGModel(dict(nx=48, dt=7200.0, log_level=0, members=3, parameterization=model, precision='auto'), 'AR1', 1)
        m.q = c['q'].astype('float64')
        assert m.closure_precision()[0] == 'auto'             # not calibrated before the first evaluation
        m.set_latent(np.concatenate([c['z32']] * 3))
        f = m.closure_eval()
        chosen, err = m.closure_precision()
        print(name, chosen, err)
        expect[name] = chosen
        assert chosen in ('tc', 'tc_fast') and err['tc_l2'] < 1e-3
        assert (chosen == 'tc_fast') == (err['tc_fast_l2'] <= 7e-4)
        ref = cnn_ref.demean(c[name + '_snapshot'])
        assert rel_l2(f[1], ref) < TC_TOL
    assert expect == {'gan': 'tc_fast', 'vae': 'tc'}, expect


def test_weighted_attached_closure_is_not_weighted_twice(tmp_path):
    """calling ``(w * model)(m)`` directly on a model the closure is attached to returns the engine's forcing (which
    already carries w), not w^2 * forcing."""
    from pyqg_generative_b200.models.cgan_regression import CGANRegression
    from pyqg_generative_b200.tools.stochastic_pyqg import stochastic_QGModel
    c = golden('closure_48.npz')
    model = CGANRegression(folder=write_model_folder(tmp_path, 'gan'), nx=48)
    par = 0.5 * model
    m = stochastic_QGModel(dict(nx=48, dt=7200.0, log_level=0, members=2, parameterization=par), 'AR1', 1)
    m.q = c['q'].astype('float64')
    m.set_latent(np.concatenate([c['z32']] * 2))
    out = par(m)
    ref = 0.5 * cnn_ref.demean(c['gan_snapshot'])
    assert np.abs(out[0] - ref).max() < 5e-5 * np.abs(ref).max()


def test_coupled_run_inherits_the_precision_the_closure_was_built_with(tmp_path):
    """README usage: ``CGANRegression(folder, precision='tc')`` + pyqg parameters WITHOUT a ``precision`` key must run the closure on
    the tensor-core path (it used to fall back to fp32 silently); the model's own keyword still wins."""
    from pyqg_generative_b200.models.cgan_regression import CGANRegression
    from pyqg_generative_b200.tools.stochastic_pyqg import stochastic_QGModel
    folder = write_model_folder(tmp_path, 'gan')
    for closure_prec, model_prec, want in (('tc', None, 'tc'), ('fp32', None, 'fp32'), ('tc', 'fp32', 'fp32'), ('fp32', 'tc', 'tc')):
        model = CGANRegression(folder=folder, nx=48, precision=closure_prec)
        params = dict(nx=48, dt=7200.0, log_level=0, members=2, parameterization=model)
        if model_prec is not None:
            params['precision'] = model_prec
        m = stochastic_QGModel(params, 'AR1', 1)
        assert m.closure_precision()[0] == want, (closure_prec, model_prec, m.closure_precision())


def test_two_devices_in_one_process_and_odd_cluster_sizes():
    """(a) function attributes are per device -- handles on two devices of one process both run the kernels that need
    > 48 KB of dynamic shared memory; (b) nx whose 8-CTA cluster split leaves a remainder (162 = 2 * 3^4) picks a cluster size
    that divides it and still matches the oracle."""
    import torch
    from pyqg_generative_b200.tools.stochastic_pyqg import EnsembleQGModel
    q0 = developed_state(64, 2, 5)
    devs = [0] + ([1] if torch.cuda.device_count() > 1 else [])
    outs = []
    for d in devs:
        m = EnsembleQGModel(nx=64, dt=14400., members=2, device=d, log_level=0)
        m.set_q(q0)
        m._step_forward(3)
        outs.append(m.q)
    for o in outs[1:]:
        assert np.array_equal(o, outs[0])
    n = 162
    q0 = developed_state(n, 1, 6)
    m = EnsembleQGModel(nx=n, dt=3600., members=1, log_level=0)
    m.set_q(q0)
    o = pyqg_shim.QGModel(nx=n, dt=3600., log_level=0)
    o.q = q0[0]
    for _ in range(3):
        m._step_forward()
        o._step_forward()
    assert np.abs(m.q[0] - o.q).max() / np.abs(o.q).max() < 1e-10


def test_cuda_graph_replay_is_bit_identical_to_plain_launches(tmp_path):
    """qgb_step replays a captured graph of the step in the steady state; a handle with kernel timing switched on launches
    the same kernels one by one.  Same seed -> bit-identical states, with and without a closure."""
    import ctypes
    from pyqg_generative_b200 import _lib
    from pyqg_generative_b200.models.cgan_regression import CGANRegression
    from pyqg_generative_b200.tools.stochastic_pyqg import EnsembleQGModel, stochastic_QGModel
    c = golden('closure_48.npz')
    q0 = np.stack([c['q'].astype('float64')] * 4) * np.array([1.0, 0.9, 1.1, 0.8]).reshape(4, 1, 1, 1)

    def run(plain, closure):
        if closure:
            model = CGANRegression(folder=write_model_folder(tmp_path, 'gan'), nx=48, precision='tc')
            m = stochastic_QGModel(dict(nx=48, dt=7200.0, log_level=0, tmax=1e12, tavestart=1e12, members=4, parameterization=model,
                                        precision='tc', seed=5), 'AR1', 1)
        else:
            m = EnsembleQGModel(nx=48, dt=7200.0, log_level=0, tmax=1e12, tavestart=1e12, members=4)
        m.set_q(q0)
        if plain:
            _lib.check(m._lib.qgb_profile_all_begin(m._h), m._h)
        m._step_forward(25)
        return m.q, int(m._lib.qgb_graph_replays(m._h))
    for closure in (False, True):
        qa, ra = run(False, closure)
        qb, rb = run(True, closure)
        assert ra >= 20 and rb == 0, (closure, ra, rb)
        assert np.array_equal(qa, qb), closure


def test_latent_noise_generated_inside_layer_one_equals_the_stored_path(tmp_path):
    """north_star (4): with white noise and the tensor-core generator the Philox latent field is generated inside layer 1 of the
    network (no latent kernel, z never stored).  Reading the noise back regenerates it from the same counters; injecting that
    field into a second model (stored-noise path) must give the bit-identical forcing, step after step."""
    from pyqg_generative_b200 import _lib
    from pyqg_generative_b200.models.cgan_regression import CGANRegression
    from pyqg_generative_b200.tools.stochastic_pyqg import stochastic_QGModel
    c = golden('closure_48.npz')
    q0 = np.stack([c['q'].astype('float64')] * 3) * np.array([1.0, 0.9, 1.1]).reshape(3, 1, 1, 1)

    def make(seed):
        model = CGANRegression(folder=write_model_folder(tmp_path, 'gan'), nx=48, precision='tc')
        m = stochastic_QGModel(dict(nx=48, dt=7200.0, log_level=0, tmax=1e12, tavestart=1e12, members=3, member_offset=5,
                                    parameterization=model, precision='tc', seed=seed), 'constant', 1)
        m.set_q(q0)
        return m
    a, b = make(11), make(99)
    for step in range(3):
        l0 = _lib.launch_count()
        fa = a.closure_eval()
        launched = _lib.launch_count() - l0
        za = a.noise_sampler.noise                     # regenerated from the Philox counters
        assert za.dtype == np.float32 and abs(za.std() - 1) < 0.05
        b.set_latent(za)                               # stored path: latent kernel copies the injected field
        fb = b.closure_eval()
        assert np.array_equal(fa, fb), step
        assert launched <= 11, launched                # 8 conv layers + epilogue + draw-counter bump (+ demean for the read-back): no latent kernel
        a._step_forward(1)
        b.set_latent(a.noise_sampler.noise)
        b._step_forward(1)
        assert np.array_equal(a.q, b.q), step
