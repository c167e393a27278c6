"""Generate the committed golden fixtures by running the UNMODIFIED reference (container only).

    python tests/golden/make_golden.py

/root/reference's in-tree modules (tools/cnn_tools.py, tools/operators.py, tools/stochastic_pyqg.py,
models/*.py) are imported as-is on top of oracle/pyqg_shim.py (pyqg itself is not installable here, see
oracle/pyqg_shim.py header).  The outputs written next to this file are what tests/ compare against on
machines where /root/reference does not exist (the GPU box).

Fixtures
--------
weights_{gan,vae,gz_mean,gz_var}.npz  shipped nx=48 networks (Google-Colab/{GAN,VAE,GZ}/*.pt) as float32 arrays
                                      keyed by the state-dict names, + x_std / y_std from the json scalers
closure_48.npz     q (2,48,48), injected latent noise, and the reference classes' predict_snapshot /
                   predict_mean_snapshot / __call__ outputs for GAN, VAE, GZ
coupled_48.npz     3 steps of reference ``stochastic_QGModel`` + ``CVAERegression`` (AR1 nsteps=1 and nsteps=4;
                   constant nsteps=2) with recorded noise and states
operators_128.npz  Operator1/2/5, cut_off, fft_interpolate, PV_subgrid_forcing(none, 3/2-rule) on a 128^2 field
operators_128_more.npz  Operator4 and PV_subgrid_forcing with the '2/3-rule' (and Operator4 with none / 3/2-rule), same field
samplers.npz       AR1 / constant sampler sequences
ispec.npz          calc_ispec (tools/spectral_tools.py:103-180) of seeded spectra at nx = 48, 64 for every option combination
initial_condition.npz  set_initial_condition (tools/simulate.py:147-168) under np.random.seed for nx = 48, 64, 96 (two
                   successive members each)
training_cvae.npz  CVAERegression.compute_loss (ELBO, adaptive and fixed decoder variance) losses and autograd gradients of a small
                   encoder / decoder pair with the recorded reparameterisation noise, and a whole ``train_CVAE`` run (4 epochs, batch 8)
training_cgan.npz  a whole ``train_CGAN`` run (WGAN-GP, 2 epochs x 6 iterations, small G, DCGAN_discriminator with ndf = 8 at 64 x 64) with
                   seeded random draws: the gradients the two Adam optimizers saw in the first iteration, the loss logs, final G / D
stats_layout.json  variables / dimensions / dtypes of the shipped training logs (Google-Colab/{GAN,VAE}/stats.nc, GZ/stats_var.nc)
training.npz       the reference's training arithmetic on a small AndrewCNN (2 -> 16 -> 12 -> 12 -> 8 -> 2, 16 x 16 images):
                   loss and autograd gradients of ``compute_loss`` in training mode for AndrewCNN and VarCNN (softplus head),
                   BatchNorm running statistics after that forward, and a whole ``cnn_tools.train`` run (4 epochs, batch 8,
                   Adam + MultiStepLR, np.random.seed(0) shuffling): final state_dict and the loss log
"""
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402

ref_import.import_reference()
import pyqg  # noqa: E402  (the shim)
from pyqg_generative.tools import cnn_tools, operators  # noqa: E402
from pyqg_generative.tools.stochastic_pyqg import stochastic_QGModel, AR1_sampler, constant_sampler  # noqa: E402
from pyqg_generative.tools.parameters import EDDY_PARAMS  # noqa: E402
from pyqg_generative.models.cgan_regression import CGANRegression  # noqa: E402
from pyqg_generative.models.cvae_regression import CVAERegression  # noqa: E402
from pyqg_generative.models.mean_var_model import MeanVarModel  # noqa: E402

COLAB = os.path.join(ref_import.REFERENCE_ROOT, 'Google-Colab')


def synthetic_q(n, seed, slope=-1.5):
    """Gaussian random field with the shipped x_scale stds (7.78e-6 / 1.05e-6), red spectrum, truncated at 0.65*pi/dx."""
    rng = np.random.RandomState(seed)
    m = pyqg.QGModel(nx=n, log_level=0)
    out = []
    for std in (7.784383342368528e-06, 1.0471941322975908e-06):
        h = np.fft.rfftn(rng.randn(n, n))
        amp = np.where(m.wv > 0, (m.wv / m.dk + 1.0) ** slope, 0.0) * (m.wv * m.dx <= 0.65 * np.pi)
        f = np.fft.irfftn(h * amp)
        out.append(f / f.std() * std)
    return np.stack(out).astype('float32').astype('float64')   # exactly representable in f32 -> small fixtures


def save_weights():
    def dump(pt, name, folder):
        sd = torch.load(os.path.join(COLAB, folder, pt), map_location='cpu')
        arrs = {k: v.numpy() for k, v in sd.items()}
        xs = cnn_tools.ChannelwiseScaler().read('x_scale.json', os.path.join(COLAB, folder))
        ys = cnn_tools.ChannelwiseScaler().read('y_scale.json', os.path.join(COLAB, folder))
        arrs['x_std'] = xs.std.reshape(-1)
        arrs['y_std'] = ys.std.reshape(-1)
        np.savez_compressed(os.path.join(HERE, name), **arrs)
    dump('G.pt', 'weights_gan.npz', 'GAN')
    dump('decoder.pt', 'weights_vae.npz', 'VAE')
    dump('net_mean.pt', 'weights_gz_mean.npz', 'GZ')
    dump('net_var.pt', 'weights_gz_var.npz', 'GZ')


def load_reference_models():
    tmp = tempfile.mkdtemp(prefix='qgb_gan_')
    for f in os.listdir(os.path.join(COLAB, 'GAN')):
        shutil.copy(os.path.join(COLAB, 'GAN', f), tmp)
    # D.pt is a missing large blob; the discriminator is never used at inference
    torch.manual_seed(0)
    torch.save(cnn_tools.DCGAN_discriminator(6, bn='None', nx=48).state_dict(), os.path.join(tmp, 'D.pt'))
    gan = CGANRegression(folder=tmp, nx=48)
    vae = CVAERegression(folder=os.path.join(COLAB, 'VAE'))
    gz = MeanVarModel(folder=os.path.join(COLAB, 'GZ'))
    return gan, vae, gz


class _M(object):
    """Minimal model view consumed by predict_snapshot / __call__ (reads m.q, m.ny, m.nx, sampler)."""


def closure_fixture(gan, vae, gz):
    n = 48
    q = synthetic_q(n, 11)
    rng = np.random.RandomState(5)
    z32 = rng.randn(1, 2, n, n).astype('float32')
    z64 = rng.randn(2, n, n)
    m = _M()
    m.q, m.ny, m.nx = q, n, n
    out = dict(q=q.astype('float32'), z32=z32, z64=z64)
    out['gan_snapshot'] = gan.predict_snapshot(m, z32)
    out['vae_snapshot'] = vae.predict_snapshot(m, z32)
    out['gz_snapshot'] = gz.predict_snapshot(m, z64)
    out['gz_mean_snapshot'] = gz.predict_mean_snapshot(m)
    # __call__ with seeded host RNG (np.random.randn is what generate_latent_noise draws from)
    for name, model in (('gan', gan), ('vae', vae), ('gz', gz)):
        m.sampling_type = 'AR1'
        m.noise_sampler = AR1_sampler(1)
        np.random.seed(77)
        out[name + '_call'] = model(m)
        out[name + '_call_noise'] = np.array(m.noise_sampler.noise)
    # batched forward of the generator itself (torch API surface: generate(x, z))
    xb = torch.as_tensor(rng.randn(3, 2, n, n).astype('float32'))
    zb = torch.as_tensor(rng.randn(3, 2, n, n).astype('float32'))
    gan.G.eval()
    with torch.no_grad():
        out['generate_x'] = xb.numpy()
        out['generate_z'] = zb.numpy()
        out['gan_generate'] = gan.generate(xb, zb).numpy()
    np.savez_compressed(os.path.join(HERE, 'closure_48.npz'), **out)


def coupled_fixture(vae):
    n = 48
    params = dict(EDDY_PARAMS.nx(n)._update({'dt': 7200.0, 'log_level': 0}))
    out = {}
    for tag, sampling, nsteps in (('ar1_1', 'AR1', 1), ('ar1_4', 'AR1', 4), ('const_2', 'constant', 2)):
        p = dict(params)
        p['parameterization'] = vae
        np.random.seed(3)
        m = stochastic_QGModel(p, sampling, nsteps)
        m.q = synthetic_q(n, 21)
        m._invert()
        qs, zs, dqs = [m.q.copy()], [], []
        np.random.seed(123)
        for _ in range(4):
            m._step_forward()
            qs.append(m.q.copy())
            zs.append(np.array(m.noise_sampler.noise).copy())
            dqs.append(np.array(m.PV_forcing).copy())
        out[tag + '_q'] = np.stack(qs)
        out[tag + '_noise'] = np.stack(zs)
        out[tag + '_forcing'] = np.stack(dqs)
    np.savez_compressed(os.path.join(HERE, 'coupled_48.npz'), **out)


def operators_fixture():
    n = 128
    q = synthetic_q(n, 31, slope=-1.0)
    params = dict(EDDY_PARAMS.nx(n))
    params.pop('nx')
    out = dict(q=q.astype('float32'))
    for nc in (32, 48, 64):
        for name in ('Operator1', 'Operator2', 'Operator5', 'cut_off'):
            out['%s_%d' % (name, nc)] = getattr(operators, name)(q, nc)
    x = np.random.RandomState(1).randn(2, 48, 48)
    out['interp_in'] = x
    out['interp_48_72'] = operators.fft_interpolate(x, 48, 72)
    out['interp_48_32'] = operators.fft_interpolate(x, 48, 32)
    for opname in ('Operator1', 'Operator2', 'Operator5'):
        for dealias, tag in (('none', 'none'), ('3/2-rule', '32')):
            forcing, mf, m = operators.PV_subgrid_forcing(q, 64, getattr(operators, opname), dict(params), dealias)
            out['S_%s_%s' % (opname, tag)] = forcing
            if dealias == 'none':
                out['qf_%s' % opname] = mf.q
                out['uf_%s' % opname] = mf.u
                out['vf_%s' % opname] = mf.v
                out['pf_%s' % opname] = mf.p
    np.savez_compressed(os.path.join(HERE, 'operators_128.npz'), **out)
    # Operator4 (:213-214) and the '2/3-rule' of advect (:253-257) on the same field (separate file: same q as above)
    out2 = {}
    for nc in (32, 48, 64):
        out2['Operator4_%d' % nc] = operators.Operator4(q, nc)
    for opname in ('Operator1', 'Operator2', 'Operator4', 'Operator5'):
        forcing, mf, m = operators.PV_subgrid_forcing(q, 64, getattr(operators, opname), dict(params), '2/3-rule')
        out2['S_%s_23' % opname] = forcing
    for dealias, tag in (('none', 'none'), ('3/2-rule', '32')):
        forcing, mf, m = operators.PV_subgrid_forcing(q, 64, operators.Operator4, dict(params), dealias)
        out2['S_Operator4_%s' % tag] = forcing
        out2['qf_Operator4'] = mf.q
        out2['uf_Operator4'] = mf.u
    np.savez_compressed(os.path.join(HERE, 'operators_128_more.npz'), **out2)


def samplers_fixture():
    out = {}
    for n in (1, 4, -1):
        s = AR1_sampler(n)
        rng = np.random.RandomState(9)
        seq, xi = [], []
        for _ in range(6):
            def gen():
                v = rng.randn(3)
                xi.append(v)
                return v
            s.update(gen)
            seq.append(np.array(s.noise))
        out['ar1_%d' % n] = np.stack(seq)
        out['ar1_%d_xi' % n] = np.stack(xi)
    for n in (1, 3):
        s = constant_sampler(n)
        rng = np.random.RandomState(9)
        seq, flags = [], []
        for _ in range(8):
            flags.append(s.update(lambda: rng.randn(3)))
            seq.append(np.array(s.noise))
        out['const_%d' % n] = np.stack(seq)
        out['const_%d_flags' % n] = np.array(flags)
    np.savez_compressed(os.path.join(HERE, 'samplers.npz'), **out)


def ispec_fixture():
    from pyqg_generative.tools.spectral_tools import calc_ispec
    out = {}
    for n in (48, 64):
        m = pyqg.QGModel(nx=n, log_level=0)
        rng = np.random.RandomState(100 + n)
        spec = rng.rand(n, n // 2 + 1) * np.exp(-(m.wv / (10 * m.dk)) ** 1.5)
        out['spec_%d' % n] = spec
        for avg in (True, False):
            for trunc in (True, False):
                for nd in (False, True):
                    for nf in (1, 2):
                        kr, ph = calc_ispec(m, spec, averaging=avg, truncate=trunc, nd_wavenumber=nd, nfactor=nf)
                        tag = '%d_%d%d%d%d' % (n, avg, trunc, nd, nf)
                        out['kr_' + tag], out['ph_' + tag] = kr, ph
    np.savez_compressed(os.path.join(HERE, 'ispec.npz'), **out)


def initial_condition_fixture():
    from pyqg_generative.tools.simulate import set_initial_condition
    out = {}
    for n in (48, 64, 96):
        models = [pyqg.QGModel(nx=n, log_level=0) for _ in range(2)]   # (the pyqg constructor draws its own default IC)
        np.random.seed(1000 + n)
        qs = []
        for m in models:
            set_initial_condition(m)
            qs.append(m.q.copy())
        out['q_%d' % n] = np.stack(qs)
    np.savez_compressed(os.path.join(HERE, 'initial_condition.npz'), **out)


TRAIN_HIDDEN = [16, 12, 12, 8]


def training_fixture():
    """tools/cnn_tools.py:645-700 ``train`` and :177-182 ``compute_loss`` run as they are (CPU torch, fp32)."""
    from pyqg_generative.models.mean_var_model import VarCNN
    out = {}
    rng = np.random.RandomState(11)
    x = rng.randn(6, 2, 16, 16).astype('float32')
    y = rng.randn(6, 2, 16, 16).astype('float32')
    out['grad_x'], out['grad_y'] = x, y
    for tag, cls, target in (('mean', cnn_tools.AndrewCNN, y), ('var', VarCNN, y ** 2)):
        torch.manual_seed(3)
        net = cls(2, 2, hidden_channels=TRAIN_HIDDEN)
        for k, v in net.state_dict().items():
            out['%s_init/%s' % (tag, k)] = v.numpy().copy()
        net.train()
        loss = net.compute_loss(torch.as_tensor(x), torch.as_tensor(target))['loss']
        loss.backward()
        out['%s_loss' % tag] = np.float64(loss.item())
        for k, p in net.named_parameters():
            out['%s_grad/%s' % (tag, k)] = p.grad.numpy().copy()
        for k, v in net.state_dict().items():
            if 'running' in k:
                out['%s_after/%s' % (tag, k)] = v.numpy().copy()
    # a whole training run
    X_train = rng.randn(20, 2, 16, 16).astype('float32')
    W = rng.randn(2, 2).astype('float32')
    Y_train = (np.einsum('ij,bjyx->biyx', W, np.roll(X_train, 1, axis=-1)) + 0.3 * X_train ** 2).astype('float32')
    X_test = rng.randn(8, 2, 16, 16).astype('float32')
    Y_test = (np.einsum('ij,bjyx->biyx', W, np.roll(X_test, 1, axis=-1)) + 0.3 * X_test ** 2).astype('float32')
    torch.manual_seed(4)
    net = cnn_tools.AndrewCNN(2, 2, hidden_channels=TRAIN_HIDDEN)
    for k, v in net.state_dict().items():
        out['run_init/%s' % k] = v.numpy().copy()
    np.random.seed(0)
    cnn_tools.train(net, X_train, Y_train, X_test, Y_test, num_epochs=4, batch_size=8, learning_rate=1e-3, device='cpu')
    for k, v in net.state_dict().items():
        out['run_final/%s' % k] = v.numpy().copy()
    out['run_loss'] = np.array(net.log_dict['loss'], dtype=np.float64)
    out['run_loss_test'] = np.array(net.log_dict['loss_test'], dtype=np.float64)
    out.update(X_train=X_train, Y_train=Y_train, X_test=X_test, Y_test=Y_test)
    np.savez_compressed(os.path.join(HERE, 'training.npz'), **out)


def _small_cvae(decoder_var='adaptive'):
    tmp = tempfile.mkdtemp()
    net = CVAERegression(folder=tmp, hidden_channels=TRAIN_HIDDEN, decoder_var=decoder_var)
    net.encoder = cnn_tools.AndrewCNN(4, 4, hidden_channels=TRAIN_HIDDEN)      # a small encoder keeps the fixture small
    shutil.rmtree(tmp, ignore_errors=True)
    return net


def cvae_fixture():
    """models/cvae_regression.py:165-230 ``forward`` / ``compute_loss`` and :250-300 ``train_CVAE`` run as they are (CPU torch, fp32)
    on a small encoder / decoder pair; the reparameterisation draws (``torch.randn_like``) are recorded so that the device trainer can
    be fed the same noise."""
    from pyqg_generative.models import cvae_regression as ref_cvae
    out = {}
    rng = np.random.RandomState(21)
    x = rng.randn(6, 2, 16, 16).astype('float32')
    y = (0.5 * np.roll(x, 1, axis=-1) + 0.3 * rng.randn(6, 2, 16, 16)).astype('float32')
    out['grad_x'], out['grad_y'] = x, y
    drawn = []
    real_randn_like = torch.randn_like

    def recording_randn_like(t, *a, **k):
        r = real_randn_like(t, *a, **k)
        drawn.append(r.numpy().copy())
        return r
    torch.randn_like = recording_randn_like
    try:
        for tag, dv in (('adaptive', 'adaptive'), ('fixed01', 0.1)):
            torch.manual_seed(5)
            net = _small_cvae(dv)
            for name, sub in (('enc', net.encoder), ('dec', net.decoder)):
                for k, v in sub.state_dict().items():
                    out['%s_%s_init/%s' % (tag, name, k)] = v.numpy().copy()
            net.encoder.train(); net.decoder.train()
            del drawn[:]
            losses = net.compute_loss(torch.as_tensor(x), torch.as_tensor(y), 0 * torch.as_tensor(y))
            losses['loss'].backward()
            out['%s_eps' % tag] = drawn[0]
            out['%s_losses' % tag] = np.array([float(losses[k]) for k in
                                               ('loss', 'loss_recon', 'loss_KL', 'MSE', 'var_latent', 'var_aggr')])
            for name, sub in (('enc', net.encoder), ('dec', net.decoder)):
                for k, p_ in sub.named_parameters():
                    out['%s_%s_grad/%s' % (tag, name, k)] = p_.grad.numpy().copy()
        # a whole train_CVAE run (the per-epoch offline scores need datasets: replaced by a stub, they do not feed back)
        X_train = rng.randn(20, 2, 16, 16).astype('float32')
        Y_train = (0.5 * np.roll(X_train, 1, axis=-1) + 0.3 * rng.randn(20, 2, 16, 16)).astype('float32')
        torch.manual_seed(6)
        net = _small_cvae('adaptive')
        for name, sub in (('enc', net.encoder), ('dec', net.decoder)):
            for k, v in sub.state_dict().items():
                out['run_%s_init/%s' % (name, k)] = v.numpy().copy()
        ref_cvae.evaluate_prediction = lambda *a, **k: dict(L2_mean=0., L2_total=0., L2_residual=0., var_ratio=[0., 0.])
        del drawn[:]
        np.random.seed(0)
        optim_loss, _, _ = ref_cvae.train_CVAE(net, None, None, X_train, Y_train, num_epochs=4, batch_size=8, learning_rate=1e-3)
        out['run_eps'] = np.concatenate([d.reshape(-1) for d in drawn])
        for name, sub in (('enc', net.encoder), ('dec', net.decoder)):
            for k, v in sub.state_dict().items():
                out['run_%s_final/%s' % (name, k)] = v.numpy().copy()
        for k, v in optim_loss.items():
            out['run_log/%s' % k] = np.array(v, dtype=np.float64)
        out.update(X_train=X_train, Y_train=Y_train)
    finally:
        torch.randn_like = real_randn_like
    np.savez_compressed(os.path.join(HERE, 'training_cvae.npz'), **out)


def cgan_data(n=24, nx=64, seed=31):
    """Synthetic (x, y) pairs of the CGAN fixture; the test regenerates them from the same seed."""
    rng = np.random.RandomState(seed)
    x = rng.randn(n, 2, nx, nx).astype('float32')
    y = (0.5 * np.roll(x, 1, axis=-1) - 0.25 * np.roll(x, 2, axis=-2) + 0.3 * rng.randn(n, 2, nx, nx)).astype('float32')
    return x, y


class SeededDraws(object):
    """The random draws of train_CGAN from seeded numpy streams (so that the fixture need not store them):
    z = float32 standard normal (RandomState(77)), eps = float32 uniform (RandomState(78))."""

    def __init__(self):
        self.rz, self.re = np.random.RandomState(77), np.random.RandomState(78)

    def z(self, shape):
        return self.rz.randn(*shape).astype('float32')

    def eps(self, n):
        return self.re.rand(n).astype('float32')


def cgan_fixture():
    """models/cgan_regression.py:222-300 ``train_CGAN`` (with ``gradient_penalty`` :173-195) run as it is on CPU torch: a small
    generator (4 -> 16 -> 12 -> 12 -> 8 -> 2) and DCGAN_discriminator(6, ndf=8, bn='None', nx=64), 24 samples of 64 x 64, batch 4,
    2 epochs (generator steps at i = 0, 5).  ``torch.randn`` / ``torch.rand`` are replaced by seeded numpy streams for the duration
    of the run (``SeededDraws``), np.random (shuffling, the coin of the gradient penalty) is seeded; a recording subclass of
    torch.optim.Adam keeps the gradients each optimizer saw in the first iteration."""
    from pyqg_generative.models import cgan_regression as ref_cgan
    out = {}
    X_train, Y_train = cgan_data()
    torch.manual_seed(8)
    tmp = tempfile.mkdtemp()
    net = CGANRegression(folder=tmp, nx=64, hidden_channels=TRAIN_HIDDEN)
    shutil.rmtree(tmp, ignore_errors=True)
    net.D = cnn_tools.DCGAN_discriminator(6, ndf=8, bn='None', nx=64)
    net.D.apply(cnn_tools.weights_init)
    with torch.no_grad():                   # N(0, 0.02) leaves D(x) ~ 1e-5 with ndf = 8: scale to N(0, 0.1) so that every loss term counts
        for p_ in net.D.parameters():
            p_.mul_(5.0)
    for k, v in net.G.state_dict().items():
        out['G_init/%s' % k] = v.numpy().copy()
    for k, v in net.D.state_dict().items():
        out['D_init/%s' % k] = v.numpy().copy()
    draws = SeededDraws()
    seen = []

    class RecordingAdam(torch.optim.Adam):
        def step(self, *a, **k):
            if len(seen) < 2:
                seen.append([p.grad.detach().numpy().copy() for g in self.param_groups for p in g['params']])
            return super().step(*a, **k)
    real = (torch.randn, torch.rand, ref_cgan.optim.Adam, ref_cgan.evaluate_prediction)
    torch.randn = lambda *shape, **k: torch.from_numpy(draws.z(tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else shape))
    torch.rand = lambda *shape, **k: torch.from_numpy(draws.eps(shape[0]).reshape(shape))
    ref_cgan.optim.Adam = RecordingAdam
    ref_cgan.evaluate_prediction = lambda *a, **k: dict(L2_mean=0., L2_total=0., L2_residual=0., var_ratio=[0., 0.])
    try:
        np.random.seed(0)
        optim_loss, _, _ = ref_cgan.train_CGAN(net, None, None, X_train, Y_train, num_epochs=2, batch_size=4, learning_rate=2e-4)
    finally:
        torch.randn, torch.rand, ref_cgan.optim.Adam, ref_cgan.evaluate_prediction = real
    for (k, _), g in zip(net.D.named_parameters(), seen[0]):
        out['D_grad0/%s' % k] = g
    for (k, _), g in zip(net.G.named_parameters(), seen[1]):
        out['G_grad0/%s' % k] = g
    for k, v in net.G.state_dict().items():
        out['G_final/%s' % k] = v.numpy().copy()
    for k, v in net.D.state_dict().items():
        out['D_final/%s' % k] = v.numpy().copy()
    for k, v in optim_loss.items():
        out['log/%s' % k] = np.array(v, dtype=np.float64)
    # D on a fixed input with the initial weights (forward parity of the discriminator alone)
    xin = np.random.RandomState(5).randn(3, 6, 64, 64).astype('float32')
    D0 = cnn_tools.DCGAN_discriminator(6, ndf=8, bn='None', nx=64)
    D0.load_state_dict({k[7:]: torch.as_tensor(v) for k, v in out.items() if k.startswith('D_init/')})
    out['D_forward'] = D0(torch.as_tensor(xin)).detach().numpy().reshape(-1)
    np.savez_compressed(os.path.join(HERE, 'training_cgan.npz'), **out)


def stats_layout_fixture():
    """Variables, dimensions, dtypes and coordinate values of the training logs the reference ships (Google-Colab/{GAN,VAE}/stats.nc,
    GZ/stats_var.nc; written by ``loss_to_xarray(...).to_netcdf`` / ``log_to_xarray``): what our ``write_log`` must reproduce."""
    import json
    from scipy.io import netcdf_file
    out = {}
    for tag, rel_path in (('GAN', 'GAN/stats.nc'), ('VAE', 'VAE/stats.nc'), ('GZ_var', 'GZ/stats_var.nc')):
        with open(os.path.join(COLAB, rel_path), 'rb') as fh:
            magic = fh.read(4)
        with netcdf_file(os.path.join(COLAB, rel_path), 'r', mmap=False) as nc:
            out[tag] = dict(magic=list(magic), variables={k: dict(dims=list(v.dimensions), dtype=v.data.dtype.str) for k, v in nc.variables.items()},
                            lev=[int(x) for x in nc.variables['lev'][:]] if 'lev' in nc.variables else None,
                            epoch_first=int(nc.variables['epoch'][0]))
    with open(os.path.join(HERE, 'stats_layout.json'), 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == '__main__':
    if '--stats-layout' in sys.argv:
        stats_layout_fixture()
        sys.exit(0)
    if '--cgan' in sys.argv:
        cgan_fixture()
        sys.exit(0)
    if '--cvae' in sys.argv:
        cvae_fixture()
        sys.exit(0)
    if '--only-new' in sys.argv:       # fixtures added in round 2 (the others are unchanged)
        ispec_fixture()
        initial_condition_fixture()
        training_fixture()
        sys.exit(0)
    if '--operators' in sys.argv:
        operators_fixture()
        sys.exit(0)
    if '--training' in sys.argv:
        training_fixture()
        sys.exit(0)
    save_weights()
    gan, vae, gz = load_reference_models()
    closure_fixture(gan, vae, gz)
    coupled_fixture(vae)
    operators_fixture()
    samplers_fixture()
    ispec_fixture()
    initial_condition_fixture()
    training_fixture()
    cvae_fixture()
    cgan_fixture()
    stats_layout_fixture()
    for f in sorted(os.listdir(HERE)):
        print('%10d  %s' % (os.path.getsize(os.path.join(HERE, f)), f))
