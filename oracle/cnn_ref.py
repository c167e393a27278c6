"""CPU ORACLE (test infrastructure, NOT product code) -- fp32 restatement of the reference closure path.

Restates, with CPU torch fp32 ops only (no reference import, so it travels to the GPU box):

* ``AndrewCNN.forward`` in eval mode            -- /root/reference/pyqg_generative/tools/cnn_tools.py:79-98,125-176
* ``ChannelwiseScaler.normalize/denormalize``   -- tools/cnn_tools.py:502-553
* ``predict_snapshot`` of the four closures     -- models/cgan_regression.py:157-162, models/cvae_regression.py:131-136,
                                                   models/mean_var_model.py:14-17,105-109, models/ols_model.py:71-75
* ``Parameterization.__call__`` demeaning       -- models/parameterization.py:23-34
* ``AR1_sampler`` / ``constant_sampler``        -- tools/stochastic_pyqg.py:30-72

Pinned against the reference itself: tests/golden/make_golden.py runs the *unmodified* reference classes with
the shipped nx=48 weights (Google-Colab/{GAN,VAE,GZ}) and commits their outputs; tests/test_oracle_pins.py
checks this restatement reproduces them.
"""
import json

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default (cnn_tools.py:97)


def conv_indices(state_dict):
    """Sequential indices of the Conv2d entries of ``AndrewCNN.conv`` (0,3,6,...,21)."""
    idx = sorted({int(k.split('.')[1]) for k in state_dict if k.startswith('conv.') and k.endswith('.weight')
                  and state_dict[k].dim() == 4})
    return idx


def andrew_cnn_forward(state_dict, x, final_softplus=False, dtype=torch.float32):
    """AndrewCNN.forward in eval mode: 8x [circular 'same' conv -> ReLU -> BatchNorm2d(running stats)],
    the last block is the bare conv (cnn_tools.py:137-160).  ``x``: torch float32 (B, Cin, ny, nx)."""
    idx = conv_indices(state_dict)
    if dtype != torch.float32:   # float64 evaluation: used by tests to separate rounding noise from real errors
        state_dict = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in state_dict.items()}
        x = x.to(dtype)
    with torch.no_grad():
        for n, i in enumerate(idx):
            w = state_dict['conv.%d.weight' % i]
            b = state_dict.get('conv.%d.bias' % i)
            p = w.shape[-1] // 2
            x = F.conv2d(F.pad(x, (p, p, p, p), mode='circular'), w, None if b is None else b)
            if n < len(idx) - 1:
                x = F.relu(x)
                j = i + 2
                if 'conv.%d.running_mean' % j in state_dict:
                    x = F.batch_norm(x, state_dict['conv.%d.running_mean' % j],
                                     state_dict['conv.%d.running_var' % j],
                                     state_dict['conv.%d.weight' % j],
                                     state_dict['conv.%d.bias' % j], False, 0.0, BN_EPS)
        if final_softplus:
            x = F.softplus(x)
    return x


def read_scale(path):
    """ChannelwiseScaler.read (cnn_tools.py:547-553): json with stringified nested lists -> float32 (1,C,1,1)."""
    with open(path) as f:
        d = json.load(f)
    return (np.array(eval(d['std'])).astype('float32'), np.array(eval(d['mean'])).astype('float32'))


def predict_snapshot(kind, nets, x_std, y_std, q, noise=None):
    """``predict_snapshot`` of CGAN/CVAE ('gan'/'vae'), MeanVarModel ('gz') and OLSModel ('ols').

    q      : float64 (2,ny,nx) or batched (B,2,ny,nx)
    noise  : gan/vae: float32 (B,2,ny,nx) [reference (1,2,ny,nx)]; gz: float64 (B,2,ny,nx) [reference (2,ny,nx)]
    returns: float64, same leading shape as q (NOT demeaned; demeaning is Parameterization.__call__).
    """
    q = np.asarray(q)
    squeeze = q.ndim == 3
    qb = q[None] if squeeze else q
    x_std = np.asarray(x_std, dtype='float32').reshape(1, 2, 1, 1)
    y_std = np.asarray(y_std, dtype='float32').reshape(1, 2, 1, 1)
    X = qb.astype('float32') / x_std                       # x_scale.normalize(m.q.astype('float32'))
    Xt = torch.as_tensor(X)
    if kind in ('gan', 'vae'):
        z = torch.as_tensor(np.asarray(noise, dtype='float32').reshape(qb.shape))
        Y = andrew_cnn_forward(nets[0], torch.cat([Xt, z], dim=1)).numpy()      # generate(): cat + G
        out = (Y * y_std).astype('float64')
    elif kind == 'gz':
        mean = andrew_cnn_forward(nets[0], Xt).numpy()
        var = andrew_cnn_forward(nets[1], Xt, final_softplus=True).numpy()
        z = np.asarray(noise, dtype='float64').reshape(qb.shape)
        out = ((mean + z * var ** 0.5) * y_std).astype('float64')             # mean_var_model.py:105-109
    elif kind == 'ols':
        out = (andrew_cnn_forward(nets[0], Xt).numpy() * y_std).astype('float64')
    else:
        raise ValueError(kind)
    return out[0] if squeeze else out


def demean(x):
    """models/parameterization.py:25 : subtract the per-layer spatial mean."""
    return x - x.mean(axis=(-2, -1), keepdims=True)


class AR1Sampler(object):
    """tools/stochastic_pyqg.py:30-54."""

    def __init__(self, nsteps):
        self.nsteps = nsteps

    def update(self, generate_noise):
        if hasattr(self, 'noise'):
            if self.nsteps > 0:
                a = 1 - 1 / self.nsteps
                b = (1 / self.nsteps * (2 - 1 / self.nsteps)) ** 0.5
            else:
                a, b = 1, 0
            self.noise = a * self.noise + b * generate_noise()
        else:
            self.noise = generate_noise()
        return True


class ConstantSampler(object):
    """tools/stochastic_pyqg.py:56-72."""

    def __init__(self, nsteps):
        self.nsteps = nsteps

    def update(self, generate_noise):
        compute = True
        if hasattr(self, 'noise'):
            if self.counter % self.nsteps == 0:
                self.noise = generate_noise()
                self.counter = 1
            else:
                self.counter += 1
                compute = False
        else:
            self.noise = generate_noise()
            self.counter = 1
        return compute


def random_state_dict(n_in, n_out, seed=0, hidden=(128, 64, 32, 32, 32, 32, 32), kernels=(5, 5, 3, 3, 3, 3, 3, 3),
                      realistic_bn=True):
    """Random-init AndrewCNN state dict with the reference's key layout (SURVEY.md §8d synthetic weights).

    ``weights_init`` (cnn_tools.py:54-65) draws N(0,0.02) conv weights and BN gamma ~ N(1,0.02); with such small
    weights the activations collapse towards the biases after a few layers, so by default conv weights use the
    torch default (kaiming-uniform) scale and BN running stats are non-trivial, which exercises the arithmetic harder.
    """
    g = torch.Generator().manual_seed(seed)
    chans = [n_in] + list(hidden) + [n_out]
    sd = {}
    for n in range(8):
        cin, cout, k = chans[n], chans[n + 1], kernels[n]
        bound = 1.0 / np.sqrt(cin * k * k)
        sd['conv.%d.weight' % (3 * n)] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound * (3 ** 0.5)
        sd['conv.%d.bias' % (3 * n)] = (torch.rand(cout, generator=g) * 2 - 1) * bound
        if n < 7:
            j = 3 * n + 2
            sd['conv.%d.weight' % j] = 1 + 0.1 * torch.randn(cout, generator=g)
            sd['conv.%d.bias' % j] = 0.1 * torch.randn(cout, generator=g)
            if realistic_bn:
                sd['conv.%d.running_mean' % j] = 0.2 + 0.1 * torch.randn(cout, generator=g)
                sd['conv.%d.running_var' % j] = 0.2 + 0.3 * torch.rand(cout, generator=g)
            else:
                sd['conv.%d.running_mean' % j] = torch.zeros(cout)
                sd['conv.%d.running_var' % j] = torch.ones(cout)
            sd['conv.%d.num_batches_tracked' % j] = torch.tensor(0)
    return sd
