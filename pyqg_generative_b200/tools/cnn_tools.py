"""Host-side mirror of ``pyqg_generative/tools/cnn_tools.py`` for the regression networks.

Same names and argument meaning as the reference (AndrewCNN :125-182, ChannelwiseScaler :502-553,
minibatch :607-622, apply_function :702-735, train :645-700, evaluate_test :624-643, prepare_PV_data :402-421) but the
forward pass runs in libqgb200's sm_100a kernels (fp32 FFMA direct convolution or the tcgen05 implicit-GEMM path) and the
training step (forward with batch statistics, MSE, backward, Adam) in its device-side trainer (csrc/train.cuh), not in torch.nn.
The adversarial / variational trainers (DCGAN_discriminator, WGAN-GP, ELBO) are not built.
"""
import collections
import ctypes
import json
from time import time

import numpy as np
import torch

from .. import _lib

BN_EPS = 1e-5


def _engine_handle(device_index):
    """A tiny model handle (nx=16, 1 member) that only hosts network weights for raw forwards."""
    lib = _lib.load()
    cfg = _lib.QgbConfig()
    lib.qgb_default_config(ctypes.byref(cfg))
    cfg.nx, cfg.members, cfg.device = 16, 1, device_index
    h = ctypes.c_void_p()
    _lib.check(lib.qgb_create(ctypes.byref(cfg), ctypes.byref(h)))
    return h


class AndrewCNN(object):
    """Inference-only AndrewCNN: 8 x [Conv2d circular 'same' -> ReLU -> BatchNorm2d], last block bare conv.

    Constructor signature and ``state_dict`` key layout follow the reference (cnn_tools.py:125-160) so the shipped
    ``*.pt`` files load unchanged.  ``forward`` accepts a torch tensor (B, n_in, ny, nx) float32 on CPU or CUDA and
    returns a tensor on the same device.
    """

    def __init__(self, n_in, n_out, ReLU='ReLU', batch_norm=True, bias=True, final_activation='None', div=False,
                 hidden_channels=[128, 64, 32, 32, 32, 32, 32], kernels=[5, 5, 3, 3, 3, 3, 3, 3], precision='fp32'):
        if div:
            raise NotImplementedError('div=True (spectral divergence head) is out of scope: every shipped '
                                      'model_args.json has div=false')
        if ReLU != 'ReLU':
            raise NotImplementedError("only ReLU='ReLU' is supported")
        if not bias:
            raise NotImplementedError('bias=False is not supported')
        if final_activation != 'None':
            raise NotImplementedError('final_activation is not supported (VarCNN applies softplus itself)')
        self.div = div
        self.n_in, self.n_out = n_in, n_out
        self.batch_norm = batch_norm
        self.precision = precision
        chans = [n_in] + list(hidden_channels) + [n_out]
        self._blocks = []   # (conv_index, bn_index or None, cin, cout, k)
        idx = 0
        for n in range(len(chans) - 1):
            last = n == len(chans) - 2
            conv_idx = idx
            idx += 1
            bn_idx = None
            if not last:
                idx += 1            # ReLU
                if batch_norm:
                    bn_idx = idx
                    idx += 1
            self._blocks.append((conv_idx, bn_idx, chans[n], chans[n + 1], kernels[n]))
        g = torch.Generator().manual_seed(0)
        self._sd = {}
        for conv_idx, bn_idx, cin, cout, k in self._blocks:
            bound = 1.0 / np.sqrt(cin * k * k)
            self._sd['conv.%d.weight' % conv_idx] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
            self._sd['conv.%d.bias' % conv_idx] = (torch.rand(cout, generator=g) * 2 - 1) * bound
            if bn_idx is not None:
                self._sd['conv.%d.weight' % bn_idx] = torch.ones(cout)
                self._sd['conv.%d.bias' % bn_idx] = torch.zeros(cout)
                self._sd['conv.%d.running_mean' % bn_idx] = torch.zeros(cout)
                self._sd['conv.%d.running_var' % bn_idx] = torch.ones(cout)
                self._sd['conv.%d.num_batches_tracked' % bn_idx] = torch.tensor(0)
        self._engines = {}   # device index -> handle with these weights loaded
        self._keep = None

    # ---- torch.nn.Module look-alike surface used by the reference call sites --------------------------------
    def state_dict(self):
        return dict(self._sd)

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self._sd if k not in sd]
        unexpected = [k for k in sd if k not in self._sd]
        if strict and (missing or unexpected):
            raise RuntimeError('Error(s) in loading state_dict for AndrewCNN: missing %s, unexpected %s'
                               % (missing, unexpected))
        for k in self._sd:
            if k in sd:
                v = torch.as_tensor(sd[k]).detach().cpu()
                if tuple(v.shape) != tuple(self._sd[k].shape):
                    raise RuntimeError('size mismatch for %s: %s vs %s' % (k, tuple(v.shape), tuple(self._sd[k].shape)))
                self._sd[k] = v.clone()
        self._release()
        return '<All keys matched successfully>'

    def to(self, device):
        return self

    def eval(self):
        return self

    def train(self, mode=True):
        return self

    softplus_output = False      # VarCNN overrides: softplus on the output, in forward() and in the loss

    def parameter_names(self):
        """state_dict keys in ``net.parameters()`` order (= the flat parameter layout of qgb_train_*)."""
        names = []
        for conv_idx, bn_idx, cin, cout, k in self._blocks:
            names += ['conv.%d.weight' % conv_idx, 'conv.%d.bias' % conv_idx]
            if bn_idx is not None:
                names += ['conv.%d.weight' % bn_idx, 'conv.%d.bias' % bn_idx]
        return names

    def buffer_names(self):
        names = []
        for conv_idx, bn_idx, cin, cout, k in self._blocks:
            if bn_idx is not None:
                names += ['conv.%d.running_mean' % bn_idx, 'conv.%d.running_var' % bn_idx]
        return names

    def compute_loss(self, x, ytrue):
        """cnn_tools.py:177-182 ``{'loss': MSELoss()(self.forward(x), ytrue)}`` (eval-mode forward, no gradient)."""
        y = self.forward(x, softplus=self.softplus_output)
        yt = torch.as_tensor(ytrue).to(y.device, torch.float32)
        return {'loss': float(((y - yt) ** 2).mean())}

    def apply(self, fn):
        """``net.apply(weights_init)`` of the reference's constructors (models/cgan_regression.py:62-63)."""
        fn(self)
        return self

    # ---- weight export ----------------------------------------------------------------------------------------
    def layers(self):
        """List of dicts (cin, cout, ksize, relu_bn, weight, bias, bn_scale, bn_shift) with float32 numpy arrays.
        BatchNorm (eval) is folded as ATen's CPU kernel does: alpha = weight/sqrt(var+eps), beta = bias - mean*alpha."""
        out = []
        for n, (conv_idx, bn_idx, cin, cout, k) in enumerate(self._blocks):
            last = n == len(self._blocks) - 1
            w = np.ascontiguousarray(self._sd['conv.%d.weight' % conv_idx].numpy().astype('float32'))
            b = np.ascontiguousarray(self._sd['conv.%d.bias' % conv_idx].numpy().astype('float32'))
            if last:
                s = t = None
            elif bn_idx is not None:
                gamma = self._sd['conv.%d.weight' % bn_idx].numpy().astype('float32')
                beta = self._sd['conv.%d.bias' % bn_idx].numpy().astype('float32')
                mean = self._sd['conv.%d.running_mean' % bn_idx].numpy().astype('float32')
                var = self._sd['conv.%d.running_var' % bn_idx].numpy().astype('float32')
                invstd = np.float32(1.0) / np.sqrt(var + np.float32(BN_EPS))
                s = np.ascontiguousarray((invstd * gamma).astype('float32'))
                t = np.ascontiguousarray((beta - mean * s).astype('float32'))
            else:
                s = np.ones(cout, 'float32')
                t = np.zeros(cout, 'float32')
            out.append(dict(cin=cin, cout=cout, ksize=k, relu_bn=0 if last else 1, weight=w, bias=b,
                            bn_scale=s, bn_shift=t))
        return out

    def c_layers(self):
        """ctypes array of qgb_cnn_layer + the numpy arrays that must stay alive during the call."""
        ls = self.layers()
        arr = (_lib.QgbCnnLayer * len(ls))()
        keep = []
        for i, L in enumerate(ls):
            arr[i].cin, arr[i].cout, arr[i].ksize, arr[i].relu_bn = L['cin'], L['cout'], L['ksize'], L['relu_bn']
            for name in ('weight', 'bias', 'bn_scale', 'bn_shift'):
                a = L[name]
                if a is not None:
                    keep.append(a)
                    setattr(arr[i], name, a.ctypes.data)
                else:
                    setattr(arr[i], name, None)
        return arr, keep

    # ---- forward -------------------------------------------------------------------------------------------------
    def _release(self):
        lib = _lib.load() if self._engines else None
        for h in self._engines.values():
            lib.qgb_destroy(h)
        self._engines = {}

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _engine(self, dev):
        if dev not in self._engines:
            h = _engine_handle(dev)
            arr, keep = self.c_layers()
            _lib.check(_lib.load().qgb_cnn_load(h, _lib.CLOSURE_RAW, 0, len(arr), arr), h)
            self._engines[dev] = h
        return self._engines[dev]

    def forward(self, x, softplus=False, precision=None):
        if not torch.cuda.is_available():
            raise RuntimeError('AndrewCNN.forward needs a CUDA device: libqgb200 has no CPU fallback')
        x = torch.as_tensor(x)
        if x.dim() != 4 or x.shape[1] != self.n_in:
            raise ValueError('expected input of shape (B, %d, ny, nx), got %s' % (self.n_in, tuple(x.shape)))
        src_device = x.device
        dev = x.device.index if x.is_cuda else torch.cuda.current_device()
        if dev is None:
            dev = torch.cuda.current_device()
        xd = x.to(device='cuda:%d' % dev, dtype=torch.float32).contiguous()
        y = torch.empty((x.shape[0], self.n_out, x.shape[2], x.shape[3]), dtype=torch.float32, device=xd.device)
        prec = _lib.PRECISIONS[precision or self.precision]
        h = self._engine(dev)
        stream = torch.cuda.current_stream(xd.device).cuda_stream
        _lib.check(_lib.load().qgb_cnn_forward(h, 0, xd.data_ptr(), y.data_ptr(), x.shape[0], x.shape[2], x.shape[3],
                                                1 if softplus else 0, prec, 1, stream), h)
        return y.to(src_device)

    __call__ = forward


def weights_init(m):
    """Reference cnn_tools.py:54-65 (DCGAN initialisation, applied by the GAN / VAE constructors): conv weights N(0, 0.02),
    BatchNorm weights N(1, 0.02) and zero bias.  Accepts our AndrewCNN; other objects are left alone."""
    if isinstance(m, DCGAN_discriminator):
        for key in m.KEYS:
            m._sd[key] = torch.randn(tuple(m._sd[key].shape)) * 0.02
        m._release()
        return None
    if not isinstance(m, AndrewCNN):
        return None
    for conv_idx, bn_idx, cin, cout, k in m._blocks:
        m._sd['conv.%d.weight' % conv_idx] = torch.randn(cout, cin, k, k) * 0.02
        if bn_idx is not None:
            m._sd['conv.%d.weight' % bn_idx] = 1.0 + torch.randn(cout) * 0.02
            m._sd['conv.%d.bias' % bn_idx] = torch.zeros(cout)
    m._release()
    return None


class DCGAN_discriminator(object):
    """cnn_tools.py:212-244 with ``bn='None'`` (what CGANRegression builds, models/cgan_regression.py:57): four
    Conv2d(4, stride 2, padding 1, bias=False) + LeakyReLU(0.2) from ``in_channels`` to ndf, 2 ndf, 4 ndf, 8 ndf, then
    Conv2d(8 ndf, 1, nx/64*4, 1, 0).  State-dict keys as in the reference's nn.Sequential (0, 2, 5, 8, 11 .weight).
    ``forward`` runs on the device through ``qgb_disc_forward``; training goes through ``CGANTrainer``."""
    KEYS = ['0.weight', '2.weight', '5.weight', '8.weight', '11.weight']

    def __init__(self, in_channels, ndf=64, nx=64, bn='None'):
        if bn != 'None':
            raise NotImplementedError("only bn='None' (the CGAN's discriminator) is built")
        self.in_channels, self.ndf, self.nx = int(in_channels), int(ndf), int(nx)
        k5 = int(nx / 64 * 4)
        chans = [in_channels, ndf, ndf * 2, ndf * 4, ndf * 8]
        g = torch.Generator().manual_seed(1)
        self._sd = {}
        for i, key in enumerate(self.KEYS):
            cin, cout, k = (chans[i], chans[i + 1], 4) if i < 4 else (chans[4], 1, k5)
            bound = 1.0 / np.sqrt(cin * k * k)
            self._sd[key] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
        self._disc = None

    def state_dict(self):
        return dict(self._sd)

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self._sd if k not in sd]
        unexpected = [k for k in sd if k not in self._sd]
        if strict and (missing or unexpected):
            raise RuntimeError('Error(s) in loading state_dict for DCGAN_discriminator: missing %s, unexpected %s'
                               % (missing, unexpected))
        for k in self._sd:
            if k in sd:
                v = torch.as_tensor(sd[k]).detach().cpu().float()
                if tuple(v.shape) != tuple(self._sd[k].shape):
                    raise RuntimeError('size mismatch for %s: %s vs %s' % (k, tuple(v.shape), tuple(self._sd[k].shape)))
                self._sd[k] = v.clone()
        self._release()
        return '<All keys matched successfully>'

    def to(self, device):
        return self

    def train(self, mode=True):
        return self

    def eval(self):
        return self

    def zero_grad(self):
        return None

    def apply(self, fn):
        fn(self)
        return self

    def flat(self):
        return np.ascontiguousarray(np.concatenate([self._sd[k].numpy().astype('float32').ravel() for k in self.KEYS]))

    def unflat(self, flat):
        out, o = {}, 0
        for k in self.KEYS:
            shape = tuple(self._sd[k].shape)
            n = int(np.prod(shape))
            out[k] = flat[o:o + n].reshape(shape).copy()
            o += n
        return out

    def _release(self):
        if getattr(self, '_disc', None) is not None:
            self._disc.close()
            self._disc = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def forward(self, x):
        """x: (B, in_channels, nx, nx) float32 tensor (CPU or CUDA) -> (B, 1, 1, 1) on the same device."""
        xt = torch.as_tensor(x, dtype=torch.float32)
        if self._disc is None or self._disc.max_batch * 4 < xt.shape[0]:
            self._release()
            self._disc = DiscState(self, max_batch=max(16, (int(xt.shape[0]) + 3) // 4))
        out = self._disc.forward(xt)
        return out.reshape(-1, 1, 1, 1)

    __call__ = forward


class DiscState(object):
    """Device-side state of a DCGAN_discriminator (``qgb_disc``): parameters, Adam moments, activations of 4 x max_batch
    samples."""

    def __init__(self, net, max_batch=64, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError('the discriminator needs a CUDA device: libqgb200 has no CPU fallback')
        self._lib = _lib.load()
        self.net = net
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.max_batch = int(max_batch)
        self._h = ctypes.c_void_p()
        _lib.check_disc(self._lib.qgb_disc_create(self.device, net.in_channels, net.ndf, net.nx, self.max_batch,
                                                  ctypes.byref(self._h)))
        self.nparams = int(self._lib.qgb_disc_num_params(self._h))
        self.upload(reset_optimizer=True)

    def close(self):
        if getattr(self, '_h', None):
            self._lib.qgb_disc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, reset_optimizer=False):
        p = self.net.flat()
        assert p.size == self.nparams
        _lib.check_disc(self._lib.qgb_disc_set_params(self._h, p.ctypes.data, 1 if reset_optimizer else 0), self._h)

    def sync_to(self, net=None):
        net = net or self.net
        p = np.empty(self.nparams, 'float32')
        _lib.check_disc(self._lib.qgb_disc_get_params(self._h, p.ctypes.data, None), self._h)
        for k, a in net.unflat(p).items():
            net._sd[k] = torch.from_numpy(a)
        if getattr(net, '_disc', None) is not self:
            net._release()                 # the inference-side copy (if any) is stale now

    def last_grads(self):
        g = np.empty(self.nparams, 'float32')
        _lib.check_disc(self._lib.qgb_disc_get_params(self._h, None, g.ctypes.data), self._h)
        return self.net.unflat(g)

    def forward(self, x):
        dev = torch.device('cuda:%d' % self.device)
        xd = x.to(dev).contiguous()
        out = torch.empty(xd.shape[0], dtype=torch.float32, device=dev)
        _lib.check_disc(self._lib.qgb_disc_forward(self._h, xd.data_ptr(), xd.shape[0], 1, out.data_ptr(),
                                                   torch.cuda.current_stream(dev).cuda_stream), self._h)
        return out.to(x.device)

    def launch_count(self):
        return int(self._lib.qgb_disc_launch_count(self._h))


class Trainer(object):
    """Device-side training state of one AndrewCNN (``qgb_trainer``): parameters, Adam moments, BatchNorm running statistics.

    ``step(x, y, lr)`` is one iteration of the loop at cnn_tools.py:685-690; ``grads`` exposes the raw gradients for the
    parity tests; ``sync_to(net)`` writes parameters and buffers back into the network's state_dict."""

    def __init__(self, net, ny, nx, max_batch=64, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError('training needs a CUDA device: libqgb200 has no CPU fallback')
        if not net.batch_norm:
            raise NotImplementedError('the trainer expects ReLU + BatchNorm2d after every layer but the last')
        self._lib = _lib.load()
        self.net = net
        self.device = torch.cuda.current_device() if device is None else int(device)
        chans = [net._blocks[0][2]] + [b[3] for b in net._blocks]
        ks = [b[4] for b in net._blocks]
        self.ny, self.nx, self.max_batch = int(ny), int(nx), int(max_batch)
        self._h = ctypes.c_void_p()
        ca = (ctypes.c_int32 * len(chans))(*chans)
        ka = (ctypes.c_int32 * len(ks))(*ks)
        _lib.check_train(self._lib.qgb_train_create(self.device, len(ks), ca, ka, self.ny, self.nx, self.max_batch,
                                                    1 if net.softplus_output else 0, ctypes.byref(self._h)))
        self.nparams = int(self._lib.qgb_train_num_params(self._h))
        self.nbuffers = int(self._lib.qgb_train_num_buffers(self._h))
        self.steps = 0
        self.upload(reset_optimizer=True)

    def close(self):
        if getattr(self, '_h', None):
            self._lib.qgb_train_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _flat(self, names):
        sd = self.net._sd
        return np.ascontiguousarray(np.concatenate([sd[n].numpy().astype('float32').ravel() for n in names])) if names \
            else np.zeros(0, 'float32')

    def upload(self, reset_optimizer=False):
        p, b = self._flat(self.net.parameter_names()), self._flat(self.net.buffer_names())
        assert p.size == self.nparams and b.size == self.nbuffers
        _lib.check_train(self._lib.qgb_train_set_params(self._h, p.ctypes.data, b.ctypes.data if b.size else None,
                                                        1 if reset_optimizer else 0), self._h)

    def _unflat(self, flat, names):
        out, o = {}, 0
        for n in names:
            shape = tuple(self.net._sd[n].shape)
            sz = int(np.prod(shape))
            out[n] = flat[o:o + sz].reshape(shape).copy()
            o += sz
        return out

    def sync_to(self, net=None):
        """Parameters, running statistics (and num_batches_tracked) -> the network's state_dict."""
        net = net or self.net
        p, b = np.empty(self.nparams, 'float32'), np.empty(max(self.nbuffers, 1), 'float32')
        _lib.check_train(self._lib.qgb_train_get_params(self._h, p.ctypes.data, b.ctypes.data), self._h)
        for n, a in self._unflat(p, net.parameter_names()).items():
            net._sd[n] = torch.from_numpy(a)
        for n, a in self._unflat(b, net.buffer_names()).items():
            net._sd[n] = torch.from_numpy(a)
        for conv_idx, bn_idx, cin, cout, k in net._blocks:
            if bn_idx is not None:
                key = 'conv.%d.num_batches_tracked' % bn_idx
                net._sd[key] = torch.tensor(int(net._sd[key]) + self.steps)
        self.steps = 0
        net._release()

    def _xy(self, x, y):
        x = np.ascontiguousarray(np.asarray(x, dtype='float32'))
        y = np.ascontiguousarray(np.asarray(y, dtype='float32'))
        if x.ndim != 4 or x.shape[2:] != (self.ny, self.nx) or y.shape[0] != x.shape[0] or y.shape[2:] != x.shape[2:]:
            raise ValueError('expected (B, C, %d, %d) minibatches, got %s and %s' % (self.ny, self.nx, x.shape, y.shape))
        return x, y

    def step(self, x, y, lr):
        x, y = self._xy(x, y)
        loss = ctypes.c_double(0.0)
        _lib.check_train(self._lib.qgb_train_step(self._h, x.ctypes.data, y.ctypes.data, x.shape[0], 0, float(lr),
                                                  ctypes.byref(loss), None), self._h)
        self.steps += 1
        return loss.value

    def grads(self, x, y, update_running=False):
        x, y = self._xy(x, y)
        g = np.empty(self.nparams, 'float32')
        loss = ctypes.c_double(0.0)
        _lib.check_train(self._lib.qgb_train_grads(self._h, x.ctypes.data, y.ctypes.data, x.shape[0], 0, g.ctypes.data,
                                                   ctypes.byref(loss), 1 if update_running else 0, None), self._h)
        return self._unflat(g, self.net.parameter_names()), loss.value

    def eval_loss(self, x, y):
        x, y = self._xy(x, y)
        loss = ctypes.c_double(0.0)
        _lib.check_train(self._lib.qgb_train_eval_loss(self._h, x.ctypes.data, y.ctypes.data, x.shape[0], 0,
                                                       ctypes.byref(loss), None), self._h)
        return loss.value

    def last_grads(self):
        """Gradients left by the last backward pass (``qgb_train_get_grads``), keyed like ``net.named_parameters()``."""
        g = np.empty(self.nparams, 'float32')
        _lib.check_train(self._lib.qgb_train_get_grads(self._h, g.ctypes.data), self._h)
        return self._unflat(g, self.net.parameter_names())

    def set_adam(self, beta1, beta2):
        _lib.check_train(self._lib.qgb_train_set_adam(self._h, float(beta1), float(beta2)), self._h)

    def launch_count(self):
        return int(self._lib.qgb_train_launch_count(self._h))


def multistep_lr(learning_rate, num_epochs, epoch, gamma=0.1):
    """Learning rate of ``epoch`` (0-based) under MultiStepLR(milestones=[E/2, 3E/4, 7E/8], gamma) stepped once per
    epoch (cnn_tools.py:672-673,691; gamma = 0.5 in train_CGAN, cgan_regression.py:248-251); coinciding milestones multiply,
    as torch's Counter-based scheduler does."""
    counts = collections.Counter([int(num_epochs / 2), int(num_epochs * 3 / 4), int(num_epochs * 7 / 8)])
    return learning_rate * gamma ** sum(c for m, c in counts.items() if 0 < m <= epoch)


class AverageLoss(object):
    """cnn_tools.py:555-600: sample-weighted epoch means of a dict of losses, appended to ``log_dict`` lists."""

    def __init__(self, log_dict):
        self.init_me = True
        self.count = {}

    def accumulate(self, log_dict, losses, n):
        if self.init_me:
            for key in losses:
                log_dict.setdefault(key, [])
                self.count[key] = 0
                log_dict[key].append(0.)
            self.init_me = False
        for key, value in losses.items():
            log_dict[key][-1] += float(value) * n
            self.count[key] += n

    def average(self, log_dict):
        for key in self.count:
            log_dict[key][-1] = log_dict[key][-1] / self.count[key]


def evaluate_test(net, *arrays, batch_size=64, postfix='_test', device=None, trainer=None):
    """cnn_tools.py:624-643: epoch-mean eval-mode loss appended to ``net.log_dict['loss' + postfix]``."""
    if not hasattr(net, 'log_dict'):
        net.log_dict = {}
    tot, cnt = 0.0, 0
    for x, y in minibatch(*arrays, batch_size=batch_size):
        loss = trainer.eval_loss(x.numpy(), y.numpy()) if trainer is not None else net.compute_loss(x, y)['loss']
        tot += loss * len(x)
        cnt += len(x)
    net.log_dict.setdefault('loss' + postfix, []).append(tot / max(cnt, 1))


def train(net, X_train, Y_train, X_test, Y_test, num_epochs, batch_size, learning_rate, device=None):
    """cnn_tools.py:645-700 on the device-side trainer: Adam(lr) + MultiStepLR, shuffled minibatches, MSE through
    ``compute_loss``, BatchNorm in training mode; epoch-mean losses in ``net.log_dict['loss' | 'loss_test']``."""
    X_train, Y_train = np.asarray(X_train), np.asarray(Y_train)
    print('Training starts on device %s, number of samples %d' % (torch.cuda.get_device_name(0) if torch.cuda.is_available()
                                                                  else 'cpu', len(X_train)))
    trainer = Trainer(net, X_train.shape[2], X_train.shape[3], max_batch=batch_size, device=device)
    if not hasattr(net, 'log_dict'):
        net.log_dict = {}
    t_s = time()
    for epoch in range(num_epochs):
        t_e = time()
        lr = multistep_lr(learning_rate, num_epochs, epoch)
        tot, cnt = 0.0, 0
        for x, y in minibatch(X_train, Y_train, batch_size=batch_size):
            loss = trainer.step(x.numpy(), y.numpy(), lr)
            tot += loss * len(x)
            cnt += len(x)
        net.log_dict.setdefault('loss', []).append(tot / max(cnt, 1))
        evaluate_test(net, X_test, Y_test, batch_size=batch_size, trainer=trainer)
        t = time()
        print('[%d/%d] [%.2f/%.2f] Loss: [%.3f, %.3f]' % (epoch + 1, num_epochs, t - t_e,
                                                          (t - t_s) * (num_epochs / (epoch + 1) - 1),
                                                          net.log_dict['loss'][-1], net.log_dict['loss_test'][-1]))
    trainer.sync_to(net)
    trainer.close()
    return net


class ChannelwiseScaler(object):
    """std / mean per channel (cnn_tools.py:502-553); json files are read and written in the reference format."""

    def __init__(self, X=None):
        if X is not None:
            X64 = np.asarray(X).astype('float64')
            self.mean = X64.mean(axis=(0, 2, 3), keepdims=True).astype('float32')
            self.std = X64.std(axis=(0, 2, 3), keepdims=True).astype('float32')

    def direct(self, X):
        return (X - self.mean) / self.std

    def inverse(self, X):
        return X * self.std + self.mean

    def normalize(self, X):
        return X / self.std

    def denormalize(self, X):
        return X * self.std

    def normalize_var(self, X):
        return X / (self.std ** 2)

    def denormalize_var(self, X):
        return X * (self.std ** 2)

    def write(self, name, folder='model'):
        to_str = lambda x: str(x.tolist())
        with open('%s/%s' % (folder, name), 'w') as f:
            json.dump(dict(mean=to_str(self.mean), std=to_str(self.std)), f)

    def read(self, name, folder='model'):
        with open('%s/%s' % (folder, name)) as f:
            d = json.load(f)
        self.std = np.array(json.loads(d['std'])).astype('float32')
        self.mean = np.array(json.loads(d['mean'])).astype('float32')
        return self


def minibatch(*arrays, batch_size=64, shuffle=True):
    """cnn_tools.py:607-622: yields tuples of torch tensors of at most ``batch_size`` rows."""
    assert len(set(len(a) for a in arrays)) == 1
    order = np.arange(len(arrays[0]))
    if shuffle:
        np.random.shuffle(order)
    steps = int(np.ceil(len(arrays[0]) / batch_size))
    for step in range(steps):
        idx = order[step * batch_size:(step + 1) * batch_size]
        yield tuple(torch.as_tensor(np.asarray(a)[idx]) for a in arrays)


def apply_function(net, *X, fun=None, batch_size=64, **kw):
    """cnn_tools.py:702-735: apply ``fun`` (default ``net.forward``) batch-wise on the GPU, return numpy array(s)."""
    if not torch.cuda.is_available():
        raise RuntimeError('apply_function needs a CUDA device: libqgb200 has no CPU fallback')
    device = torch.device('cuda:%d' % torch.cuda.current_device())
    if fun is None:
        fun = net.forward
    preds = []
    for x in minibatch(*X, batch_size=batch_size, shuffle=False):
        xx = [t.to(device) for t in x]
        y = fun(*xx, **kw)
        y = [y] if not isinstance(y, tuple) else y
        preds.append([yy.cpu().numpy() for yy in y])
    preds = [np.vstack(p) for p in zip(*preds)]
    return preds[0] if len(preds) == 1 else preds


def prepare_PV_data(ds_train, ds_test):
    """cnn_tools.py:402-421: q -> q_forcing_advection pairs, scaled by the per-channel std of the training set."""
    X_train, Y_train = extract(ds_train, 'q'), extract(ds_train, 'q_forcing_advection')
    X_test, Y_test = extract(ds_test, 'q'), extract(ds_test, 'q_forcing_advection')
    x_scale, y_scale = ChannelwiseScaler(X_train), ChannelwiseScaler(Y_train)
    return (x_scale.normalize(X_train), y_scale.normalize(Y_train), x_scale.normalize(X_test), y_scale.normalize(Y_test),
            x_scale, y_scale)


def write_log(log_dict, path):
    """cnn_tools.py:12-19 ``log_to_xarray(log_dict).to_netcdf(path)`` (and the ``stats.nc`` of the CVAE / CGAN trainers,
    models/cvae_regression.py:245-255) without xarray, in the layout of the shipped logs (Google-Colab/{GAN,VAE}/stats.nc,
    GZ/stats_var.nc: NetCDF-3 64-bit offset): coordinate ``epoch`` = 1..E (int32, long_name 'epoch'); one float64 variable over
    ``epoch`` per list-valued key; 2-D arrays (E, 2) as float32 over (``epoch``, ``lev``) with the coordinate ``lev`` = [1, 2]
    (long_name 'vertical levels'); python scalars as 0-d float64 variables.  Opens with xarray."""
    from scipy.io import netcdf_file
    series = {k: np.asarray(v) for k, v in log_dict.items()}
    n = max(a.shape[0] for a in series.values() if a.ndim >= 1)
    with netcdf_file(path, 'w', version=2) as f:
        f.createDimension('epoch', n)
        v = f.createVariable('epoch', 'i', ('epoch',))
        v[:] = np.arange(1, n + 1, dtype=np.int32)
        v.long_name = 'epoch'
        if any(a.ndim == 2 for a in series.values()):
            f.createDimension('lev', 2)
            v = f.createVariable('lev', 'i', ('lev',))
            v[:] = np.array([1, 2], dtype=np.int32)
            v.long_name = 'vertical levels'
        for k, a in series.items():
            if a.ndim == 2:
                v = f.createVariable(k, 'f', ('epoch', 'lev'))
                v[:] = a.astype(np.float32)
                v._FillValue = np.float32(np.nan)
            elif a.ndim == 1:
                v = f.createVariable(k, 'd', ('epoch',))
                v[:] = a.astype(np.float64)
                v._FillValue = np.float64(np.nan)
            else:
                v = f.createVariable(k, 'd', ())
                v.data[...] = float(a)
                v._FillValue = np.float64(np.nan)
    return path


def save_model_args(model, folder='model', **kw):
    """cnn_tools.py:21-25."""
    with open('%s/model_args.json' % folder, 'w') as f:
        json.dump(dict(model=model, **kw), f)


def extract(ds, key):
    """cnn_tools.py:398-400 for xarray or dict-like datasets: (run,time,lev,y,x) -> (run*time, lev, y, x)."""
    v = ds[key]
    v = np.asarray(getattr(v, 'values', v))
    return v.reshape((-1,) + v.shape[2:])
