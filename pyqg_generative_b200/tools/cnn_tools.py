"""Host-side mirror of the inference half of ``pyqg_generative/tools/cnn_tools.py``.

Same names and argument meaning as the reference (AndrewCNN :125-182, ChannelwiseScaler :502-553,
minibatch :607-622, apply_function :702-735) but the forward pass runs in libqgb200's sm_100a kernels
(fp32 FFMA direct convolution or the tcgen05 implicit-GEMM path), not in torch.nn.
Training utilities (train, DCGAN_discriminator, ...) are out of scope.
"""
import ctypes
import json

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

    def apply(self, fn):
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
    """Reference cnn_tools.py:54-65 re-initialises weights for training; inference-only here: no-op."""
    return None


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


def extract(ds, key):
    """cnn_tools.py:398-400 for xarray or dict-like datasets: (run,time,lev,y,x) -> (run*time, lev, y, x)."""
    v = ds[key]
    v = np.asarray(getattr(v, 'values', v))
    return v.reshape((-1,) + v.shape[2:])
