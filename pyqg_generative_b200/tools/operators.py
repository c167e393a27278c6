"""Coarse-graining operators of ``pyqg_generative/tools/operators.py`` on the GPU.

``cut_off`` (:117-132), ``Operator1`` = model_filter o cut_off (:204-205), ``Operator2`` = gauss_filter(., nc//2) o cut_off
(:207-208), ``Operator5`` = cut_off (:216-217) and ``PV_subgrid_forcing`` (:283-287, dealias='none') keep the reference
signatures.  Inputs may be numpy arrays (2-D field, (nlev,ny,nx) like the reference's ``array_format`` numpy branch, or
any leading batch axes) or CUDA torch tensors (processed in place on the device, a CUDA tensor is returned).
All FFTs / truncations / Jacobians run in libqgb200 (qgb_operator, qgb_subgrid_forcing).
``fft_interpolate`` (:134-190), ``Operator4`` = model_filter o Operator2 (:213-214) and the dealiased forcings (``advect``
:253-266, '2/3-rule' and '3/2-rule') are built as well.  Not built: Operator3 (gcm_filters is not a dependency here).
"""
import ctypes

import numpy as np

from .. import _lib


def _is_cuda_tensor(x):
    try:
        import torch
        return isinstance(x, torch.Tensor) and x.is_cuda
    except ImportError:
        return False


def _run_operator(op, X, nc):
    import torch
    if nc is None:
        raise ValueError('nc must be given')
    lib = _lib.load()
    if _is_cuda_tensor(X):
        x = X.to(torch.float64).contiguous()
        if x.dim() < 2 or x.shape[-1] != x.shape[-2]:
            raise ValueError('numpy array should be 2 or 3 dimensional')
        n = x.shape[-1]
        batch = int(np.prod(x.shape[:-2])) if x.dim() > 2 else 1
        out = torch.empty(tuple(x.shape[:-2]) + (nc, nc), dtype=torch.float64, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(lib.qgb_operator(x.device.index, op, n, nc, batch, x.data_ptr(), out.data_ptr(), 1, stream))
        return out
    if not torch.cuda.is_available():
        raise RuntimeError('coarse-graining operators need a CUDA device: libqgb200 has no CPU fallback')
    x = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
    if x.ndim < 2:
        raise ValueError('numpy array should be 2 or 3 dimensional')
    if x.shape[-1] != x.shape[-2]:
        raise ValueError('only square fields are supported')
    n = x.shape[-1]
    batch = int(np.prod(x.shape[:-2])) if x.ndim > 2 else 1
    out = np.empty(x.shape[:-2] + (nc, nc))
    dev = torch.cuda.current_device()
    _lib.check(lib.qgb_operator(dev, op, n, nc, batch, x.ctypes.data, out.ctypes.data, 0,
                                torch.cuda.current_stream(dev).cuda_stream))
    return out


def cut_off(X, nc):
    return _run_operator(5, X, nc)


def Operator1(X, nc):
    return _run_operator(1, X, nc)


def Operator2(X, nc):
    return _run_operator(2, X, nc)


def Operator4(X, nc):
    return _run_operator(4, X, nc)


def Operator5(X, nc):
    return _run_operator(5, X, nc)


def fft_interpolate(x, n, N, truncate_2h=True):
    """Reference :134-190: spectral interpolation n -> N (either direction) of 2-D / 3-D (or batched) real fields."""
    import torch
    if not truncate_2h:
        raise NotImplementedError('truncate_2h=False is not built')
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    if x.shape[-2] != n or x.shape[-1] != n:
        raise ValueError('Input variable must be n*n points')
    if n % 2 != 0 or N % 2 != 0:
        raise ValueError('Grid sizes (n,N) must be even')
    if not torch.cuda.is_available():
        raise RuntimeError('fft_interpolate needs a CUDA device: libqgb200 has no CPU fallback')
    batch = int(np.prod(x.shape[:-2])) if x.ndim > 2 else 1
    out = np.empty(x.shape[:-2] + (N, N))
    dev = torch.cuda.current_device()
    _lib.check(_lib.load().qgb_fft_interpolate(dev, n, N, batch, x.ctypes.data, out.ctypes.data, 0,
                                               torch.cuda.current_stream(dev).cuda_stream))
    return out


_OP_ID = {'Operator1': 1, 'Operator2': 2, 'Operator4': 4, 'Operator5': 5, 'cut_off': 5}
_DEALIAS_ID = {'none': 0, '2/3-rule': 1, '3/2-rule': 2}


def _config(pyqg_params, n):
    cfg = _lib.QgbConfig()
    _lib.load().qgb_default_config(ctypes.byref(cfg))
    cfg.nx = n
    for k in ('L', 'dt', 'rek', 'filterfac', 'beta', 'rd', 'delta', 'H1', 'U1', 'U2'):
        if k in pyqg_params:
            setattr(cfg, k, float(pyqg_params[k]))
    return cfg


def PV_subgrid_forcing(q, nc, operator, pyqg_params, dealias='none', return_fields=False):
    """Reference :283-287.  ``q``: (2,n,n) or batched (B,2,n,n), numpy or CUDA tensor.

    Returns ``(forcing, mf, m)`` like the reference when ``return_fields`` is False -- ``mf`` is a light object carrying
    q, u, v, p of the coarse model, ``m`` is None (the fine model is never materialised) -- or ``(forcing, dict)`` with
    the coarse fields when ``return_fields`` is True."""
    import torch
    if dealias not in _DEALIAS_ID:
        raise ValueError('dealias should be none or 2/3-rule or 3/2-rule')
    dealias_id = _DEALIAS_ID[dealias]
    op = _OP_ID.get(getattr(operator, '__name__', str(operator)))
    if op is None:
        raise NotImplementedError('operator %r is not on the accelerated path (Operator1, Operator2, Operator4, Operator5)' % (operator,))
    lib = _lib.load()
    cuda_in = _is_cuda_tensor(q)
    if cuda_in:
        x = q.to(torch.float64).contiguous()
        single = x.dim() == 3
        x = x.unsqueeze(0) if single else x
        B, n = x.shape[0], x.shape[-1]
        outs = [torch.empty((B, 2, nc, nc), dtype=torch.float64, device=x.device) for _ in range(5)]
        ptrs = [o.data_ptr() for o in outs]
        dev, src, on_dev = x.device.index, x.data_ptr(), 1
    else:
        if not torch.cuda.is_available():
            raise RuntimeError('PV_subgrid_forcing needs a CUDA device: libqgb200 has no CPU fallback')
        x = np.ascontiguousarray(np.asarray(q, dtype=np.float64))
        single = x.ndim == 3
        x = x[None] if single else x
        B, n = x.shape[0], x.shape[-1]
        outs = [np.empty((B, 2, nc, nc)) for _ in range(5)]
        ptrs = [o.ctypes.data for o in outs]
        dev, src, on_dev = torch.cuda.current_device(), x.ctypes.data, 0
    cfg = _config(pyqg_params, n)
    cfg.device = dev
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.qgb_subgrid_forcing(ctypes.byref(cfg), op, nc, dealias_id, B, src, ptrs[0], ptrs[1], ptrs[2], ptrs[3], ptrs[4],
                                       on_dev, stream))
    if cuda_in:
        outs = [o.cpu().numpy() for o in outs]
    if single:
        outs = [o[0] for o in outs]
    forcing, qf, uf, vf, pf = outs
    fields = dict(q=qf, u=uf, v=vf, psi=pf)
    if return_fields:
        return forcing, fields
    mf = type('CoarseModel', (), dict(q=qf, u=uf, v=vf, p=pf, nx=nc, ny=nc))()
    return forcing, mf, None
