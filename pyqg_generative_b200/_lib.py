"""ctypes binding of ``libqgb200.so`` (C ABI declared in include/qgb200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` / ``pyqg_generative_b200/build.py``.  There is no
CPU fallback: if the library is missing, or no CUDA device is present, every entry point raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libqgb200.so')

QGB_OK, QGB_EINVAL, QGB_ECUDA, QGB_ESTATE, QGB_EUNSUPPORTED = 0, -1, -2, -3, -4
(F_Q, F_QH, F_PH, F_U, F_V, F_DQHDT, F_FORCING, F_NOISE, F_P) = range(9)
CLOSURE_NONE, CLOSURE_GAN, CLOSURE_VAE, CLOSURE_GZ, CLOSURE_OLS, CLOSURE_RAW = range(6)
SAMPLER_AR1, SAMPLER_CONSTANT, SAMPLER_DETERMINISTIC = range(3)
PROF_SLOTS = 20
PROF_SLOT_NAMES = ['net0.L%d' % (i + 1) for i in range(8)] + ['net1.L%d' % (i + 1) for i in range(8)] + \
    ['spectral_step', 'latent_noise', 'closure_epilogue', 'diagnostics']
PREC_FP32, PREC_TC, PREC_TC_FAST, PREC_AUTO = 0, 1, 2, 3
PRECISIONS = {'fp32': PREC_FP32, 'tc': PREC_TC, 'tc_fast': PREC_TC_FAST, 'auto': PREC_AUTO}
PRECISION_NAMES = {v: k for k, v in PRECISIONS.items()}


class QgbConfig(ctypes.Structure):
    _fields_ = [('nx', ctypes.c_int32), ('members', ctypes.c_int32), ('member_offset', ctypes.c_int32),
                ('device', ctypes.c_int32)] + [(n, ctypes.c_double) for n in
                                               ('L', 'dt', 'rek', 'filterfac', 'beta', 'rd', 'delta', 'H1', 'U1', 'U2')]


class QgbCnnLayer(ctypes.Structure):
    _fields_ = [('cin', ctypes.c_int32), ('cout', ctypes.c_int32), ('ksize', ctypes.c_int32),
                ('relu_bn', ctypes.c_int32), ('weight', ctypes.c_void_p), ('bias', ctypes.c_void_p),
                ('bn_scale', ctypes.c_void_p), ('bn_shift', ctypes.c_void_p)]


# every symbol include/qgb200.h declares: name -> (restype, argtypes)
_vp, _i, _d = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
SYMBOLS = {
    'qgb_default_config': (None, [ctypes.POINTER(QgbConfig)]),
    'qgb_create': (_i, [ctypes.POINTER(QgbConfig), ctypes.POINTER(_vp)]),
    'qgb_destroy': (None, [_vp]),
    'qgb_last_error': (ctypes.c_char_p, [_vp]),
    'qgb_set_q': (_i, [_vp, _vp, _i, _vp]),
    'qgb_reset_time': (_i, [_vp]),
    'qgb_get': (_i, [_vp, _i, _vp, _i, _vp]),
    'qgb_get_f32': (_i, [_vp, _i, _vp, _i, _i, _vp]),
    'qgb_invert': (_i, [_vp, _vp]),
    'qgb_step': (_i, [_vp, _i, _vp]),
    'qgb_graph_replays': (ctypes.c_int64, [_vp]),
    'qgb_get_time': (_i, [_vp, ctypes.POINTER(_d), ctypes.POINTER(ctypes.c_int64)]),
    'qgb_cnn_load': (_i, [_vp, _i, _i, _i, ctypes.POINTER(QgbCnnLayer)]),
    'qgb_closure_config': (_i, [_vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), _d, _i]),
    'qgb_closure_precision': (_i, [_vp, ctypes.POINTER(_i), ctypes.POINTER(_d)]),
    'qgb_set_sampler': (_i, [_vp, _i, _i, _i]),
    'qgb_seed': (_i, [_vp, ctypes.c_uint64]),
    'qgb_set_latent': (_i, [_vp, _vp, _i, _i, _vp]),
    'qgb_closure_eval': (_i, [_vp, _vp]),
    'qgb_set_forcing': (_i, [_vp, _vp, _i, _vp]),
    'qgb_cnn_forward': (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    'qgb_step_host': (_i, [_vp, _vp, _vp, _i, _vp]),
    'qgb_step_host_async': (_i, [_vp, _vp, _vp, _i, _vp]),
    'qgb_diag': (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    'qgb_diag_spectra': (_i, [_vp, _vp, _vp, _i, _vp]),
    'qgb_diag_budget': (_i, [_vp, _vp, _i, _vp]),
    'qgb_diag_config': (_i, [_vp, _d, _d]),
    'qgb_diag_averages': (_i, [_vp, _vp, ctypes.POINTER(ctypes.c_int64), _i, _i, _vp]),
    'qgb_operator': (_i, [_i, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    'qgb_subgrid_forcing': (_i, [ctypes.POINTER(QgbConfig), _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    'qgb_fft_interpolate': (_i, [_i, _i, _i, _i, _vp, _vp, _i, _vp]),
    'qgb_profile_begin': (_i, [_vp, _i, _i]),
    'qgb_profile_end': (_i, [_vp, ctypes.POINTER(_d), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
    'qgb_profile_all_begin': (_i, [_vp]),
    'qgb_profile_all_end': (_i, [_vp, ctypes.POINTER(_d), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
    'qgb_launch_count': (ctypes.c_int64, []),
    'qgb_version': (ctypes.c_char_p, []),
    # training (tools/cnn_tools.py:645-700)
    'qgb_train_create': (_i, [_i, _i, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), _i, _i, _i, _i,
                              ctypes.POINTER(_vp)]),
    'qgb_train_destroy': (None, [_vp]),
    'qgb_train_last_error': (ctypes.c_char_p, [_vp]),
    'qgb_train_num_params': (ctypes.c_int64, [_vp]),
    'qgb_train_num_buffers': (ctypes.c_int64, [_vp]),
    'qgb_train_launch_count': (ctypes.c_int64, [_vp]),
    'qgb_train_set_params': (_i, [_vp, _vp, _vp, _i]),
    'qgb_train_get_params': (_i, [_vp, _vp, _vp]),
    'qgb_train_step': (_i, [_vp, _vp, _vp, _i, _i, _d, ctypes.POINTER(_d), _vp]),
    'qgb_train_grads': (_i, [_vp, _vp, _vp, _i, _i, _vp, ctypes.POINTER(_d), _i, _vp]),
    'qgb_train_eval_loss': (_i, [_vp, _vp, _vp, _i, _i, ctypes.POINTER(_d), _vp]),
    'qgb_train_get_grads': (_i, [_vp, _vp]),
    'qgb_train_set_adam': (_i, [_vp, _d, _d]),
    # CVAE / CGAN training steps (models/cvae_regression.py:250-300, models/cgan_regression.py:227-300)
    'qgb_train_cvae_step': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _d, _d, _i, ctypes.POINTER(_d), _vp]),
    'qgb_disc_create': (_i, [_i, _i, _i, _i, _i, ctypes.POINTER(_vp)]),
    'qgb_disc_destroy': (None, [_vp]),
    'qgb_disc_last_error': (ctypes.c_char_p, [_vp]),
    'qgb_disc_num_params': (ctypes.c_int64, [_vp]),
    'qgb_disc_launch_count': (ctypes.c_int64, [_vp]),
    'qgb_disc_set_params': (_i, [_vp, _vp, _i]),
    'qgb_disc_get_params': (_i, [_vp, _vp, _vp]),
    'qgb_disc_forward': (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    'qgb_train_cgan_step': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _d, _d, _i, _i, ctypes.POINTER(_d), _vp]),
}

_lib = None


def load():
    """Load libqgb200.so (once).  Raises ImportError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError('%s not found: build it with `python -m pyqg_generative_b200.build` '
                          '(nvcc, sm_100a).  There is no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class QgbError(RuntimeError):
    pass


def check(rc, handle=None):
    if rc == QGB_OK:
        return
    msg = load().qgb_last_error(handle)
    msg = msg.decode() if msg else 'error %d' % rc
    if rc == QGB_EINVAL:
        raise ValueError(msg)
    if rc == QGB_EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise QgbError(msg)


def check_train(rc, trainer=None):
    if rc == QGB_OK:
        return
    msg = load().qgb_train_last_error(trainer)
    msg = msg.decode() if msg else 'error %d' % rc
    if rc == QGB_EINVAL:
        raise ValueError(msg)
    if rc == QGB_EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise QgbError(msg)


def check_disc(rc, disc=None):
    if rc == QGB_OK:
        return
    msg = load().qgb_disc_last_error(disc)
    msg = msg.decode() if msg else 'error %d' % rc
    if rc == QGB_EINVAL:
        raise ValueError(msg)
    if rc == QGB_EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise QgbError(msg)


def launch_count():
    return int(load().qgb_launch_count())
