"""Batched two-layer QG model on the GPU + the noise samplers of ``pyqg_generative/tools/stochastic_pyqg.py``.

``EnsembleQGModel`` is a look-alike of ``pyqg.QGModel`` / ``stochastic_QGModel`` (reference :74-88) with a leading
member axis: B independent members are integrated by libqgb200's fused sm_100a kernels (one CTA per member and time
step), and a CNN closure loaded into the engine is evaluated on the device inside ``_step_forward`` -- no per-step
host round trip.  ``stochastic_QGModel(pyqg_params, sampling_type, nsteps)`` keeps the reference constructor.

Attribute names follow pyqg 0.7.2: q, qh, ph, u, v, p, dqhdt, ik, il, k, l, wv, wv2, filtr,
dx, L, nx, ny, t, tc, dt, Ubg, Hi, H, set_q1q2, _invert, _step_forward, run_with_snapshots, ...
"""
import ctypes
import logging

import numpy as np

from .. import _lib


# ---------------------------------------------------------------------------------------------------------
# noise samplers (host-side mirrors; the engine keeps the device-resident equivalent, csrc/api.cu closure_update)
# ---------------------------------------------------------------------------------------------------------
class noise_time_sampler(object):
    """Base class: ``update(generate_noise) -> bool`` says whether the SGS force must be recomputed."""

    def __init__(self, nsteps):
        self.nsteps = nsteps

    def update(self, generate_noise):
        raise NotImplementedError


class AR1_sampler(noise_time_sampler):
    """AR(1) in time latent noise, decorrelation ``nsteps`` steps (reference :30-54).
    nsteps=1 is white noise, nsteps<0 freezes the first draw."""

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


class constant_sampler(noise_time_sampler):
    """Piecewise-constant latent noise redrawn every ``nsteps`` steps; the force is reused in between (reference :56-72)."""

    def update(self, generate_noise):
        if not hasattr(self, 'noise'):
            self.noise = generate_noise()
            self.counter = 1
            return True
        if self.counter % self.nsteps == 0:
            self.noise = generate_noise()
            self.counter = 1
            return True
        self.counter += 1
        return False


_SAMPLERS = {'AR1': _lib.SAMPLER_AR1, 'constant': _lib.SAMPLER_CONSTANT, 'deterministic': _lib.SAMPLER_DETERMINISTIC}


class _DeviceSamplerView(object):
    """``m.noise_sampler`` of a device-coupled model: ``.noise`` reads the latent field back from the engine."""

    def __init__(self, model, nsteps):
        self._m = model
        self.nsteps = nsteps

    @property
    def noise(self):
        m = self._m
        gz = m._closure_kind == _lib.CLOSURE_GZ
        out = np.empty((m.members, 2, m.ny, m.nx), dtype='float64' if gz else 'float32')
        m._get_into(_lib.F_NOISE, out)
        return out


# ---------------------------------------------------------------------------------------------------------
class EnsembleQGModel(object):
    """``members`` independent two-layer QG models advanced together on one GPU."""

    def __init__(self, members=1, member_offset=0, device=None, nz=2, nx=64, ny=None, L=1e6, W=None, dt=7200.,
                 twrite=1000., tmax=1576800000., tavestart=315360000., taveint=86400., useAB2=False,
                 rek=5.787e-7, filterfac=23.6, f=None, g=9.81, q_parameterization=None, uv_parameterization=None,
                 parameterization=None, diagnostics_list='all', ntd=1, log_level=1, logfile=None,
                 beta=1.5e-11, rd=15000.0, delta=0.25, H1=500, U1=0.025, U2=0.0,
                 sampling_type='AR1', nsteps=1, precision=None, seed=None, squeeze=False, **kwargs):
        if kwargs:      # pyqg.Model.__init__ takes no **kwargs: unknown keywords are a TypeError there too
            raise TypeError("__init__() got an unexpected keyword argument '%s'" % sorted(kwargs)[0])
        if nz != 2:
            raise ValueError('QGModel is a two-layer model')
        if (ny is not None and ny != nx) or (W is not None and W != L):
            raise NotImplementedError('only square domains (ny == nx, W == L) are supported')
        if useAB2:
            raise NotImplementedError('useAB2 is not supported (pyqg default is AB3)')
        if uv_parameterization is not None:
            raise NotImplementedError('uv parameterizations are outside the hot path')
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('EnsembleQGModel needs a CUDA device: libqgb200 has no CPU fallback')
        self._torch = torch
        self._lib = _lib.load()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.members, self.member_offset = int(members), int(member_offset)
        self.squeeze = bool(squeeze) and self.members == 1
        self.nz, self.nx, self.ny = 2, int(nx), int(nx)
        self.L = self.W = float(L)
        self.dt, self.twrite, self.tmax = float(dt), twrite, float(tmax)
        self.tavestart, self.taveint = float(tavestart), float(taveint)
        self.rek, self.filterfac, self.beta, self.rd, self.delta = rek, filterfac, beta, rd, delta
        self.H1, self.U1, self.U2 = H1, U1, U2
        self.log_level = log_level
        self.logger = logging.getLogger('pyqg_generative_b200')
        self.log = []
        self._init_host_grid()

        cfg = _lib.QgbConfig()
        self._lib.qgb_default_config(ctypes.byref(cfg))
        cfg.nx, cfg.members, cfg.member_offset, cfg.device = self.nx, self.members, self.member_offset, self.device_index
        cfg.L, cfg.dt, cfg.rek, cfg.filterfac, cfg.beta = self.L, self.dt, float(rek), float(filterfac), float(beta)
        cfg.rd, cfg.delta, cfg.H1, cfg.U1, cfg.U2 = float(rd), float(delta), float(H1), float(U1), float(U2)
        self._cfg = cfg
        self._h = ctypes.c_void_p()
        _lib.check(self._lib.qgb_create(ctypes.byref(cfg), ctypes.byref(self._h)))
        _lib.check(self._lib.qgb_diag_config(self._h, self.tavestart, self.taveint), self._h)

        self.t, self.tc = 0.0, 0
        self.sampling_type = sampling_type
        self._precision = precision
        self._closure = None
        self._closure_kind = _lib.CLOSURE_NONE
        self._host_param = None
        self._pv_host = None
        self.q_parameterization = None
        self.diag_count = 0
        if sampling_type not in _SAMPLERS:
            raise ValueError('Unknown sampling type')
        self._nsteps = nsteps
        if seed is not None:
            _lib.check(self._lib.qgb_seed(self._h, int(seed)), self._h)
        if parameterization is not None and q_parameterization is None:
            if getattr(parameterization, 'parameterization_type', None) != 'q_parameterization':
                raise ValueError('only q parameterizations are supported')
            q_parameterization = parameterization
        if q_parameterization is not None:
            self.set_parameterization(q_parameterization, sampling_type, nsteps)

    # ---- grid (pyqg Model._initialize_grid / _initialize_filter / QGModel._initialize_background) ----------
    def _init_host_grid(self):
        N, L = self.nx, self.L
        self.x, self.y = np.meshgrid(np.arange(0.5, N, 1.) / N * L, np.arange(0.5, N, 1.) / N * L)
        self.nl, self.nk = N, N // 2 + 1
        self.dk = self.dl = 2. * np.pi / L
        self.ll = self.dl * np.append(np.arange(0., N / 2), np.arange(-N / 2, 0.))
        self.kk = self.dk * np.arange(0., self.nk)
        self.k, self.l = np.meshgrid(self.kk, self.ll)
        self.ik, self.il = 1j * self.k, 1j * self.l
        self.dx = self.dy = L / N
        self.M = N * N
        self.wv2 = self.k ** 2 + self.l ** 2
        self.wv = np.sqrt(self.wv2)
        cphi = 0.65 * np.pi
        wvx = np.sqrt((self.k * self.dx) ** 2. + (self.l * self.dy) ** 2.)
        with np.errstate(over='ignore', under='ignore'):
            filtr = np.exp(-self.filterfac * (wvx - cphi) ** 4.)
        filtr[wvx <= cphi] = 1.
        self.filtr = filtr
        self.Hi = np.array([self.H1, self.H1 / self.delta])
        self.H = self.Hi.sum()
        self.Ubg = np.array([self.U1, self.U2])
        self.F1 = self.rd ** -2 / (1. + self.delta)
        self.F2 = self.delta * self.F1
        self.Qy = np.array([self.beta + self.F1 * (self.U1 - self.U2), self.beta - self.F2 * (self.U1 - self.U2)])

    # ---- low-level helpers -------------------------------------------------------------------------------------
    def _stream(self):
        return self._torch.cuda.current_stream(self.device_index).cuda_stream

    def _get_into(self, field, out):
        _lib.check(self._lib.qgb_get(self._h, field, out.ctypes.data, 0, self._stream()), self._h)
        return out

    def _real(self, field):
        out = np.empty((self.members, 2, self.ny, self.nx))
        self._get_into(field, out)
        return out[0] if self.squeeze else out

    def real32(self, name, out=None, stream=None, wait=True):
        """float32 copy of ``q`` / ``u`` / ``v`` / ``p`` converted on the device (what ``drop_vars`` stores, tools/simulate.py:
        16-36): half the PCIe bytes of ``np.asarray(m.q, 'float32')``.  ``out``: optional (pinned) float32 buffer."""
        field = {'q': _lib.F_Q, 'u': _lib.F_U, 'v': _lib.F_V, 'p': _lib.F_P, 'psi': _lib.F_P}[name]
        if out is None:
            out = np.empty((self.members, 2, self.ny, self.nx), dtype=np.float32)
        ptr = out.data_ptr() if hasattr(out, 'data_ptr') else out.ctypes.data
        s = stream.cuda_stream if stream is not None else self._stream()
        _lib.check(self._lib.qgb_get_f32(self._h, field, ptr, 0, 0 if wait else 1, s), self._h)
        return out

    def _cplx(self, field):
        out = np.empty((self.members, 2, self.nl, self.nk), dtype=complex)
        self._get_into(field, out)
        return out[0] if self.squeeze else out

    def _sync_time(self):
        t, tc = ctypes.c_double(), ctypes.c_int64()
        self._lib.qgb_get_time(self._h, ctypes.byref(t), ctypes.byref(tc))
        self.t, self.tc = t.value, int(tc.value)

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                self._lib.qgb_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- state (pyqg kernel properties) ------------------------------------------------------------------------
    @property
    def q(self):
        return self._real(_lib.F_Q)

    @q.setter
    def q(self, value):
        self.set_q(value)

    def set_q(self, value):
        """pyqg ``q`` setter: stores q and refreshes qh.  Accepts (2,ny,nx) (broadcast to all members) or
        (B,2,ny,nx); numpy, or a CUDA torch tensor (used in place, no host copy)."""
        torch = self._torch
        shape = (self.members, 2, self.ny, self.nx)
        if isinstance(value, torch.Tensor) and value.is_cuda:
            v = value.to(dtype=torch.float64)
            if v.dim() == 3:
                v = v.unsqueeze(0).expand(shape)
            v = v.contiguous()
            if tuple(v.shape) != shape:
                raise ValueError('q must have shape %s' % (shape,))
            _lib.check(self._lib.qgb_set_q(self._h, v.data_ptr(), 1, self._stream()), self._h)
            torch.cuda.current_stream(self.device_index).synchronize()
            return
        v = np.asarray(value, dtype=np.float64)
        if v.ndim == 3:
            v = np.broadcast_to(v[None], shape)
        if v.shape != shape:
            raise ValueError('q must have shape %s' % (shape,))
        v = np.ascontiguousarray(v)
        _lib.check(self._lib.qgb_set_q(self._h, v.ctypes.data, 0, self._stream()), self._h)

    def set_q1q2(self, q1, q2, check=False):
        q1, q2 = np.asarray(q1, dtype=np.float64), np.asarray(q2, dtype=np.float64)
        self.set_q(np.stack([q1, q2], axis=-3))

    def device_q(self):
        """Copy of q as a CUDA float64 torch tensor (B,2,ny,nx) (device-to-device, no host traffic)."""
        out = self._torch.empty((self.members, 2, self.ny, self.nx), dtype=self._torch.float64,
                                device='cuda:%d' % self.device_index)
        _lib.check(self._lib.qgb_get(self._h, _lib.F_Q, out.data_ptr(), 1, self._stream()), self._h)
        return out

    @property
    def qh(self):
        return self._cplx(_lib.F_QH)

    @property
    def ph(self):
        return self._cplx(_lib.F_PH)

    @property
    def u(self):
        return self._real(_lib.F_U)

    @property
    def v(self):
        return self._real(_lib.F_V)

    @property
    def p(self):
        return self._real(_lib.F_P)

    @property
    def dqhdt(self):
        return self._cplx(_lib.F_DQHDT)

    @property
    def PV_forcing(self):
        if self._closure is None:
            return self._pv_host      # set by a host-side Parameterization.__call__
        return self._real(_lib.F_FORCING)

    @PV_forcing.setter
    def PV_forcing(self, value):
        self._pv_host = value

    @property
    def ufull(self):
        return self.u + self.Ubg[:, np.newaxis, np.newaxis]

    @property
    def vfull(self):
        return self.v

    # ---- closure coupling ------------------------------------------------------------------------------------
    def set_parameterization(self, param, sampling_type='AR1', nsteps=1):
        """Attach a q-parameterization.  Device closures (pyqg_generative_b200.models.*, optionally wrapped as
        ``weight * model``) are loaded into the engine; any other callable ``dq = param(m)`` is honoured through a
        per-step host callback (generic pyqg.QParameterization behaviour)."""
        from ..models.parameterization import DeviceClosure, WeightedParameterization
        if sampling_type not in _SAMPLERS:
            raise ValueError('Unknown sampling type')
        self.sampling_type = sampling_type
        self._nsteps = nsteps
        self.q_parameterization = param
        weight, inner = 1.0, param
        while isinstance(inner, WeightedParameterization):
            weight *= inner.weight
            inner = inner.param
        if isinstance(inner, DeviceClosure):
            # CNN precision of the coupled run: the model's ``precision`` keyword when given, else the one the closure's networks were
            # built with (CGANRegression(..., precision='tc') alone must not fall back to the fp32 path), else fp32
            prec = self._precision
            if prec is None:
                nets = inner._nets()
                prec = getattr(nets[0], 'precision', None) if nets else None
            inner._attach(self, weight, prec or 'fp32')
            self._closure = inner
            self._closure_kind = inner.closure_kind
            self._host_param = None
            _lib.check(self._lib.qgb_set_sampler(self._h, _SAMPLERS[sampling_type], int(nsteps),
                                                 getattr(inner, 'n_mean', 100)), self._h)
            if sampling_type != 'deterministic':
                self.noise_sampler = _DeviceSamplerView(self, nsteps)
        else:
            self._closure = None
            self._host_param = param
            if sampling_type == 'AR1':
                self.noise_sampler = AR1_sampler(nsteps)
            elif sampling_type == 'constant':
                self.noise_sampler = constant_sampler(nsteps)

    def set_latent(self, xi):
        """Inject the white noise used by the next sampler update (parity tests): float32 (B,2,ny,nx) for
        GAN/VAE, float64 for GZ."""
        xi = np.ascontiguousarray(xi)
        dtype = 1 if xi.dtype == np.float64 else 0
        if xi.dtype not in (np.float32, np.float64):
            raise ValueError('latent noise must be float32 or float64')
        xi = xi.reshape(self.members, 2, self.ny, self.nx)
        _lib.check(self._lib.qgb_set_latent(self._h, xi.ctypes.data, dtype, 0, self._stream()), self._h)

    def closure_precision(self):
        """(name, errors): the CNN precision in effect ('auto' until the first closure evaluation has calibrated it) and, for
        precision='auto', the measured errors against the fp32 path: dict(tc_l2, tc_max, tc_fast_l2, tc_fast_max)."""
        pr = ctypes.c_int(0)
        err = (ctypes.c_double * 4)()
        _lib.check(self._lib.qgb_closure_precision(self._h, ctypes.byref(pr), err), self._h)
        names = ('tc_l2', 'tc_max', 'tc_fast_l2', 'tc_fast_max')
        return _lib.PRECISION_NAMES[pr.value], {k: (v if v >= 0 else None) for k, v in zip(names, err)}

    def seed(self, seed):
        _lib.check(self._lib.qgb_seed(self._h, int(seed)), self._h)

    def closure_eval(self):
        """``Parameterization.__call__(m)`` on the device; returns the demeaned forcing."""
        _lib.check(self._lib.qgb_closure_eval(self._h, self._stream()), self._h)
        return self.PV_forcing

    # ---- dynamics (pyqg kernel / Model methods) -----------------------------------------------------------------
    def _invert(self):
        _lib.check(self._lib.qgb_invert(self._h, self._stream()), self._h)

    def _calc_derived_fields(self):
        self._invert()

    def _step_forward(self, nsteps=1):
        if self._host_param is not None:
            for _ in range(nsteps):
                dq = np.asarray(self._host_param(self), dtype=np.float64)
                dq = np.ascontiguousarray(np.broadcast_to(dq.reshape((-1, 2, self.ny, self.nx)),
                                                          (self.members, 2, self.ny, self.nx)))
                _lib.check(self._lib.qgb_set_forcing(self._h, dq.ctypes.data, 0, self._stream()), self._h)
                _lib.check(self._lib.qgb_step(self._h, 1, self._stream()), self._h)
                self._sync_time()
                self._after_step()
            return
        done = 0
        while done < nsteps:
            n = min(nsteps - done, self._steps_to_next_event())
            _lib.check(self._lib.qgb_step(self._h, int(n), self._stream()), self._h)
            done += n
            self._sync_time()
            self._after_step()

    def _steps_to_next_event(self):
        """Steps that can be fused into one engine call before the host has to look (log line / diagnostics)."""
        nxt = self.tc + (1 << 30)
        if self.log_level and self.twrite:
            tw = int(self.twrite)
            nxt = min(nxt, (self.tc // tw + 1) * tw)
        return max(1, nxt - self.tc)      # (time-averaged diagnostics are sampled on the device inside qgb_step)

    def _after_step(self):
        if self.log_level and self.twrite and self.tc % int(self.twrite) == 0:
            self._print_status()

    def diagnostics(self):
        """Per-member (KE, CFL, flags): pyqg _calc_ke / _calc_cfl; flags bit0 non-finite, bit1 CFL>=1."""
        ke = np.empty(self.members)
        cfl = np.empty(self.members)
        flags = np.empty(self.members, dtype=np.int32)
        _lib.check(self._lib.qgb_diag(self._h, ke.ctypes.data, cfl.ctypes.data, flags.ctypes.data, 0, self._stream()),
                   self._h)
        return ke, cfl, flags

    def _calc_ke(self):
        ke = self.diagnostics()[0]
        return ke[0] if self.squeeze else ke

    def _calc_cfl(self):
        cfl = self.diagnostics()[1]
        return cfl[0] if self.squeeze else cfl

    def _print_status(self):
        ke, cfl, flags = self.diagnostics()
        self.ke, self.cfl, self.flags = ke, cfl, flags
        self.log.append((self.tc, self.t, float(np.nanmean(ke)), float(np.nanmax(cfl))))
        self.logger.info('Step: %i, Time: %3.2e, KE: %3.2e, CFL: %4.3f', self.tc, self.t, np.nanmean(ke), np.nanmax(cfl))
        if self.squeeze:
            assert cfl[0] < 1., 'CFL condition violated'

    # ---- pyqg time-averaged diagnostics (Model tavestart / taveint; sampled on the device before each eligible step) ----
    DIAG_LAYERED = ('KEspec', 'Ensspec')
    DIAG_BUDGET = ('KEflux', 'APEflux', 'APEgenspec', 'KEfrictionspec', 'entspec', 'paramspec_KEflux', 'paramspec_APEflux',
                   'ENSflux', 'ENSgenspec', 'ENSfrictionspec', 'Dissspec', 'ENSDissspec', 'ENSparamspec')

    def _split_terms(self, a, first):
        a = a.reshape((-1, self.nl, self.nk))
        return {name: a[first + i] for i, name in enumerate(self.DIAG_BUDGET)}

    def budget_sums(self):
        """Spectral energy budget of the CURRENT state summed over the local members: dict name -> (nl, nk)
        (pyqg diagnostics KEflux, APEflux, APEgenspec, KEfrictionspec, entspec, paramspec_KEflux, paramspec_APEflux, ENSflux,
        ENSgenspec, ENSfrictionspec, Dissspec, ENSDissspec, ENSparamspec; consumers: tools/comparison_tools.py:91,164-189,
        222-263)."""
        out = np.empty(len(self.DIAG_BUDGET) * self.nl * self.nk)
        _lib.check(self._lib.qgb_diag_budget(self._h, out.ctypes.data, 0, self._stream()), self._h)
        return self._split_terms(out, 0)

    def diagnostic_sums(self):
        """(dict name -> sum over local members and averaging times, count = members * samples): the accumulators that
        ``parallel.ensemble_diagnostics`` all-reduces.  KEspec / Ensspec are (2, nl, nk), the budget terms (nl, nk)."""
        n = (4 + len(self.DIAG_BUDGET)) * self.nl * self.nk
        out = np.empty(n)
        ns = ctypes.c_int64(0)
        _lib.check(self._lib.qgb_diag_averages(self._h, out.ctypes.data, ctypes.byref(ns), 0, 0, self._stream()), self._h)
        a = out.reshape((-1, self.nl, self.nk))
        d = {'KEspec': a[0:2].copy(), 'Ensspec': a[2:4].copy()}
        d.update({k: v.copy() for k, v in self._split_terms(out, 4).items()})
        self.diag_count = int(ns.value)
        return d, int(ns.value) * self.members

    def averaged_diagnostics(self):
        """Ensemble- (local members) and time-mean diagnostics like pyqg's ``m.get_diagnostic``; adds ``paramspec``."""
        d, count = self.diagnostic_sums()
        if not count:
            return {}
        out = {k: v / count for k, v in d.items()}
        out['paramspec'] = out['paramspec_KEflux'] + out['paramspec_APEflux']
        out.update(self.derived_scalars(out['KEspec']))
        return out

    def derived_scalars(self, kespec):
        """pyqg's scalar diagnostics that are linear in KEspec: ``EKE`` = 0.5 (u^2 + v^2).mean() per layer (Parseval: the
        half-plane sum of KEspec with the k = 0 and Nyquist columns counted once, ``Model.spec_var``) and ``EKEdiss`` =
        Hi[-1]/H rek (u_2^2 + v_2^2).mean().  Linear, so the averaged KEspec gives the averaged scalars."""
        w = np.full(self.nk, 2.0)
        w[0] = w[-1] = 1.0
        eke = 0.5 * (np.asarray(kespec) * w).sum(axis=(-1, -2))
        return {'EKE': eke, 'EKEdiss': self.Hi[-1] / self.H * self.rek * 2.0 * eke[-1]}

    def spectra_sums(self):
        """(KEspec_sum, Ensspec_sum, count): sums over local members and averaging times, shape (2,nl,nk);
        ``count`` = members * samples.  These are the accumulators all-reduced over NCCL (parallel.py)."""
        d, count = self.diagnostic_sums()
        return d['KEspec'], d['Ensspec'], count

    def run_with_snapshots(self, tsnapstart=0., tsnapint=432000.):
        tsnapints = int(np.ceil(tsnapint / self.dt))
        while self.t < self.tmax:
            if self._host_param is not None:
                n = 1
            else:
                n = tsnapints - self.tc % tsnapints
                n = min(n, int(np.ceil((self.tmax - self.t) / self.dt)))
            self._step_forward(max(1, n))
            if self.t >= tsnapstart and (self.tc % tsnapints) == 0:
                yield self.t
        return

    def run(self):
        for _ in self.run_with_snapshots(tsnapint=1e30):
            pass

    def step_host(self, q_in, q_out, nsteps=1, stream=None, wait=True):
        """Host-buffer step: upload ``q_in`` (pinned CPU torch tensor or numpy array, (B,2,ny,nx) float64), advance
        ``nsteps``, download q into ``q_out``.  With ``wait=False`` the call only enqueues work on ``stream`` (a
        torch.cuda.Stream); drive several member groups on different streams to overlap PCIe transfers and kernels."""
        def ptr(a):
            if a is None:
                return None
            return a.data_ptr() if hasattr(a, 'data_ptr') else a.ctypes.data
        s = (stream.cuda_stream if stream is not None else self._stream())
        fn = self._lib.qgb_step_host if wait else self._lib.qgb_step_host_async
        _lib.check(fn(self._h, ptr(q_in), ptr(q_out), int(nsteps), s), self._h)
        self.tc += int(nsteps)
        self.t += int(nsteps) * self.dt

    def reset_time(self):
        _lib.check(self._lib.qgb_reset_time(self._h), self._h)
        self._sync_time()

    # spectral helpers kept for API compatibility (host numpy: not on the step path)
    def fft(self, x):
        return np.fft.rfftn(np.asarray(x, dtype=np.float64), axes=(-2, -1))

    def ifft(self, xh):
        return np.fft.irfftn(np.asarray(xh), s=(self.ny, self.nx), axes=(-2, -1))

    def spec_var(self, ph):
        var_dens = 2. * np.abs(ph) ** 2 / self.M ** 2
        var_dens[..., 0] /= 2
        var_dens[..., -1] /= 2
        return var_dens.sum(axis=(-1, -2))


class stochastic_QGModel(EnsembleQGModel):
    """Reference constructor ``stochastic_QGModel(pyqg_params, sampling_type='AR1', nsteps=1)`` (:74-88).
    ``pyqg_params`` may carry the extra keys ``members``, ``member_offset``, ``device``, ``precision``, ``seed``."""

    def __init__(self, pyqg_params, sampling_type='AR1', nsteps=1):
        if sampling_type not in ('AR1', 'constant', 'deterministic'):
            raise ValueError('Unknown sampling type')
        params = dict(pyqg_params)
        params.setdefault('squeeze', params.get('members', 1) == 1)
        super().__init__(sampling_type=sampling_type, nsteps=nsteps, **params)
