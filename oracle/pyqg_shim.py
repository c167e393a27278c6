"""CPU ORACLE (test infrastructure, NOT product code) -- numpy restatement of pyqg 0.7.2 ``QGModel``.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module.  The product (``pyqg_generative_b200``) never does.

PARITY PINNING STATUS
---------------------
The arithmetic of the dynamical core lives in the third-party dependency ``pyqg`` (resolved
version 0.7.2, FFT backend pyFFTW 0.13.1 -- /root/reference/Google-Colab/online-simulations.ipynb:35-40,92,99),
which is NOT vendored under /root/reference and is NOT installable here (no network).  This file
restates the published algorithm of upstream ``pyqg/model.py``, ``pyqg/qg_model.py``,
``pyqg/kernel.pyx`` and ``pyqg/parameterizations.py`` (function names are cited next to each block).
The reference ships no test-suite, so for the *stepping arithmetic* this oracle is
**parity unpinned** by any reference golden vector; it is anchored on

* the reference's own call sites (tools/simulate.py:83,121,125,132,137,168; tools/operators.py:229-234,
  241-257,292-326; tools/stochastic_pyqg.py:74-88) which exercise every attribute defined here,
* the recorded notebook identities / logs replayed in tests/test_oracle_pins.py
  (q-setter round trip = 0.0, notebooks/3-2-dealiasing.ipynb:88; initial CFL 0.023 / 0.009,
  notebooks/3-2-dealiasing.ipynb:1431 and Google-Colab/online-simulations.ipynb:318; KE growth curve).

Everything *in-tree* in the reference (CNN, samplers, operators, closures) is pinned properly: those
modules are imported unmodified on top of this shim by tests/golden/make_golden.py and their outputs
are committed as fixtures.
"""
import numpy as np

try:  # scipy's pocketfft is ~2x faster than numpy's for these sizes; same convention
    from scipy import fft as _fft
    _rfft2 = lambda x: _fft.rfft2(x, axes=(-2, -1))
    _irfft2 = lambda x, s: _fft.irfft2(x, s=s, axes=(-2, -1))
except Exception:  # pragma: no cover
    _rfft2 = lambda x: np.fft.rfftn(x, axes=(-2, -1))
    _irfft2 = lambda x, s: np.fft.irfftn(x, s=s, axes=(-2, -1))


# --------------------------------------------------------------------------------------
# pyqg/parameterizations.py : Parameterization / QParameterization / Weighted / Composite
# --------------------------------------------------------------------------------------
class Parameterization(object):
    """pyqg.parameterizations.Parameterization: callable closure + algebra (``__add__``, ``__mul__``)."""

    @property
    def parameterization_type(self):
        raise NotImplementedError

    def __call__(self, m):
        raise NotImplementedError

    def __add__(self, other):
        return CompositeParameterization(self, other)

    def __mul__(self, constant):
        return WeightedParameterization(self, constant)

    __rmul__ = __mul__


class QParameterization(Parameterization):
    @property
    def parameterization_type(self):
        return 'q_parameterization'


class UVParameterization(Parameterization):
    @property
    def parameterization_type(self):
        return 'uv_parameterization'


class CompositeParameterization(Parameterization):
    def __init__(self, *params):
        assert len(set(p.parameterization_type for p in params)) == 1
        self.params = params

    @property
    def parameterization_type(self):
        return self.params[0].parameterization_type

    def __call__(self, m):
        return np.sum([np.array(p(m)) for p in self.params], axis=0)


class WeightedParameterization(Parameterization):
    def __init__(self, param, weight):
        self.param = param
        self.weight = weight

    @property
    def parameterization_type(self):
        return self.param.parameterization_type

    def __call__(self, m):
        return np.array(self.param(m)) * self.weight


# --------------------------------------------------------------------------------------
# pyqg/model.py : Model  +  pyqg/qg_model.py : QGModel  +  pyqg/kernel.pyx
# --------------------------------------------------------------------------------------
class QGModel(object):
    """Two-layer quasi-geostrophic pseudo-spectral model, pyqg 0.7.2 semantics."""

    def __init__(self, nz=2, nx=64, ny=None, L=1e6, W=None, dt=7200., twrite=1000., tmax=1576800000.,
                 tavestart=315360000., taveint=86400., useAB2=False, rek=5.787e-7, filterfac=23.6,
                 f=None, g=9.81, q_parameterization=None, uv_parameterization=None, parameterization=None,
                 diagnostics_list='all', ntd=1, log_level=1, logfile=None,
                 beta=1.5e-11, rd=15000.0, delta=0.25, H1=500, U1=0.025, U2=0.0, **kwargs):
        # Model.__init__
        if ny is None:
            ny = nx
        if W is None:
            W = L
        self.nz, self.nx, self.ny = 2, int(nx), int(ny)
        self.L, self.W = float(L), float(W)
        self.dt, self.twrite, self.tmax = float(dt), twrite, float(tmax)
        self.tavestart, self.taveint = float(tavestart), float(taveint)
        self.useAB2 = useAB2
        self.rek, self.filterfac = rek, filterfac
        self.log_level = log_level
        self.beta, self.rd, self.delta = beta, rd, delta
        self.H1, self.U1, self.U2 = H1, U1, U2
        self.f, self.g = f, g
        self.log = []           # (step, t, ke, cfl) tuples written by _print_status

        self.q_parameterization = None
        self.uv_parameterization = None
        if parameterization is not None:
            ptype = getattr(parameterization, 'parameterization_type', None)
            if ptype == 'q_parameterization':
                q_parameterization = parameterization
            elif ptype == 'uv_parameterization':
                uv_parameterization = parameterization
            else:
                raise ValueError('unknown parameterization type')
        self.q_parameterization = q_parameterization
        self.uv_parameterization = uv_parameterization
        if uv_parameterization is not None:
            raise NotImplementedError('uv parameterizations are outside the hot path')

        self._initialize_grid()
        self._initialize_background()
        self._initialize_filter()
        self._initialize_time()
        self._initialize_inversion_matrix()
        self._initialize_diagnostics()

        # kernel state
        shp_r = (self.nz, self.ny, self.nx)
        shp_c = (self.nz, self.nl, self.nk)
        self._q = np.zeros(shp_r)
        self.qh = np.zeros(shp_c, complex)
        self.ph = np.zeros(shp_c, complex)
        self.u = np.zeros(shp_r)
        self.v = np.zeros(shp_r)
        self.dqhdt = np.zeros(shp_c, complex)
        self.dqhdt_p = np.zeros(shp_c, complex)
        self.dqhdt_pp = np.zeros(shp_c, complex)

        # QGModel.__init__ : default initial condition (overridden by tools/simulate.py:147-168)
        self.set_q1q2(1e-7 * np.random.rand(self.ny, self.nx)
                      + 1e-6 * (np.ones((self.ny, 1)) * np.random.rand(1, self.nx)),
                      np.zeros((self.ny, self.nx)))

    # ---- Model._initialize_grid -------------------------------------------------------
    def _initialize_grid(self):
        self.x, self.y = np.meshgrid(np.arange(0.5, self.nx, 1.) / self.nx * self.L,
                                     np.arange(0.5, self.ny, 1.) / self.ny * self.W)
        self.nl = self.ny
        self.nk = self.nx // 2 + 1
        self.dk = 2. * np.pi / self.L
        self.dl = 2. * np.pi / self.W
        self.ll = self.dl * np.append(np.arange(0., self.nx / 2), np.arange(-self.nx / 2, 0.))
        self.kk = self.dk * np.arange(0., self.nk)
        self.k, self.l = np.meshgrid(self.kk, self.ll)
        self.ik = 1j * self.k
        self.il = 1j * self.l
        self.dx = self.L / self.nx
        self.dy = self.W / self.ny
        self.M = self.nx * self.ny
        self.wv2 = self.k ** 2 + self.l ** 2
        self.wv = np.sqrt(self.wv2)
        iwv2 = self.wv2 != 0.
        self.wv2i = np.zeros_like(self.wv2)
        self.wv2i[iwv2] = self.wv2[iwv2] ** -1

    # ---- QGModel._initialize_background -----------------------------------------------
    def _initialize_background(self):
        self.Hi = np.array([self.H1, self.H1 / self.delta])
        self.H = self.Hi.sum()
        self.Ubg = np.array([self.U1, self.U2])
        self.U = self.U1 - self.U2
        self.F1 = self.rd ** -2 / (1. + self.delta)
        self.F2 = self.delta * self.F1
        self.Qy1 = self.beta + self.F1 * (self.U1 - self.U2)
        self.Qy2 = self.beta - self.F2 * (self.U1 - self.U2)
        self.Qy = np.array([self.Qy1, self.Qy2])
        self.ikQy = self.Qy[:, np.newaxis, np.newaxis] * 1j * self.k
        self.ilQx = 0.
        self.del1 = self.delta / (self.delta + 1.)
        self.del2 = (self.delta + 1.) ** -1

    # ---- Model._initialize_filter ------------------------------------------------------
    def _initialize_filter(self):
        cphi = 0.65 * np.pi
        wvx = np.sqrt((self.k * self.dx) ** 2. + (self.l * self.dy) ** 2.)
        with np.errstate(over='ignore', under='ignore'):
            filtr = np.exp(-self.filterfac * (wvx - cphi) ** 4.)
        filtr[wvx <= cphi] = 1.
        self.filtr = filtr

    def _initialize_time(self):
        self.t = 0.
        self.tc = 0
        self.ablevel = 0

    # ---- QGModel._initialize_inversion_matrix -----------------------------------------
    def _initialize_inversion_matrix(self):
        a = np.ma.zeros((self.nz, self.nz, self.nl, self.nk), np.dtype('float64'))
        det_inv = np.ma.masked_equal(self.wv2 * (self.wv2 + self.F1 + self.F2), 0.) ** -1
        a[0, 0] = -(self.wv2 + self.F2) * det_inv
        a[0, 1] = -self.F1 * det_inv
        a[1, 0] = -self.F2 * det_inv
        a[1, 1] = -(self.wv2 + self.F1) * det_inv
        self.a = np.ma.masked_invalid(a).filled(0.)

    # ---- FFT convention: rfft2 unnormalised, irfft2 carries 1/(nx*ny) ------------------
    def fft(self, x):
        return _rfft2(np.asarray(x, dtype=np.float64))

    def ifft(self, xh):
        return _irfft2(np.asarray(xh, dtype=complex), (self.ny, self.nx))

    # ---- kernel ``q`` property: the setter refreshes qh (relied on by operators.py:232-233) --
    @property
    def q(self):
        return self._q

    @q.setter
    def q(self, value):
        self._q = np.array(value, dtype=np.float64).reshape(self.nz, self.ny, self.nx)
        self.qh = self.fft(self._q)

    def set_q(self, q):
        self.q = q

    def set_q1q2(self, q1, q2, check=False):
        self.set_q(np.vstack([np.asarray(q1)[np.newaxis, :, :], np.asarray(q2)[np.newaxis, :, :]]))

    def set_qh(self, qh):
        self.qh = np.array(qh, dtype=complex)
        self._q = self.ifft(self.qh)

    # ---- kernel.pyx : _invert ------------------------------------------------------------
    def _invert(self):
        self.ph = np.einsum('ijlk,jlk->ilk', self.a, self.qh)
        self.uh = -self.il * self.ph
        self.vh = self.ik * self.ph
        self.u = self.ifft(self.uh)
        self.v = self.ifft(self.vh)

    # ---- kernel.pyx : _do_advection ------------------------------------------------------
    def _do_advection(self):
        uq = (self.u + self.Ubg[:, np.newaxis, np.newaxis]) * self._q
        vq = self.v * self._q
        self.uqh = self.fft(uq)
        self.vqh = self.fft(vq)
        self.dqhdt = -(self.ik * self.uqh + self.il * self.vqh + self.ikQy * self.ph)

    # ---- kernel.pyx : _do_friction -------------------------------------------------------
    def _do_friction(self):
        if self.rek:
            self.dqhdt[-1] = self.dqhdt[-1] + self.rek * self.wv2 * self.ph[-1]

    # ---- model.py : _do_q_subgrid_parameterization ---------------------------------------
    def _do_q_subgrid_parameterization(self):
        if self.q_parameterization is not None:
            self.dq = np.array(self.q_parameterization(self), dtype=np.float64)
            self.dqh = self.fft(self.dq)
            self.dqhdt = self.dqhdt + self.dqh

    # ---- kernel.pyx : _forward_timestep --------------------------------------------------
    def _forward_timestep(self):
        if self.ablevel == 0:
            dt1, dt2, dt3 = self.dt, 0.0, 0.0
            self.ablevel = 1
        elif self.ablevel == 1 and not self.useAB2:
            dt1, dt2, dt3 = 1.5 * self.dt, -0.5 * self.dt, 0.0
            self.ablevel = 2
        elif self.useAB2:
            dt1, dt2, dt3 = 1.5 * self.dt, -0.5 * self.dt, 0.0
        else:
            dt1, dt2, dt3 = 23. / 12. * self.dt, -16. / 12. * self.dt, 5. / 12. * self.dt
        self.qh = self.filtr * (self.qh + dt1 * self.dqhdt + dt2 * self.dqhdt_p + dt3 * self.dqhdt_pp)
        self.dqhdt_pp = self.dqhdt_p
        self.dqhdt_p = self.dqhdt
        self._q = self.ifft(self.qh)
        self.tc += 1
        self.t += self.dt

    # ---- model.py : _step_forward ----------------------------------------------------------
    def _step_forward(self):
        self._invert()
        self._do_advection()
        self._do_friction()
        self._do_q_subgrid_parameterization()
        self._calc_diagnostics()
        self._forward_timestep()
        self._print_status()

    def run_with_snapshots(self, tsnapstart=0., tsnapint=432000.):
        tsnapints = np.ceil(tsnapint / self.dt)
        while self.t < self.tmax:
            self._step_forward()
            if self.t >= tsnapstart and (self.tc % tsnapints) == 0:
                yield self.t
        return

    def run(self):
        while self.t < self.tmax:
            self._step_forward()

    # ---- qg_model.py : _calc_cfl / _calc_ke ; model.py : spec_var / _print_status -----------
    def _calc_cfl(self):
        return np.abs(np.hstack([self.u + self.Ubg[:, np.newaxis, np.newaxis], self.v])).max() * self.dt / self.dx

    def spec_var(self, ph):
        var_dens = 2. * np.abs(ph) ** 2 / self.M ** 2
        var_dens[..., 0] /= 2
        var_dens[..., -1] /= 2
        return var_dens.sum(axis=(-1, -2))

    def _calc_ke(self):
        ke1 = .5 * self.Hi[0] * self.spec_var(self.wv * self.ph[0])
        ke2 = .5 * self.Hi[1] * self.spec_var(self.wv * self.ph[1])
        return (ke1 + ke2) / self.H

    def _print_status(self):
        if self.log_level and (self.tc % self.twrite) == 0:
            self.ke = self._calc_ke()
            self.cfl = self._calc_cfl()
            self.log.append((self.tc, self.t, self.ke, self.cfl))
            assert self.cfl < 1., "CFL condition violated"

    # ---- model.py / qg_model.py : diagnostics -------------------------------------------------
    # Restated from pyqg 0.7.2 (``Model._initialize_diagnostics`` and ``QGModel._initialize_model_diagnostics``: the
    # ``add_diagnostic`` lambdas; ``Model._increment_diagnostics`` running mean).  pyqg is not installable here, so these
    # formulas are PARITY UNPINNED against pyqg itself; tests/test_oracle_pins.py pins them through identities that must
    # hold for the true definitions (energy conservation of the Jacobian terms, paramspec = KE part + APE part, and the
    # spectral energy budget d/dt E(k) = sum of the terms measured by stepping the model).
    # Consumers in the reference: tools/comparison_tools.py:91,164-189,222-263, tools/spectral_tools.py:103-180.
    def _initialize_diagnostics(self):
        self.diag_count = 0
        self.diag = {}

    def diagnostic_fields(self):
        """Instantaneous values of the time-averaged diagnostics for the current (inverted) state."""
        M2 = self.M ** 2
        self._calc_derived_fields()
        ph, qh = self.ph, self.qh
        hr = (self.Hi / self.H)[:, np.newaxis, np.newaxis]
        d = {
            'KEspec': self.wv2 * np.abs(ph) ** 2 / M2,
            'Ensspec': np.abs(qh) ** 2 / M2,
            'entspec': np.abs(self.del1 * qh[0] + self.del2 * qh[1]) ** 2 / M2,
            'KEflux': (np.real(self.del1 * ph[0] * np.conj(self.Jpxi[0])) + np.real(self.del2 * ph[1] * np.conj(self.Jpxi[1]))) / M2,
            'APEflux': self.rd ** -2 * self.del1 * self.del2 * np.real((ph[0] - ph[1]) * np.conj(self.Jptpc)) / M2,
            'APEgenspec': self.U * self.rd ** -2 * self.del1 * self.del2 * np.real(
                1j * self.k * (self.del1 * ph[0] + self.del2 * ph[1]) * np.conj(ph[0] - ph[1])) / M2,
            'KEfrictionspec': -self.rek * self.del2 * self.wv2 * np.abs(ph[1]) ** 2 / M2,
            'EKE': 0.5 * (self.u ** 2 + self.v ** 2).mean(axis=(-1, -2)),
            # Model._initialize_core_diagnostics: dissipation by bottom drag, enstrophy budget
            'EKEdiss': self.Hi[-1] / self.H * self.rek * (self.v[-1] ** 2 + self.u[-1] ** 2).mean(),
            'ENSflux': -(hr * np.real(np.conj(qh) * self.Jq)).sum(axis=0) / M2,
            'ENSgenspec': -(hr * np.real(self.ikQy * np.conj(qh) * ph)).sum(axis=0) / M2,
            'ENSfrictionspec': self.rek * self.del2 * self.wv2 * np.real(np.conj(qh[-1]) * ph[-1]) / M2,
        }
        # Model.dissipation_spectrum: what the exponential filter removes in the coming _forward_timestep (the diagnostics are
        # evaluated inside _step_forward after the tendencies of this step are complete)
        if self.ablevel == 0:
            dt1, dt2, dt3 = self.dt, 0.0, 0.0
        elif self.ablevel == 1 or self.useAB2:
            dt1, dt2, dt3 = 1.5 * self.dt, -0.5 * self.dt, 0.0
        else:
            dt1, dt2, dt3 = 23. / 12. * self.dt, -16. / 12. * self.dt, 5. / 12. * self.dt
        diss = (self.filtr - 1.0) * (qh + dt1 * self.dqhdt + dt2 * self.dqhdt_p + dt3 * self.dqhdt_pp)
        d['Dissspec'] = -(hr * np.real(np.conj(ph) * diss)).sum(axis=0) / self.dt / M2
        d['ENSDissspec'] = (hr * np.real(np.conj(qh) * diss)).sum(axis=0) / self.dt / M2
        if self.q_parameterization is not None and getattr(self, 'dqh', None) is not None:
            dqh = self.dqh
            d['ENSparamspec'] = np.real((hr * np.conj(qh) * dqh).sum(axis=0)) / M2
            d['paramspec'] = -np.real((hr * np.conj(ph) * dqh).sum(axis=0)) / M2
            dph = np.einsum('ij...,j...->i...', self.a, dqh)          # streamfunction tendency of the parameterization
            d['paramspec_KEflux'] = self.wv2 * (self.del1 * np.real(np.conj(ph[0]) * dph[0])
                                               + self.del2 * np.real(np.conj(ph[1]) * dph[1])) / M2
            d['paramspec_APEflux'] = self.rd ** -2 * self.del1 * self.del2 * np.real(
                np.conj(ph[0] - ph[1]) * (dph[0] - dph[1])) / M2
        return d

    def _calc_diagnostics(self):
        taveints = np.ceil(self.taveint / self.dt)
        if self.t >= self.dt and self.t >= self.tavestart and (self.tc % taveints) == 0:
            vals = self.diagnostic_fields()
            n = self.diag_count
            for k, v in vals.items():
                self.diag[k] = np.array(v, copy=True) if n == 0 else (self.diag[k] * n + v) / (n + 1)
            self.diag_count = n + 1

    # ---- qg_model.py : _calc_derived_fields ---------------------------------------------------
    def _calc_derived_fields(self):
        self.p = self.ifft(self.ph)
        self.xi = self.ifft(-self.wv2 * self.ph)
        self.Jptpc = -self._advect(self.p[0] - self.p[1], self.del1 * self.u[0] + self.del2 * self.u[1],
                                   self.del1 * self.v[0] + self.del2 * self.v[1])
        self.Jpxi = self._advect(self.xi, self.u, self.v)
        self.Jq = self._advect(self.q, self.u, self.v)

    def _advect(self, q, u=None, v=None):
        if u is None:
            u = self.u
        if v is None:
            v = self.v
        return self.ik * self.fft(u * q) + self.il * self.fft(v * q)

    def to_dataset(self):  # xarray is not installed; only out-of-scope paths need it
        raise NotImplementedError('xarray export is out of scope for the oracle')


# module layout expected by ``import pyqg`` / ``import pyqg.parameterizations as p``
class _ParamModule(object):
    Parameterization = Parameterization
    QParameterization = QParameterization
    UVParameterization = UVParameterization
    CompositeParameterization = CompositeParameterization
    WeightedParameterization = WeightedParameterization


parameterizations = _ParamModule()
