"""Index arithmetic of the CUDA phase programs (csrc/qg_core.cuh: packed FFTs, Hermitian packing, AB3 update),
executed thread-by-thread on the host by tests/emu and compared with the oracle.  The product never uses the
emulation; it only de-risks the kernels before GPU time is spent."""
import ctypes

import numpy as np
import pytest

from oracle import pyqg_shim


class Cfg(ctypes.Structure):
    _fields_ = [('nx', ctypes.c_int32), ('members', ctypes.c_int32), ('member_offset', ctypes.c_int32),
                ('device', ctypes.c_int32)] + [(n, ctypes.c_double) for n in
                                               'L dt rek filterfac beta rd delta H1 U1 U2'.split()]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def run(lib, cfg, prog, qh, q, dcur=None, dp=None, dpp=None, dq=None, x=None, ab=2, ph=None, u=None, v=None, p=None,
        red=None, nt=96, bud=None, scr=None, demean=1):
    lib.qgbemu_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 7 + \
        [ctypes.c_float, ctypes.c_float, ctypes.c_int] + [ctypes.c_void_p] * 7 + [ctypes.c_int]
    assert lib.qgbemu_run(ctypes.byref(cfg), prog, nt, _p(qh), _p(q), _p(dcur), _p(dp), _p(dpp), _p(dq), _p(x),
                          7.78e-6, 1.05e-6, ab, _p(ph), _p(u), _p(v), _p(p), _p(red), _p(bud), _p(scr), demean) == 0


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize('N,dt,jet', [(64, 14400., False), (48, 7200., True), (96, 7200., False), (32, 14400., False),
                                      (128, 7200., True)])
def test_programs_match_oracle(emu_lib, N, dt, jet):
    rng = np.random.RandomState(N)
    phys = dict(rek=7e-8, delta=0.1, beta=1e-11) if jet else dict(rek=5.787e-7, delta=0.25, beta=1.5e-11)
    m = pyqg_shim.QGModel(nx=N, dt=dt, log_level=0, **phys)
    cfg = Cfg(N, 2, 0, 0, 1e6, dt, phys['rek'], 23.6, phys['beta'], 15000., phys['delta'], 500., 0.025, 0.)
    B = 2
    # white noise (full Nyquist content) exercises the self-conjugate columns of the packed transforms
    q0 = rng.randn(B, 2, N, N) * np.array([7e-6, 1e-6])[None, :, None, None]
    qh = np.zeros((B, 2, N, N // 2 + 1), complex)
    q = q0.copy()
    run(emu_lib, cfg, 2, qh, q)                                             # PROG_SET_Q
    m.q = q0[1]
    assert rel(qh[1], m.qh) < 1e-14
    m._invert()
    m._calc_derived_fields()
    ph = np.zeros_like(qh)
    u, v, p = np.zeros_like(q), np.zeros_like(q), np.zeros_like(q)
    run(emu_lib, cfg, 3, qh, q, ph=ph, u=u, v=v, p=p)                        # PROG_INVERT
    assert rel(ph[1], m.ph) < 1e-14 and rel(u[1], m.u) < 1e-14 and rel(v[1], m.v) < 1e-14 and rel(p[1], m.p) < 1e-14

    class Par(pyqg_shim.QParameterization):
        def __call__(self, mm):
            return self.dq
    par = Par()
    m.q_parameterization = par
    hist = [np.zeros_like(qh) for _ in range(3)]
    for step in range(4):                                                   # Euler, AB2, AB3, AB3
        dq = rng.randn(B, 2, N, N) * np.array([7e-12, 2e-13])[None, :, None, None]
        par.dq = dq[1] - dq[1].mean(axis=(1, 2), keepdims=True)              # models/parameterization.py:25
        x = np.zeros((B, 2, N, N), np.float32)
        run(emu_lib, cfg, 1, qh, q, hist[step % 3], hist[(step + 2) % 3], hist[(step + 1) % 3], dq=dq.copy(), x=x,
            ab=min(step, 2))                                                 # PROG_STEP_DQ
        m._step_forward()
        assert rel(qh[1], m.qh) < 1e-13 and rel(q[1], m.q) < 1e-13 and rel(hist[step % 3][1], m.dqhdt_p) < 1e-13
        xr = m.q.astype('float32') / np.array([7.78e-6, 1.05e-6], 'float32')[:, None, None]
        assert np.abs(x[1] - xr).max() <= 2e-7 * np.abs(xr).max()
    red = np.zeros((B, 4))
    run(emu_lib, cfg, 4, qh, q, red=red)                                      # PROG_DIAG
    m._invert()
    assert abs(red[1, 0] - m._calc_ke()) < 1e-12 * m._calc_ke()
    assert abs(max(red[1, 1], red[1, 2]) * dt / m.dx - m._calc_cfl()) < 1e-12


def test_raw_forcing_program_keeps_the_mean(emu_lib):
    N, dt = 32, 14400.
    rng = np.random.RandomState(3)
    m = pyqg_shim.QGModel(nx=N, dt=dt, log_level=0)
    cfg = Cfg(N, 1, 0, 0, 1e6, dt, 5.787e-7, 23.6, 1.5e-11, 15000., 0.25, 500., 0.025, 0.)
    q0 = rng.randn(1, 2, N, N) * 1e-6
    qh = np.zeros((1, 2, N, N // 2 + 1), complex)
    q = q0.copy()
    run(emu_lib, cfg, 2, qh, q)
    dq = rng.randn(1, 2, N, N) * 1e-12 + 3e-12

    class Par(pyqg_shim.QParameterization):
        def __call__(self, mm):
            return dq[0]
    m.q = q0[0]
    m.q_parameterization = Par()
    hist = [np.zeros_like(qh) for _ in range(3)]
    run(emu_lib, cfg, 6, qh, q, hist[0], hist[2], hist[1], dq=dq.copy(), ab=0)   # PROG_STEP_DQ_RAW
    m._step_forward()
    assert rel(qh[0], m.qh) < 1e-13
    assert abs(q[0, 0].mean() - m.q[0].mean()) < 1e-20


BUDGET_TERMS = ['KEflux', 'APEflux', 'APEgenspec', 'KEfrictionspec', 'entspec', 'paramspec_KEflux', 'paramspec_APEflux',
                'ENSflux', 'ENSgenspec', 'ENSfrictionspec', 'Dissspec', 'ENSDissspec', 'ENSparamspec']


@pytest.mark.parametrize('N,jet', [(64, False), (48, True), (32, False)])
def test_budget_program_matches_oracle(emu_lib, N, jet):
    """PROG_BUDGET (spectral energy budget terms) against the oracle's restatement of the pyqg diagnostics."""
    rng = np.random.RandomState(7 + N)
    dt = 7200.
    phys = dict(rek=7e-8, delta=0.1, beta=1e-11) if jet else dict(rek=5.787e-7, delta=0.25, beta=1.5e-11)
    B = 2
    dq = rng.randn(B, 2, N, N) * np.array([7e-12, 2e-13])[None, :, None, None]

    class Par(pyqg_shim.QParameterization):
        def __call__(self, mm):
            return dq[1] - dq[1].mean(axis=(1, 2), keepdims=True)
    m = pyqg_shim.QGModel(nx=N, dt=dt, log_level=0, q_parameterization=Par(), **phys)
    cfg = Cfg(N, B, 0, 0, 1e6, dt, phys['rek'], 23.6, phys['beta'], 15000., phys['delta'], 500., 0.025, 0.)
    q0 = rng.randn(B, 2, N, N) * np.array([7e-6, 1e-6])[None, :, None, None]
    qh = np.zeros((B, 2, N, N // 2 + 1), complex)
    q = q0.copy()
    run(emu_lib, cfg, 2, qh, q)
    bud = np.zeros((B, len(BUDGET_TERMS), N, N // 2 + 1))
    scr = np.zeros((B, 3, N, N))
    # tendency history of an AB3 step (the filter-dissipation spectra describe the coming _forward_timestep)
    dp = (rng.randn(B, 2, N, N // 2 + 1) + 1j * rng.randn(B, 2, N, N // 2 + 1)) * 3e-10
    dpp = (rng.randn(B, 2, N, N // 2 + 1) + 1j * rng.randn(B, 2, N, N // 2 + 1)) * 3e-10
    run(emu_lib, cfg, 9, qh, q, dp=dp, dpp=dpp, dq=dq.copy(), bud=bud, scr=scr, ab=2)   # PROG_BUDGET
    m.q = q0[1]
    m.dqhdt_p, m.dqhdt_pp, m.ablevel = dp[1].copy(), dpp[1].copy(), 2
    m._invert()
    m._do_advection()
    m._do_friction()
    m._do_q_subgrid_parameterization()
    d = m.diagnostic_fields()
    for i, name in enumerate(BUDGET_TERMS):
        assert rel(bud[1, i], d[name]) < 1e-12, name
    assert rel(bud[1, 5] + bud[1, 6], d['paramspec']) < 1e-12
