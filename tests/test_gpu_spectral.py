"""GPU parity of the fused spectral step (csrc/qg_core.cuh + spectral.cuh) against the oracle, through the C ABI.
Tolerance: BASELINE.json north_star -- spectral-state relative error <= 1e-10 per step in fp64."""
import numpy as np
import pytest

from oracle import operators_ref as opr
from oracle import pyqg_shim

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def make(nx, members, **kw):
    from pyqg_generative_b200.tools.stochastic_pyqg import EnsembleQGModel
    kw.setdefault('log_level', 0)
    return EnsembleQGModel(members=members, nx=nx, **kw)


@pytest.mark.parametrize('N,dt,phys', [
    (64, 14400., {}), (48, 14400., dict(rek=7e-8, delta=0.1, beta=1e-11)), (96, 7200., {}), (32, 14400., {}),
    (128, 7200., {}), (256, 3600., dict(rek=7e-8, delta=0.1, beta=1e-11))])     # 128/256: thread-block-cluster path
def test_set_q_invert_and_steps_match_oracle(N, dt, phys):
    rng = np.random.RandomState(N)
    B = 3
    m = make(N, B, dt=dt, **phys)
    q0 = rng.randn(B, 2, N, N) * np.array([7e-6, 1e-6])[None, :, None, None]   # white noise: full Nyquist content
    m.q = q0
    refs = []
    for b in range(B):
        o = pyqg_shim.QGModel(nx=N, dt=dt, log_level=0, **phys)
        o.q = q0[b]
        o._invert()
        o._calc_derived_fields()
        refs.append(o)
    assert np.array_equal(m.q, q0)                      # q setter round trip is exact (notebook 3-2-dealiasing:88)
    m._invert()
    qh, ph, u, v, p = m.qh, m.ph, m.u, m.v, m.p
    for b, o in enumerate(refs):
        assert rel(qh[b], o.qh) < TOL and rel(ph[b], o.ph) < TOL
        assert rel(u[b], o.u) < TOL and rel(v[b], o.v) < TOL and rel(p[b], o.p) < TOL
    for step in range(5):                               # Euler, AB2, AB3, AB3, AB3
        m._step_forward()
        q, qh, d = m.q, m.qh, m.dqhdt
        for b, o in enumerate(refs):
            o._step_forward()
            assert rel(q[b], o.q) < TOL and rel(qh[b], o.qh) < TOL and rel(d[b], o.dqhdt_p) < TOL, (step, b)
    assert m.tc == 5 and abs(m.t - 5 * dt) < 1e-6
    ke, cfl, flags = m.diagnostics()
    for b, o in enumerate(refs):
        o._invert()
        assert abs(ke[b] - o._calc_ke()) < 1e-10 * o._calc_ke() and abs(cfl[b] - o._calc_cfl()) < 1e-10
    assert not flags.any()


def test_many_steps_in_one_call_equal_single_steps_and_oracle_ke():
    """1500 free-running steps from the JAMES initial condition: the linear-instability phase is not chaotic yet
    (SURVEY.md section 7 'Chaos'), so the state itself must still agree with the oracle."""
    N, dt = 64, 14400.
    np.random.seed(5)
    o = pyqg_shim.QGModel(nx=N, dt=dt, log_level=0)
    opr.set_initial_condition(o)
    m = make(N, 2, dt=dt)
    m.q = o.q
    m._step_forward(1500)
    for _ in range(1500):
        o._step_forward()
    assert m.tc == 1500
    q = m.q
    assert np.array_equal(q[0], q[1])
    assert rel(q[0], o.q) < 1e-8
    o._invert()
    assert abs(m.diagnostics()[0][0] - o._calc_ke()) < 1e-9 * o._calc_ke()


@pytest.mark.parametrize('N,dt', [(32, 14400.), (64, 14400.), (128, 7200.), (256, 3600.)])   # generic, register-FFT and cluster kernels
def test_host_callback_parameterization_and_weighting(N, dt):
    from pyqg_generative_b200.models.parameterization import QParameterization
    rng = np.random.RandomState(2)
    dq = rng.randn(2, N, N) * 1e-12 + 2e-12

    class Const(QParameterization):
        def __call__(self, mm):
            assert np.asarray(mm.q).shape[-3:] == (2, N, N)
            return dq
    q0 = rng.randn(2, N, N) * 1e-6
    m = make(N, 2, dt=dt, parameterization=0.5 * Const())
    m.q = q0
    o = pyqg_shim.QGModel(nx=N, dt=dt, log_level=0, parameterization=0.5 * _OracleConst(dq))
    o.q = q0
    for _ in range(3):
        m._step_forward()
        o._step_forward()
    assert rel(m.q[1], o.q) < TOL


class _OracleConst(pyqg_shim.QParameterization):
    def __init__(self, dq):
        self.dq = dq

    def __call__(self, m):
        return self.dq


def test_run_with_snapshots_cadence_and_log():
    N, dt = 32, 14400.
    m = make(N, 2, dt=dt, tmax=40 * dt, twrite=10, log_level=1, tavestart=20 * dt, taveint=5 * dt)
    np.random.seed(0)
    from pyqg_generative_b200.tools.simulate import set_initial_condition
    set_initial_condition(m)
    times = [t for t in m.run_with_snapshots(tsnapint=8 * dt)]
    assert np.allclose(times, [8 * dt * i for i in range(1, 6)])
    assert [s for s, _, _, _ in m.log] == [10, 20, 30, 40]
    ke, en, count = m.spectra_sums()
    # pyqg samples BEFORE the step when t >= tavestart and tc % taveints == 0: tc = 20, 25, 30, 35 (no step starts at 40)
    assert count == 2 * 4 and ke.shape == (2, N, N // 2 + 1) and (ke >= 0).all()
    kespec = 0
    for b in range(2):
        o = pyqg_shim.QGModel(nx=N, dt=dt, log_level=0)
        o.q = m.q[b]
        o._invert()
        kespec = kespec + o.wv2 * np.abs(o.ph) ** 2 / o.M ** 2
    from pyqg_generative_b200 import _lib
    k1, e1 = np.empty(ke.size), np.empty(ke.size)
    _lib.check(m._lib.qgb_diag_spectra(m._h, k1.ctypes.data, e1.ctypes.data, 0, m._stream()), m._h)
    assert rel(k1.reshape(ke.shape), kespec) < 1e-10


def test_blow_up_is_flagged_not_fatal():
    N = 32
    m = make(N, 3, dt=14400.)
    q = np.random.RandomState(0).randn(3, 2, N, N) * 1e-6
    q[1] *= 1e6            # absurd amplitude: CFL >> 1, goes non-finite within a few steps
    m.q = q
    m._step_forward(60)
    ke, cfl, flags = m.diagnostics()
    assert flags[1] != 0 and flags[0] == 0 and flags[2] == 0
    assert np.isfinite(ke[0]) and np.isfinite(ke[2])


def test_full_size_ensemble_members_are_independent_and_identical():
    """BASELINE config size (1024 members at 64^2): replicated initial condition -> every member must produce the same
    bits, and member 0 must match the oracle."""
    N, dt, B = 64, 14400., 1024
    np.random.seed(7)
    o = pyqg_shim.QGModel(nx=N, dt=dt, log_level=0)
    opr.set_initial_condition(o)
    m = make(N, B, dt=dt)
    m.q = o.q
    m._step_forward(3)
    for _ in range(3):
        o._step_forward()
    q = m.q
    assert rel(q[0], o.q) < TOL
    assert (q == q[0][None]).all()


def test_host_buffer_stepping_sync_and_pipelined():
    """qgb_step_host / qgb_step_host_async (the end-to-end path of bench.py) equal set_q + step + get."""
    import torch
    N, dt, B = 48, 14400., 4
    rng = np.random.RandomState(4)
    q0 = rng.randn(B, 2, N, N) * 1e-6
    ref = make(N, B, dt=dt)
    ref.q = q0
    ref._step_forward(3)
    a = make(N, B, dt=dt)
    qout = np.empty_like(q0)
    a.step_host(q0, qout, 3)
    assert np.array_equal(qout, ref.q) and a.tc == 3
    # two member groups on two streams, pinned buffers, no host synchronisation in between
    groups = []
    for g in range(2):
        mg = make(N, 2, dt=dt)
        qi = torch.from_numpy(q0[2 * g:2 * g + 2].copy()).pin_memory()
        qo = torch.empty_like(qi).pin_memory()
        groups.append((mg, qi, qo, torch.cuda.Stream()))
    for mg, qi, qo, st in groups:
        mg.step_host(qi, qo, 3, stream=st, wait=False)
    for mg, qi, qo, st in groups:
        st.synchronize()
    got = np.concatenate([g[2].numpy() for g in groups])
    assert np.array_equal(got, ref.q)
    # float32 snapshots converted on the device (qgb_get_f32): the same bits as the host-side cast the reference does
    # (drop_vars, tools/simulate.py:16-36), synchronously and enqueued on a stream without a q_out
    ref._invert()
    for name, full in (('q', ref.q), ('u', ref.u), ('v', ref.v), ('p', ref.p)):
        assert np.array_equal(ref.real32(name), full.astype('float32')), name
    mg, qi, qo, st = groups[0]
    q32 = torch.empty(qi.shape, dtype=torch.float32).pin_memory()
    mg.step_host(qo, None, 2, stream=st, wait=False)
    mg.real32('q', out=q32, stream=st, wait=False)
    st.synchronize()
    ref._step_forward(2)
    assert np.array_equal(q32.numpy(), ref.q[:2].astype('float32'))


BUDGET = ('KEflux', 'APEflux', 'APEgenspec', 'KEfrictionspec', 'entspec', 'paramspec_KEflux', 'paramspec_APEflux',
          'ENSflux', 'ENSgenspec', 'ENSfrictionspec', 'Dissspec', 'ENSDissspec', 'ENSparamspec')


@pytest.mark.parametrize('N,dt,phys', [(64, 14400., {}), (48, 7200., dict(rek=7e-8, delta=0.1, beta=1e-11)), (128, 7200., {}),
                                      (256, 3600., dict(rek=7e-8, delta=0.1, beta=1e-11)), (96, 7200., {}), (32, 14400., {})])
def test_spectral_energy_budget_matches_oracle(N, dt, phys):
    """qgb_diag_budget (PROG_BUDGET) against the oracle's restatement of the pyqg diagnostics, with an
    external forcing standing in for the closure output (paramspec terms).  Sum over members, tolerance 1e-10."""
    rng = np.random.RandomState(N + 1)
    B = 3
    m = make(N, B, dt=dt, **phys)
    q0 = rng.randn(B, 2, N, N) * np.array([7e-6, 1e-6])[None, :, None, None]
    dq = rng.randn(2, N, N) * np.array([7e-12, 2e-13])[:, None, None]
    dq -= dq.mean(axis=(1, 2), keepdims=True)
    m.q = q0
    m.set_parameterization(_OracleConst(dq), 'AR1', 1)      # host callback -> qgb_set_forcing -> PROG_STEP_DQ_RAW
    m._step_forward()
    terms = m.budget_sums()
    ref = {k: 0 for k in BUDGET + ('paramspec',)}
    for b in range(B):
        # the oracle takes the same step (the dissipation spectra need its tendency history and Adams-Bashforth level), then
        # evaluates the diagnostics where pyqg does: after the tendencies of the NEXT step are complete
        o = pyqg_shim.QGModel(nx=N, dt=dt, log_level=0, q_parameterization=_OracleConst(dq), **phys)
        o.q = q0[b]
        o._step_forward()
        o._invert()
        o._do_advection()
        o._do_friction()
        o._do_q_subgrid_parameterization()
        d = o.diagnostic_fields()
        for k in ref:
            ref[k] = ref[k] + d[k]
    for k in BUDGET:
        assert rel(terms[k], ref[k]) < TOL, k
    assert rel(terms['paramspec_KEflux'] + terms['paramspec_APEflux'], ref['paramspec']) < TOL


def test_time_averaged_diagnostics_follow_pyqg_sampling():
    """Device-side running sums (qgb_diag_config / qgb_diag_averages) against the oracle's _calc_diagnostics running mean."""
    N, dt, B = 32, 14400., 2
    rng = np.random.RandomState(5)
    kw = dict(dt=dt, tavestart=2 * dt, taveint=2 * dt, tmax=1e12)
    m = make(N, B, **kw)
    q0 = rng.randn(B, 2, N, N) * np.array([7e-6, 1e-6])[None, :, None, None]
    m.q = q0
    m._step_forward(9)
    avg = m.averaged_diagnostics()
    refs = []
    for b in range(B):
        o = pyqg_shim.QGModel(nx=N, log_level=0, **kw)
        o.q = q0[b]
        for _ in range(9):
            o._step_forward()
        refs.append(o)
    assert m.diag_count == refs[0].diag_count == 4           # sampled before the steps starting at tc = 2, 4, 6, 8
    for k in ('KEspec', 'Ensspec', 'KEflux', 'APEflux', 'APEgenspec', 'KEfrictionspec', 'entspec', 'ENSflux', 'ENSgenspec',
              'ENSfrictionspec', 'Dissspec', 'ENSDissspec', 'EKE', 'EKEdiss'):
        ref = sum(o.diag[k] for o in refs) / B
        assert rel(avg[k], ref) < 1e-9, k
    assert np.abs(avg['paramspec']).max() == 0.0 and np.abs(avg['ENSparamspec']).max() == 0.0
