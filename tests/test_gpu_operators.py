"""GPU parity of the coarse-graining path (qgb_operator / qgb_subgrid_forcing) against the reference outputs committed in
tests/golden/operators_128.npz (produced by the unmodified reference functions) and the oracle at 256 -> 64."""
import numpy as np
import pytest

from conftest import golden
from oracle import operators_ref as opr

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max()


def test_operators_match_reference_outputs():
    from pyqg_generative_b200.tools import operators as ops
    o = golden('operators_128.npz')
    q = o['q'].astype('float64')
    for nc in (32, 48, 64):
        for name in ('Operator1', 'Operator2', 'Operator5', 'cut_off'):
            y = getattr(ops, name)(q, nc)
            assert y.shape == (2, nc, nc)
            assert rel(y, o['%s_%d' % (name, nc)]) < TOL, (name, nc)
    y2d = ops.Operator1(q[0], 48)                                    # 2-D input like the reference's numpy branch
    assert y2d.shape == (48, 48) and rel(y2d, o['Operator1_48'][0]) < TOL
    with pytest.raises(ValueError, match='nc must be even'):
        ops.cut_off(q, 33)


def test_subgrid_forcing_matches_reference_outputs():
    from pyqg_generative_b200.tools import operators as ops
    o = golden('operators_128.npz')
    q = o['q'].astype('float64')
    params = dict(dt=14400.0, tmax=1.0, tavestart=1.0)               # make_golden: EDDY_PARAMS.nx(128) without nx
    for name in ('Operator1', 'Operator2', 'Operator5'):
        forcing, mf, _ = ops.PV_subgrid_forcing(q, 64, getattr(ops, name), params)
        assert rel(forcing, o['S_%s_none' % name]) < 1e-9, name
        assert rel(mf.q, o['qf_%s' % name]) < TOL and rel(mf.u, o['uf_%s' % name]) < TOL
        assert rel(mf.v, o['vf_%s' % name]) < TOL and rel(mf.p, o['pf_%s' % name]) < TOL
        f32, _, _ = ops.PV_subgrid_forcing(q, 64, getattr(ops, name), params, dealias='3/2-rule')
        assert rel(f32, o['S_%s_32' % name]) < 1e-9, name              # 3/2-rule: 128 -> 192 and 64 -> 96 and back
    with pytest.raises(ValueError, match='dealias should be'):
        ops.PV_subgrid_forcing(q, 64, ops.Operator1, params, dealias='1/2-rule')
    assert rel(ops.fft_interpolate(o['interp_in'], 48, 72), o['interp_48_72']) < TOL
    assert rel(ops.fft_interpolate(o['interp_in'], 48, 32), o['interp_48_32']) < TOL
    x = np.random.RandomState(1).randn(64, 64)                        # notebooks/3-2-dealiasing.ipynb:586
    assert rel(ops.cut_off(x, 16), ops.fft_interpolate(x, 64, 16)) < 1e-13


def test_operator4_and_two_thirds_rule_match_reference_outputs():
    """Operator4 = model_filter(Operator2) (tools/operators.py:213-214) and advect(..., '2/3-rule') (:253-257)."""
    from pyqg_generative_b200.tools import operators as ops
    q = golden('operators_128.npz')['q'].astype('float64')
    o = golden('operators_128_more.npz')
    params = dict(dt=14400.0, tmax=1.0, tavestart=1.0)
    for nc in (32, 48, 64):
        assert rel(ops.Operator4(q, nc), o['Operator4_%d' % nc]) < TOL, nc
    for name in ('Operator1', 'Operator2', 'Operator4', 'Operator5'):
        f23, _, _ = ops.PV_subgrid_forcing(q, 64, getattr(ops, name), params, dealias='2/3-rule')
        assert rel(f23, o['S_%s_23' % name]) < 1e-9, name
    f, mf, _ = ops.PV_subgrid_forcing(q, 64, ops.Operator4, params)
    assert rel(f, o['S_Operator4_none']) < 1e-9
    assert rel(mf.q, o['qf_Operator4']) < TOL and rel(mf.u, o['uf_Operator4']) < TOL
    f, _, _ = ops.PV_subgrid_forcing(q, 64, ops.Operator4, params, dealias='3/2-rule')
    assert rel(f, o['S_Operator4_32']) < 1e-9


def test_hires_256_to_64_batched_on_device():
    """configs[4] shape: 256^2 hi-res snapshots coarse-grained to 64^2, batched, device tensors in and out."""
    import torch
    from pyqg_generative_b200.tools import operators as ops
    rng = np.random.RandomState(0)
    B = 3
    q = rng.randn(B, 2, 256, 256) * np.array([7e-6, 1e-6])[None, :, None, None]
    qd = torch.as_tensor(q).cuda()
    for name in ('Operator1', 'Operator2'):
        y = getattr(ops, name)(qd, 64)
        assert y.is_cuda and tuple(y.shape) == (B, 2, 64, 64)
        ref = np.stack([getattr(opr, name)(q[b], 64) for b in range(B)])
        assert rel(y.cpu().numpy(), ref) < TOL, name
    jet = dict(rek=7e-8, delta=0.1, beta=1e-11)
    forcing, fields = ops.PV_subgrid_forcing(qd, 64, ops.Operator2, jet, return_fields=True)
    for b in range(B):
        f, mf, m = opr.PV_subgrid_forcing(q[b], 64, opr.Operator2, jet)
        assert rel(forcing[b], f) < 1e-9
        assert rel(fields['u'][b], mf.u) < TOL and rel(fields['psi'][b], mf.p) < TOL
    # mean is preserved by the operators (cut_off divides by ratio^2)
    assert abs(ops.Operator1(q[0, 0] + 3e-6, 64).mean() - (q[0, 0].mean() + 3e-6)) < 1e-18
