"""Throughput of the other BASELINE.json configurations (not the headline bench line): one JSON line per config."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from oracle import cnn_ref
from pyqg_generative_b200.tools.cnn_tools import ChannelwiseScaler
from pyqg_generative_b200.tools.stochastic_pyqg import stochastic_QGModel
from pyqg_generative_b200.models.cvae_regression import CVAERegression
from pyqg_generative_b200.models.mean_var_model import MeanVarModel
from pyqg_generative_b200.tools.parameters import EDDY_PARAMS, JET_PARAMS


def scal(m):
    m.x_scale, m.y_scale = ChannelwiseScaler(), ChannelwiseScaler()
    m.x_scale.std = np.array(bench.X_STD, 'float32').reshape(1, 2, 1, 1)
    m.y_scale.std = np.array(bench.Y_STD, 'float32').reshape(1, 2, 1, 1)
    return m


def run(name, nx, members, closure, base, steps=20, precision='tc'):
    p = dict(base.nx(nx))
    p.update(dict(log_level=0, tmax=1e12, tavestart=1e12, members=members, precision=precision, seed=1))
    par = None
    if closure == 'vae':
        par = scal(CVAERegression(folder='/nonexistent', precision=precision))
        par.decoder.load_state_dict(cnn_ref.random_state_dict(4, 2, seed=1))
    elif closure == 'gz':
        par = scal(MeanVarModel(folder='/nonexistent', precision=precision))
        par.net_mean.load_state_dict(cnn_ref.random_state_dict(2, 2, seed=2))
        par.net_var.load_state_dict(cnn_ref.random_state_dict(2, 2, seed=3))
    if par is not None:
        p['parameterization'] = par
    m = stochastic_QGModel(p, 'constant', 1)
    m.squeeze = False
    m.set_q(bench.synthetic_states(members, nx, 7))
    m._step_forward(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m._step_forward(steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ke, cfl, flags = m.diagnostics()
    print(json.dumps({'config': name, 'nx': nx, 'members': members, 'closure': closure, 'precision': precision if closure else None,
                      'dt': p['dt'], 'member_steps_per_s': members * steps / ms * 1e3, 'ms_per_step': ms / steps,
                      'healthy': bool(np.isfinite(ke).all() and not flags.any())}))


def run_forcing(members=64, steps=2000, every=1000):
    """configs[4]: hi-res 256^2 ensemble coarse-grained to 64^2 by Operator1/Operator2 with the subgrid forcing S every
    ``every`` steps (tools/simulate.py generate_subgrid_forcing, run_forcing_datasets.py); wall clock incl. host copies."""
    from pyqg_generative_b200.tools import operators as ops
    from pyqg_generative_b200.tools.simulate import generate_subgrid_forcing
    p = dict(EDDY_PARAMS.nx(256))
    p.update(dict(log_level=0, tmax=steps * p['dt'], tavestart=1e12, members=members))
    generate_subgrid_forcing([64], dict(p, tmax=2 * every * p['dt'] if False else every * p['dt']), every * p['dt'],
                             [ops.Operator1, ops.Operator2], 'none', np.random.RandomState(0))      # warm-up (handles, tables)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = generate_subgrid_forcing([64], p, every * p['dt'], [ops.Operator1, ops.Operator2], 'none', np.random.RandomState(0))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    key = sorted(out)[0]
    print(json.dumps({'config': 'configs[4] forcing datasets 256->64, Operator1+Operator2, S every %d steps' % every, 'nx': 256,
                      'members': members, 'steps': steps, 'snapshots': int(out[key]['q'].shape[1]), 'wall_s': round(dt, 3),
                      'member_steps_per_s': round(members * steps / dt, 1), 'keys': sorted(out)}))


if __name__ == '__main__':
    run_forcing()
    run('configs[1] hi-res reference ensemble', 256, 64, None, EDDY_PARAMS)
    run('hi-res 128', 128, 64, None, EDDY_PARAMS)
    run('configs[0]-like, no closure', 64, 1024, None, EDDY_PARAMS)
    for nx in (64, 48, 96):
        run('configs[3] jet + CVAE', nx, 512, 'vae', JET_PARAMS)
        run('configs[3] jet + GZ', nx, 512, 'gz', JET_PARAMS)
