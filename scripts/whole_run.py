"""Wall clock of a whole run_simulation call (configs[2] as the reference runs it, shortened): initial condition, the coupled steps with
pyqg's daily diagnostics averaged over the second half, a snapshot (q, u, v, psi as float32 on the host) every 1000 steps, the final
dataset -- against the member-steps/s of the bare step loop bench.py times.
usage: python scripts/whole_run.py [nx=64] [members=1024] [steps=3000] [closure=gan|none]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import cnn_ref                       # random_state_dict only (synthetic weights)
from pyqg_generative_b200.tools.parameters import EDDY_PARAMS
from pyqg_generative_b200.tools.simulate import run_simulation

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
closure = sys.argv[4] if len(sys.argv) > 4 else 'gan'
p = dict(EDDY_PARAMS.nx(nx))
dt = p['dt']
p.update(members=B, tmax=steps * dt, tavestart=0.5 * steps * dt, log_level=0)
par = None
if closure == 'gan':
    from pyqg_generative_b200.models.cgan_regression import CGANRegression
    from pyqg_generative_b200.tools.cnn_tools import ChannelwiseScaler
    model = CGANRegression(folder='/nonexistent', nx=nx, precision='auto')    # the coupled run inherits the closure's precision
    model.G.load_state_dict(cnn_ref.random_state_dict(4, 2, seed=0))
    model.x_scale, model.y_scale = ChannelwiseScaler(), ChannelwiseScaler()
    model.x_scale.std = np.array([7.784383e-06, 1.0471941e-06], 'float32').reshape(1, 2, 1, 1)
    model.y_scale.std = np.array([7.606111e-12, 1.656513e-13], 'float32').reshape(1, 2, 1, 1)
    par = dict(self=model, sampling='constant', nsteps=1)
out = {}
for rep in range(2):                              # the first call pays allocations, calibration and graph capture
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ds = run_simulation(p, par, rng=np.random.RandomState(rep))
    torch.cuda.synchronize()
    out['wall_s_call%d' % rep] = time.perf_counter() - t0
q = ds['q'] if isinstance(ds, dict) else ds['q'].values
out.update(nx=nx, members=B, steps=steps, closure=closure, snapshots=int(np.shape(q)[1]), finite=bool(np.isfinite(q).all()),
           member_steps_per_s=B * steps / out['wall_s_call1'], has_spectra='KEflux' in ds)
print(json.dumps(out))
