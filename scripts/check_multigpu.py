"""Multi-GPU check (run under torchrun, one rank per GPU): members sharded over ranks + NCCL reduction of the online
diagnostics must equal the single-GPU ensemble (Philox keyed by the GLOBAL member id -> sharding-invariant)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from oracle import cnn_ref
from pyqg_generative_b200 import parallel
from pyqg_generative_b200.models.cgan_regression import CGANRegression
from pyqg_generative_b200.tools.cnn_tools import ChannelwiseScaler
from pyqg_generative_b200.tools.stochastic_pyqg import stochastic_QGModel

rank, world, local = parallel.init_from_env()
torch.cuda.set_device(local)
TOTAL, N, STEPS = 16, 64, 12
sd = cnn_ref.random_state_dict(4, 2, seed=0)
q_all = bench.synthetic_states(TOTAL, N, 99)


def run(count, offset):
    gan = CGANRegression(folder='/nonexistent', nx=N, precision='tc')
    gan.G.load_state_dict(sd)
    gan.x_scale, gan.y_scale = ChannelwiseScaler(), ChannelwiseScaler()
    gan.x_scale.std = np.array(bench.X_STD, 'float32').reshape(1, 2, 1, 1)
    gan.y_scale.std = np.array(bench.Y_STD, 'float32').reshape(1, 2, 1, 1)
    m = stochastic_QGModel(dict(nx=N, dt=14400., log_level=0, tmax=1e12, tavestart=4 * 14400., taveint=2 * 14400.,
                                members=count, member_offset=offset, device=local, parameterization=gan,
                                precision='tc', seed=5), 'AR1', 2)
    m.squeeze = False
    m.set_q(q_all[offset:offset + count])
    m._step_forward(STEPS)
    return m


count, offset = parallel.shard_members(TOTAL, rank, world)
m = run(count, offset)
ke, en, n = parallel.ensemble_spectra(m.spectra_sums())
diag, nd = parallel.ensemble_diagnostics(m.diagnostic_sums())       # all time-averaged pyqg diagnostics in one all-reduce
kebar = parallel.ensemble_ke(m.diagnostics()[0])
if rank == 0:
    ref = run(TOTAL, 0)
    ke0, en0, n0 = ref.spectra_sums()
    assert n == n0 == nd == TOTAL * 4, (n, n0, nd)      # sampled before the steps starting at tc = 4, 6, 8, 10
    err = np.abs(ke - ke0 / n0).max() / np.abs(ke0 / n0).max()
    d0 = ref.averaged_diagnostics()
    for k in ('KEflux', 'APEflux', 'APEgenspec', 'KEfrictionspec', 'paramspec'):
        e = np.abs(diag[k] - d0[k]).max() / np.abs(d0[k]).max()
        assert e < 1e-11, (k, e)
    print('world %d: budget terms of the sharded ensemble equal the single-GPU ensemble to 1e-11' % world)
    errq = np.abs(ref.q[offset:offset + count] - m.q).max()
    print('world %d: ensemble KEspec rel diff vs single GPU %.2e, ensemble KE %.6e vs %.6e, shard state diff %.1e'
          % (world, err, kebar, ref.diagnostics()[0].mean(), errq))
    assert err < 1e-12 and errq == 0.0
    print('multi-GPU check ok')
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
