"""world_size=2 gloo test of the multi-GPU host logic: member sharding + the diagnostics all-reduce."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    from pyqg_generative_b200 import parallel
    r, w, _ = parallel.init_from_env(backend='gloo')
    count, offset = parallel.shard_members(total, r, w)
    # fake per-member spectra keyed by the GLOBAL member id: the reduction must not depend on the sharding
    members = np.arange(offset, offset + count)
    ke = sum(np.full((2, 8, 5), float(m + 1)) for m in members)
    en = sum(np.full((2, 8, 5), float(m + 1) ** 2) for m in members)
    kem, enm, n = parallel.ensemble_spectra((ke, en, count))
    kbar = parallel.ensemble_ke(np.array([float(m) for m in members]))
    tmax = parallel.allreduce_max(10.0 + r)
    # the full pyqg diagnostic set travels as ONE bucket: dict of sums + count -> dict of ensemble means (+ paramspec)
    sums = {'KEspec': ke, 'KEflux': ke[0] * 2, 'paramspec_KEflux': ke[0], 'paramspec_APEflux': -0.25 * ke[0]}
    dmean, dn = parallel.ensemble_diagnostics((sums, count))
    assert dn == n and abs(dmean['KEflux'][0, 0] - 2 * kem[0, 0, 0]) < 1e-12
    assert abs(dmean['paramspec'][0, 0] - 0.75 * kem[0, 0, 0]) < 1e-12 and dmean['KEspec'].shape == (2, 8, 5)
    out[rank] = (kem[0, 0, 0], enm[0, 0, 0], n, kbar, tmax, count, offset)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_diagnostics_reduce_to_the_ensemble_mean():
    total, world = 7, 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), total, out), nprocs=world, join=True)
    ids = np.arange(total)
    for r in range(world):
        kem, enm, n, kbar, tmax, count, offset = out[r]
        assert n == total
        assert abs(kem - (ids + 1).mean()) < 1e-12 and abs(enm - ((ids + 1.0) ** 2).mean()) < 1e-12
        assert abs(kbar - ids.mean()) < 1e-12 and tmax == 11.0
    assert out[0][5] + out[1][5] == total and out[1][6] == out[0][5]
