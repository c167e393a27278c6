"""Multi-GPU orchestration: ensemble members are sharded over ranks (one process per GPU); nothing is exchanged on the
step path.  The only collective is the reduction of online diagnostics (KE(t), KE / enstrophy spectra accumulators) at
output cadence -- NCCL over NVLink on GPUs, gloo in the CPU tests.

The reference has no counterpart: its ensembles are SLURM arrays of single-member processes gathered through the file
system (scripts/run_parameterized.py:55-63, tools/comparison_tools.py:423).
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def shard_members(total, rank, world):
    """Contiguous block partition: rank r owns members [offset, offset+count).  Sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return count, offset


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).  Returns (rank, world, local)."""
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def _device():
    if dist.is_initialized() and dist.get_backend() == 'nccl':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def allreduce_sum(arrays):
    """Sum a list of float64 numpy arrays over all ranks with ONE collective (flattened bucket)."""
    arrays = [np.asarray(a, dtype=np.float64) for a in arrays]
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return arrays
    flat = torch.from_numpy(np.concatenate([a.ravel() for a in arrays])).to(_device())
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat = flat.cpu().numpy()
    out, o = [], 0
    for a in arrays:
        out.append(flat[o:o + a.size].reshape(a.shape))
        o += a.size
    return out


def allreduce_max(value):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def ensemble_spectra(spec_sums):
    """Ensemble- and time-mean KEspec / Ensspec over ALL ranks.
    ``spec_sums`` = (kespec_sum, ensspec_sum, count) as returned by ``EnsembleQGModel.spectra_sums()``."""
    ke, en, count = spec_sums
    ke, en, cnt = allreduce_sum([ke, en, np.array([float(count)])])
    n = max(cnt[0], 1.0)
    return ke / n, en / n, int(cnt[0])


def ensemble_diagnostics(diag_sums):
    """Ensemble- and time-mean pyqg diagnostics over ALL ranks (one all-reduce of the concatenated accumulators).
    ``diag_sums`` = (dict name -> sum, count) as returned by ``EnsembleQGModel.diagnostic_sums()``; returns
    (dict name -> mean incl. ``paramspec``, total count)."""
    d, count = diag_sums
    names = sorted(d)
    flat = np.concatenate([np.asarray(d[k], dtype=np.float64).ravel() for k in names] + [np.array([float(count)])])
    red = allreduce_sum([flat])[0]
    n = max(red[-1], 1.0)
    out, o = {}, 0
    for k in names:
        sz = int(np.prod(np.shape(d[k])))
        out[k] = (red[o:o + sz] / n).reshape(np.shape(d[k]))
        o += sz
    if 'paramspec_KEflux' in out:
        out['paramspec'] = out['paramspec_KEflux'] + out['paramspec_APEflux']
    return out, int(red[-1])


def ensemble_diagnostics_device(model, reset=False):
    """Same result as ``ensemble_diagnostics(model.diagnostic_sums())`` without the host bounce: the engine hands out its
    device accumulators (``qgb_diag_averages(on_device=1)``), ONE ``all_reduce`` (NCCL over NVLink) sums them together with
    the sample count, and only the reduced means travel to the host.  Returns (dict name -> mean, total count)."""
    import ctypes
    from . import _lib
    nl, nk = model.nl, model.nk
    nterms = 4 + len(model.DIAG_BUDGET)
    dev = torch.device('cuda', model.device_index)
    buf = torch.zeros(nterms * nl * nk + 1, dtype=torch.float64, device=dev)
    ns = ctypes.c_int64(0)
    _lib.check(model._lib.qgb_diag_averages(model._h, buf.data_ptr(), ctypes.byref(ns), 1 if reset else 0, 1, model._stream()),
               model._h)
    buf[-1] = float(ns.value) * model.members
    if dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == 'nccl':
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        else:                                   # gloo (CPU tests of the orchestration): reduce on the host
            hb = buf.cpu()
            dist.all_reduce(hb, op=dist.ReduceOp.SUM)
            buf = hb
    red = buf.cpu().numpy()
    n = max(red[-1], 1.0)
    a = (red[:-1] / n).reshape((nterms, nl, nk))
    out = {'KEspec': a[0:2].copy(), 'Ensspec': a[2:4].copy()}
    for i, name in enumerate(model.DIAG_BUDGET):
        out[name] = a[4 + i].copy()
    out['paramspec'] = out['paramspec_KEflux'] + out['paramspec_APEflux']
    out.update(model.derived_scalars(out['KEspec']))
    return out, int(red[-1])


def ensemble_ke(ke_members):
    """Ensemble-mean kinetic energy over all ranks from the per-member values of the local shard."""
    s, c = allreduce_sum([np.array([np.nansum(ke_members)]), np.array([float(np.isfinite(ke_members).sum())])])
    return s[0] / max(c[0], 1.0)


def calc_ispec(k, l, spec2d, averaging=True, truncate=True, nd_wavenumber=False, nfactor=1):
    """Isotropic spectrum of a (nl, nk) half-plane density given the wavenumber arrays instead of a model:
    ``tools.spectral_tools.calc_ispec`` (= pyqg_generative/tools/spectral_tools.py:103-180, bit-for-bit) on a ``GridView``.
    Used to turn the all-reduced KEspec into KE(kappa).  Returns (kr, spectrum)."""
    from .tools.spectral_tools import GridView, calc_ispec as _calc_ispec
    return _calc_ispec(GridView(k, l), spec2d, averaging=averaging, truncate=truncate, nd_wavenumber=nd_wavenumber,
                       nfactor=nfactor)
