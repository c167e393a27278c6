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


def ensemble_ke(ke_members):
    """Ensemble-mean kinetic energy over all ranks from the per-member values of the local shard."""
    s, c = allreduce_sum([np.array([np.nansum(ke_members)]), np.array([float(np.isfinite(ke_members).sum())])])
    return s[0] / max(c[0], 1.0)


def calc_ispec(k, l, spec2d, averaging=True, truncate=True, nd_wavenumber=False, nfactor=1):
    """Isotropic spectrum of a (nl, nk) half-plane density -- pyqg_generative/tools/spectral_tools.py:103-180
    (``calc_ispec``), used to turn the reduced KEspec into KE(kappa).  Returns (kr, spectrum)."""
    kk, ll = k[0], l[:, 0]
    dk, dl = kk[1] - kk[0], ll[1] - ll[0]
    dkr = nfactor * np.sqrt(dk ** 2 + dl ** 2)
    kmax = min(np.abs(ll).max(), np.abs(kk).max()) if truncate else np.sqrt(np.abs(ll).max() ** 2 + np.abs(kk).max() ** 2)
    kr = np.arange(dkr / 2., kmax + dkr, dkr)
    wv = np.sqrt(k ** 2 + l ** 2)
    spec = np.array(spec2d, dtype=np.float64).copy()
    out = np.zeros(kr.size - 1)
    keep = np.ones(kr.size - 1, dtype=bool)
    for i in range(kr.size - 1):
        mask = (wv >= kr[i]) & (wv < kr[i + 1])
        n = mask.sum()
        if n == 0:
            keep[i] = False
            continue
        if averaging:
            # density in |kappa|: mean over the shell times the shell circumference (half-plane -> factor pi)
            out[i] = spec[mask].mean() * (kr[i] + kr[i + 1]) / 2 * np.pi / (dk * dl)
        else:
            out[i] = spec[mask].sum() / dkr
    krm = (kr[:-1] + kr[1:]) / 2
    return krm[keep], out[keep]
