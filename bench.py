#!/usr/bin/env python
"""bench.py -- ensemble-member-steps/s of the online-simulation hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (this repository's engine)
    python bench.py --impl reference --gpus N ...            # reference CPU implementation (oracle port) on host cores

Workload (BASELINE.json configs[2]): nx=64 eddy configuration, dt=14400 s, CGAN closure (AndrewCNN 4->2), 1024 ensemble
members PER GPU (weak scaling: members are independent, no data-path collective), synthetic developed-turbulence-like
initial states and random-init generator weights of the named shapes, shipped x/y scalers, white latent noise each
step (sampling 'constant', nsteps=1 as in scripts/run_parameterized.py:50).  A "step" is one model time step of every
member: spectral step kernel + noise + 8 convolution layers + denormalisation.

Timed region: W untimed warm-up steps (W >= 3 also completes the Adams-Bashforth start-up), then exactly K steps
bracketed by barrier + cudaDeviceSynchronize, CUDA events on the launching stream, max over ranks.  State + activations
(> 5 GB) exceed the 126 MB L2, so no explicit flush is needed (config.l2: "inputs larger than L2").
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'ensemble_member_steps_per_s'
UNIT = 'member-steps/s'
NX, DT = 64, 14400.0
X_STD = [7.784383342368528e-06, 1.0471941322975908e-06]      # Google-Colab/GAN/x_scale.json
Y_STD = [7.60611105349307e-12, 1.656513061486578e-13]         # Google-Colab/GAN/y_scale.json
MAC_PER_PIXEL = [12800, 204800, 18432, 9216, 9216, 9216, 9216, 576]   # SURVEY.md Appendix B (GAN/VAE generator)


def synthetic_states(members, n, seed):
    """Developed-turbulence-like states: Gaussian random fields with the shipped x_scale stds and a q-amplitude
    spectrum ~ kappa^0.5 truncated at 0.65*pi/dx, which gives KE ~ 5e-4 and CFL ~ 0.2 at dt=14400 -- the saturated
    values recorded in notebooks/3-2-dealiasing.ipynb:1434-1440."""
    rng = np.random.RandomState(seed)
    dk = 2 * np.pi / 1e6
    ll = dk * np.append(np.arange(0., n / 2), np.arange(-n / 2, 0.))
    kk = dk * np.arange(0., n // 2 + 1)
    k, l = np.meshgrid(kk, ll)
    wv = np.sqrt(k ** 2 + l ** 2)
    amp = np.sqrt(wv / dk) * (wv * (1e6 / n) <= 0.65 * np.pi)
    out = np.empty((members, 2, n, n))
    for z, std in enumerate(X_STD):
        h = np.fft.rfftn(rng.randn(members, n, n), axes=(-2, -1)) * amp
        f = np.fft.irfftn(h, s=(n, n), axes=(-2, -1))
        out[:, z] = f / f.std(axis=(-2, -1), keepdims=True) * std
    return out


def clock_sampler(stop, samples, device):
    """Sample SM clock, power and clock-event (throttle) reasons DURING the timed region: NVML every 20 ms, falling back to
    the nvidia-smi query of /opt/skills/guides/B200_PROFILING.md."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = None
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(device).uuid)
        except Exception:
            pass
        h = None
        if uuid:
            for cand in ('GPU-' + uuid, uuid):
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                    break
                except Exception:
                    h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(device)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        get_reasons = getattr(pynvml, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
            getattr(pynvml, 'nvmlDeviceGetCurrentClocksThrottleReasons')
        bits = {'hw_slowdown': 0x8, 'sw_power_cap': 0x4, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20}
        while not stop.is_set():
            r = int(get_reasons(h))
            samples.append({'sm': float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), 'max': float(mx),
                            'power': pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                            'reasons': [k for k, b in bits.items() if r & b]})
            stop.wait(0.02)
        return
    except Exception:
        pass
    q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    while not stop.is_set():
        try:
            out = subprocess.run(['nvidia-smi', '-i', str(device), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                 capture_output=True, text=True, timeout=5).stdout.strip()
            f = [x.strip() for x in out.split(',')]
            samples.append({'sm': float(f[0]), 'max': float(f[1]), 'power': float(f[2]),
                            'reasons': [n for i, n in enumerate(names) if f[3 + i].lower().startswith('active')]})
        except Exception:
            pass
        stop.wait(0.1)


def summarize_clocks(samples):
    if not samples:
        return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no clock samples (NVML and nvidia-smi unavailable)']}
    sm = sorted(s['sm'] for s in samples)
    reasons = sorted({r for s in samples for r in s['reasons']})
    return {'sm_mhz': sm[len(sm) // 2], 'sm_min_mhz': sm[0], 'sm_max_mhz': samples[0]['max'],
            'power_w_max': max(s['power'] for s in samples), 'samples': len(samples), 'reasons': reasons}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return json.load(f), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


# ------------------------------------------------------------------------------------------------------------------
# reference CPU implementation of the path (oracle port): pyqg shim step + AndrewCNN on CPU torch, all host threads
# ------------------------------------------------------------------------------------------------------------------
class CpuEnsemble(object):
    def __init__(self, members, sd, seed=0):
        import torch
        from oracle import cnn_ref, pyqg_shim
        self.torch, self.cnn_ref = torch, cnn_ref
        self.sd = sd
        q0 = synthetic_states(members, NX, seed)
        self.models = []
        for b in range(members):
            m = pyqg_shim.QGModel(nx=NX, dt=DT, log_level=0, parameterization=_Slot())
            m.q = q0[b]
            self.models.append(m)
        self.rng = np.random.RandomState(seed + 1)

    def step(self):
        q = np.stack([m.q for m in self.models])
        z = self.rng.randn(len(self.models), 2, NX, NX).astype('float32')        # constant sampler, nsteps=1
        dq = self.cnn_ref.predict_snapshot('gan', [self.sd], X_STD, Y_STD, q, z)   # apply_function + generate
        dq = self.cnn_ref.demean(dq)
        for b, m in enumerate(self.models):
            m.q_parameterization.dq = dq[b]
            m._step_forward()


class _Slot(object):
    parameterization_type = 'q_parameterization'
    dq = None

    def __call__(self, m):
        return self.dq


def cpu_baseline(sample_members, steps, warmup, sd):
    import torch
    # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses every host core it can see regardless of the launcher
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, RuntimeError):
        pass
    ens = CpuEnsemble(sample_members, sd)
    for _ in range(warmup):
        ens.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        ens.step()
    dt = time.perf_counter() - t0
    return sample_members * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return 0
    from oracle import cnn_ref
    sd = cnn_ref.random_state_dict(4, 2, seed=0)
    sample = args.ref_members
    value, ms, cores = cpu_baseline(sample, args.steps, args.warmup, sd)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64 spectral step + f32 CNN', 'data': 'synthetic',
        'config': {'workload': 'nx=64 eddy + CGAN closure (configs[2]); each step = %d members (bounded sample of the '
                               '1024-member ensemble)' % sample, 'nx': NX, 'dt': DT, 'closure': 'gan'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d members x %d steps, oracle/pyqg_shim.py + oracle/cnn_ref.py (CPU torch, %d threads); '
                                   'pyqg itself is not installable here' % (sample, args.steps, cores)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------------------
# library baseline on the same B200: torch eager (cuFFT fp64 + cuDNN convolutions), batched over the members -- "the
# number the hand-written kernels must beat" (reference call sites
# tools/cnn_tools.py:85-98,163-176 for the network, pyqg kernel.pyx for the step)
# ------------------------------------------------------------------------------------------------------------------
def library_baseline(members, steps, sd, q0, device):
    import torch
    import torch.nn.functional as F
    dev = torch.device('cuda', device)
    N, B = NX, members
    dk = 2 * np.pi / 1e6
    ll = dk * np.append(np.arange(0., N / 2), np.arange(-N / 2, 0.))
    kk = dk * np.arange(0., N // 2 + 1)
    k, l = np.meshgrid(kk, ll)
    wv2 = k ** 2 + l ** 2
    rd, delta, beta, rek, U1, U2 = 15000.0, 0.25, 1.5e-11, 5.787e-7, 0.025, 0.0
    F1 = rd ** -2 / (1 + delta)
    F2 = delta * F1
    det = wv2 * (wv2 + F1 + F2)
    with np.errstate(divide='ignore', invalid='ignore'):
        a = np.stack([-(wv2 + F2) / det, -F1 / det, -F2 / det, -(wv2 + F1) / det])
    a[:, 0, 0] = 0.0
    wvx = np.sqrt((k * 1e6 / N) ** 2 + (l * 1e6 / N) ** 2)
    filtr = np.where(wvx <= 0.65 * np.pi, 1.0, np.exp(-23.6 * (wvx - 0.65 * np.pi) ** 4))
    t = lambda x, dt=torch.float64: torch.as_tensor(np.asarray(x), dtype=dt, device=dev)
    a_t, ik, il, filt_t, wv2_t = t(a), 1j * t(k), 1j * t(l), t(filtr), t(wv2)
    Ubg = t([U1, U2]).reshape(1, 2, 1, 1)
    ikQy = ik[None, None] * t([beta + F1 * (U1 - U2), beta - F2 * (U1 - U2)]).reshape(1, 2, 1, 1)
    xs, ys = t(X_STD, torch.float32).reshape(1, 2, 1, 1), t(Y_STD, torch.float32).reshape(1, 2, 1, 1)
    w = {kk_: v.to(dev) for kk_, v in sd.items()}

    def cnn(x):
        for n in range(8):
            wt = w['conv.%d.weight' % (3 * n)]
            p = wt.shape[-1] // 2
            x = F.conv2d(F.pad(x, (p, p, p, p), mode='circular'), wt, w['conv.%d.bias' % (3 * n)])
            if n < 7:
                j = 3 * n + 2
                x = F.batch_norm(F.relu(x), w['conv.%d.running_mean' % j], w['conv.%d.running_var' % j],
                                 w['conv.%d.weight' % j], w['conv.%d.bias' % j], False, 0.0, 1e-5)
        return x

    q = t(q0[:B])
    qh = torch.fft.rfft2(q)
    hist = [torch.zeros_like(qh), torch.zeros_like(qh)]
    dtc = (23. / 12 * DT, -16. / 12 * DT, 5. / 12 * DT)

    def closure(q):
        x = torch.cat([q.float() / xs, torch.randn(B, 2, N, N, device=dev)], dim=1)
        y = (cnn(x) * ys).double()
        return y - y.mean(dim=(-2, -1), keepdim=True)

    def spectral(q, qh, dq):
        ph = torch.stack([a_t[0] * qh[:, 0] + a_t[1] * qh[:, 1], a_t[2] * qh[:, 0] + a_t[3] * qh[:, 1]], dim=1)
        u = torch.fft.irfft2(-il * ph, s=(N, N))
        v = torch.fft.irfft2(ik * ph, s=(N, N))
        d = -(ik * torch.fft.rfft2((u + Ubg) * q) + il * torch.fft.rfft2(v * q) + ikQy * ph)
        d[:, 1] += rek * wv2_t * ph[:, 1]
        d = d + torch.fft.rfft2(dq)
        qh = filt_t * (qh + dtc[0] * d + dtc[1] * hist[0] + dtc[2] * hist[1])
        hist[1], hist[0] = hist[0], d
        return torch.fft.irfft2(qh, s=(N, N)), qh

    out = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for name, tf32 in (('tf32', True), ('fp32', False)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        t_c = t_s = 0.0
        for it in range(steps + 2):
            e0, e1, e2 = ev(), ev(), ev()
            e0.record()
            dq = closure(q)
            e1.record()
            q, qh = spectral(q, qh, dq)
            e2.record()
            torch.cuda.synchronize(dev)
            if it >= 2:
                t_c += e0.elapsed_time(e1)
                t_s += e1.elapsed_time(e2)
        out[name] = {'value': B * steps / ((t_c + t_s) * 1e-3), 'cnn_ms_per_step': t_c / steps, 'spectral_ms_per_step': t_s / steps}
    out['healthy'] = bool(torch.isfinite(q).all().item())
    torch.backends.cudnn.allow_tf32 = True
    del q, qh, hist, w
    torch.cuda.empty_cache()
    return out


# activation bytes per pixel that a conv layer reads + writes in the format the tensor-core path stores them (fp16 hi plane +
# e4m3 lo plane = 3 B per element; the raw fp32 network input; the fp32 output); halos and weights excluded
def conv_algorithmic_bytes(li, fast):
    cin = [4, 128, 64, 32, 32, 32, 32, 32][li]
    cout = [128, 64, 32, 32, 32, 32, 32, 2][li]
    bin_ = 4.0 * cin if li == 0 else (2.0 if (fast and li == 1) else 3.0) * cin
    bout = 4.0 * cout if li == 7 else (2.0 if (fast and li == 0) else 3.0) * cout
    return (bin_ + bout) * NX * NX


def make_model(count, offset, local, precision, sd, **extra):
    from pyqg_generative_b200.models.cgan_regression import CGANRegression
    from pyqg_generative_b200.tools.cnn_tools import ChannelwiseScaler
    from pyqg_generative_b200.tools.stochastic_pyqg import stochastic_QGModel
    gan = CGANRegression(folder='/nonexistent', nx=NX, precision=precision)
    gan.G.load_state_dict(sd)
    gan.x_scale, gan.y_scale = ChannelwiseScaler(), ChannelwiseScaler()
    gan.x_scale.std = np.array(X_STD, 'float32').reshape(1, 2, 1, 1)
    gan.y_scale.std = np.array(Y_STD, 'float32').reshape(1, 2, 1, 1)
    params = dict(nx=NX, dt=DT, log_level=0, tmax=1e12, tavestart=1e12, members=count, member_offset=offset,
                  device=local, parameterization=gan, precision=precision, seed=2024)
    params.update(extra)
    return stochastic_QGModel(params, 'constant', 1), params


def run_b200(args):
    import torch
    from pyqg_generative_b200 import _lib, build, parallel
    build.build()
    rank, world, local = parallel.init_from_env()
    torch.cuda.set_device(local)
    from oracle import cnn_ref                      # only for the synthetic random-init weights and cpu_baseline

    B = args.members
    count, offset = B, rank * B                     # weak scaling: every GPU integrates ``members`` members
    sd = cnn_ref.random_state_dict(4, 2, seed=0)
    m, params = make_model(count, offset, local, args.precision, sd)
    q0 = synthetic_states(count, NX, 1234 + rank)
    m.set_q(q0)
    lib, h, stream = m._lib, m._h, m._stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()

    def timed(model, nsteps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        _lib.check(model._lib.qgb_step(model._h, nsteps, model._stream()), model._h)
        e1.record()
        barrier()
        return parallel.allreduce_max(e0.elapsed_time(e1))

    _lib.check(lib.qgb_step(h, max(args.warmup, 3), stream), h)
    barrier()
    chosen, calib = m.closure_precision()           # 'auto' is calibrated by the first closure evaluation (inside the warm-up)
    fast = chosen == 'tc_fast'
    # ---- device-resident timed region -------------------------------------------------------------------------
    samples, stop = [], threading.Event()
    th = threading.Thread(target=clock_sampler, args=(stop, samples, local), daemon=True)
    th.start()
    _lib.check(lib.qgb_profile_begin(h, 0, 1), h)              # layer 2 (128->64, 5x5): 75 % of the FLOPs
    l0 = _lib.launch_count()
    ms = timed(m, args.steps)
    launches = _lib.launch_count() - l0
    pms, pl, pim = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
    _lib.check(lib.qgb_profile_end(h, ctypes.byref(pms), ctypes.byref(pl), ctypes.byref(pim)), h)
    stop.set()
    th.join(timeout=2)
    value = world * count * args.steps / (ms * 1e-3)
    ke, cfl, flags = m.diagnostics()
    healthy = bool(np.isfinite(ke).all() and not flags.any())

    # ---- per-kernel table: every kernel of the step bracketed by CUDA events (separate short pass) ------------------------
    peaks, which = measured_peaks()
    # denominator: the burst figure when the timed region ran at (nearly) full SM clock, the sustained one when it sat under the power cap
    clk = summarize_clocks(samples)
    at_full_clock = bool(clk.get('sm_mhz') and clk.get('sm_max_mhz') and clk['sm_mhz'] >= 0.93 * clk['sm_max_mhz'])
    peak_key = 'bf16_tflops' if at_full_clock else 'bf16_tflops_sustained'
    peak = peaks[peak_key]
    ksteps = max(2, min(args.steps, 10))
    _lib.check(lib.qgb_profile_all_begin(h), h)
    _lib.check(lib.qgb_step(h, ksteps, stream), h)
    kms = (ctypes.c_double * _lib.PROF_SLOTS)()
    kl = (ctypes.c_int64 * _lib.PROF_SLOTS)()
    ku = (ctypes.c_int64 * _lib.PROF_SLOTS)()
    _lib.check(lib.qgb_profile_all_end(h, kms, kl, ku), h)
    kernels = []
    for slot in range(_lib.PROF_SLOTS):
        if not kl[slot]:
            continue
        ms_launch = kms[slot] / kl[slot]
        per = ku[slot] / kl[slot]                    # images / members per launch
        rec = {'kernel': _lib.PROF_SLOT_NAMES[slot], 'ms_per_launch': ms_launch, 'launches_per_step': kl[slot] / ksteps,
               'units_per_launch': per}
        if slot < 8:
            flops = 2.0 * MAC_PER_PIXEL[slot] * NX * NX * per
            byts = conv_algorithmic_bytes(slot, fast) * per
            rec.update(algorithmic_flops=flops, algorithmic_bytes=byts,
                       tensor_frac=flops / (ms_launch * 1e-3) / 1e12 / peak,
                       hbm_frac=byts / (ms_launch * 1e-3) / 1e9 / peaks['hbm_gbs'])
        else:
            per_unit = {16: 5 * (2 * NX * (NX // 2 + 1) * 16) + 2 * (2 * NX * NX * 8),     # 5 S_c + 2 S_r = 468 992 B
                        17: 2 * NX * NX * 4, 18: 2 * NX * NX * (4 + 8), 19: 0}[slot]
            rec.update(algorithmic_bytes=per_unit * per, hbm_frac=per_unit * per / (ms_launch * 1e-3) / 1e9 / peaks['hbm_gbs'])
        rec['binding_frac'] = max(rec.get('tensor_frac', 0.0), rec.get('hbm_frac', 0.0))
        kernels.append(rec)
    kernel_ms_per_step = sum(kms[s] for s in range(_lib.PROF_SLOTS)) / ksteps

    # ---- multi-GPU only: the one collective of the path -- all-reduce of the online spectral diagnostics (NCCL), timed, and
    # checked against a rank-ordered recomputation from the gathered per-rank accumulators (sharding invariance)
    diag_obj = None
    if world > 1:
        _lib.check(lib.qgb_diag_config(h, 0.0, 4 * DT), h)
        _lib.check(lib.qgb_step(h, 12, stream), h)
        parallel.ensemble_diagnostics_device(m)                  # warm-up (NCCL communicator set-up)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        red, cnt = parallel.ensemble_diagnostics_device(m)
        e1.record()
        torch.cuda.synchronize()
        d, c = m.diagnostic_sums()
        names = sorted(d)
        flat = torch.as_tensor(np.concatenate([d[k_].ravel() for k_ in names] + [np.array([float(c)])]), device='cuda')
        gathered = [torch.empty_like(flat) for _ in range(world)]
        torch.distributed.all_gather(gathered, flat)
        tot = torch.stack(gathered).cpu().numpy().sum(axis=0)
        dev_, o = 0.0, 0
        for k_ in names:
            sz = d[k_].size
            ref = tot[o:o + sz] / tot[-1]
            dev_ = max(dev_, float(np.abs(red[k_].ravel() - ref).max() / max(np.abs(ref).max(), 1e-300)))
            o += sz
        diag_obj = {'collective': 'all_reduce(sum, f64) of KEspec, Ensspec and the %d spectral budget terms + sample count (NCCL)' % len(m.DIAG_BUDGET),
                    'bytes': int(flat.numel() * 8), 'ms': parallel.allreduce_max(e0.elapsed_time(e1)), 'samples_x_members': int(cnt),
                    'max_rel_dev_vs_rank_ordered_gather': dev_, 'sharding_invariant': bool(dev_ < 1e-12 and cnt == tot[-1])}
        _lib.check(lib.qgb_diag_config(h, 1e12, 86400.0), h)

    # ---- informational: what pyqg's time-averaged diagnostics add when they are on (the timed region runs with them off, like a
    # run before ``tavestart``).  One sample = KEspec, Ensspec and the 13 budget terms of all local members, accumulated on the device
    # every ``taveint`` = 1 day = 6 steps of this configuration (profiles/r2_diag_overhead.md).
    diag_sample = None
    if world == 1:
        try:
            _lib.check(lib.qgb_diag_config(h, 0.0, DT), h)               # a sample before every step
            _lib.check(lib.qgb_step(h, 3, stream), h)
            t_on = timed(m, 12) / 12
            _lib.check(lib.qgb_diag_config(h, 1e12, 86400.0), h)
            _lib.check(lib.qgb_step(h, 3, stream), h)
            t_off = timed(m, 12) / 12
            diag_sample = {'ms_per_sample': t_on - t_off, 'members': count, 'steps_per_sample_in_the_reference_runs': int(round(86400.0 / DT)),
                           'ms_per_step_amortised': (t_on - t_off) / round(86400.0 / DT),
                           'what': 'KEspec, Ensspec and the %d spectral budget terms of every member (PROG_BUDGET on the register-FFT '
                                   'kernel) + the reduction over the members; step time with a sample before every step minus without' % len(m.DIAG_BUDGET)}
        except Exception as e:                                   # informational: must never take the product line down
            diag_sample = {'error': '%s: %s' % (type(e).__name__, e)}
            try:
                _lib.check(lib.qgb_diag_config(h, 1e12, 86400.0), h)
            except Exception:
                pass

    # ---- strong scaling of configs[2]: 1024 members IN TOTAL over the N ranks ------------------------------------------------
    strong = {'members_total': B, 'members_per_gpu': count, 'value': value, 'ms_per_step': ms / args.steps}
    if world > 1:
        sc, so = parallel.shard_members(B, rank, world)
        ms_, _ = make_model(sc, so, local, args.precision, sd)
        ms_.set_q(synthetic_states(B, NX, 99)[so:so + sc])
        _lib.check(ms_._lib.qgb_step(ms_._h, max(args.warmup, 3), ms_._stream()), ms_._h)
        t_s = timed(ms_, args.steps)
        strong = {'members_total': B, 'members_per_gpu': sc, 'value': B * args.steps / (t_s * 1e-3), 'ms_per_step': t_s / args.steps}
        del ms_

    # ---- end-to-end through the host-buffer API: every step uploads q from pinned host memory, advances one step and
    # downloads q.  The members of this GPU are driven as G groups on G streams (public API: step_host(wait=False)), so
    # the PCIe transfers of one group overlap the kernels of the others; a step of a group still waits for its own q.
    del m
    torch.cuda.empty_cache()
    G = args.e2e_groups
    gm = count // G
    groups = []
    for g in range(G):
        mg, _ = make_model(gm, offset + g * gm, local, chosen if args.precision == 'auto' else args.precision, sd)
        qin = torch.from_numpy(q0[g * gm:(g + 1) * gm].copy()).pin_memory()
        qout = torch.empty_like(qin).pin_memory()
        groups.append([mg, qin, qout, torch.cuda.Stream(device=local)])
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def e2e_round(n):
        for _ in range(n):
            for grp in groups:
                grp[0].step_host(grp[1], grp[2], 1, stream=grp[3], wait=False)
                grp[1], grp[2] = grp[2], grp[1]
        for grp in groups:
            grp[3].synchronize()
    e2e_round(3)
    barrier()
    t0 = time.perf_counter()
    e2e_round(e2e_steps)
    barrier()
    e2e_s = parallel.allreduce_max(time.perf_counter() - t0)
    e2e_value = world * G * gm * e2e_steps / e2e_s
    nbytes = int(G * gm * 2 * NX * NX * 8)
    e2e_healthy = all(bool(np.isfinite(grp[0].diagnostics()[0]).all()) for grp in groups)
    # same loop with the result read back as the float32 snapshot the reference stores (drop_vars, tools/simulate.py:16-36):
    # float64 q in, one step, float32 q out converted on the device -- 3/4 of the PCIe bytes (the 8-GPU line is bound by the
    # host's D2H fabric, profiles/r2_pcie_8gpu.json)
    for grp in groups:
        grp.append(torch.empty(grp[1].shape, dtype=torch.float32).pin_memory())

    def e2e32_round(n):
        for _ in range(n):
            for grp in groups:
                grp[0].step_host(grp[1], None, 1, stream=grp[3], wait=False)
                grp[0].real32('q', out=grp[4], stream=grp[3], wait=False)
        for grp in groups:
            grp[3].synchronize()
    e2e32_round(2)
    barrier()
    t0 = time.perf_counter()
    e2e32_round(e2e_steps)
    barrier()
    e2e32_s = parallel.allreduce_max(time.perf_counter() - t0)
    e2e32_value = world * G * gm * e2e_steps / e2e32_s
    del groups
    torch.cuda.empty_cache()

    lib_base = None
    if world == 1 and not args.no_library_baseline:
        try:
            lib_base = library_baseline(count, 3, sd, q0, local)
        except Exception as e:                                   # the baseline must never take the product line down
            lib_base = {'error': '%s: %s' % (type(e).__name__, e)}

    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return 0
    flops_per_image = 2.0 * MAC_PER_PIXEL[1] * NX * NX
    achieved = (flops_per_image * pim.value / max(pl.value, 1)) / (pms.value / max(pl.value, 1) * 1e-3) / 1e12 if pl.value else 0.0
    if world == 1:
        cpu_v, cpu_ms, cores = cpu_baseline(args.ref_members, args.cpu_steps, 1, sd)
        cpu_obj = {'value': cpu_v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                   'sample': '%d members x %d steps of the same workload, oracle/pyqg_shim.py + oracle/cnn_ref.py '
                             '(CPU torch, %d threads)' % (args.ref_members, args.cpu_steps, cores)}
    else:       # the CPU port is timed beside the 1-GPU line only (the other ranks would share its cores here)
        cpu_obj = {'value': None, 'unit': UNIT, 'cores': 0, 'kind': 'port', 'sample': 'not measured at n_gpus > 1; see the n_gpus = 1 line'}
    traffic = None
    traffic_src = 'no ncu capture of this configuration committed'
    try:                                 # dram bytes per 1024-image launch from the committed ncu --set full capture of this kernel
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            tj = json.load(f)
        ent = tj['layer2'][chosen]
        traffic = ent['dram_bytes_per_1024_images'] * (pim.value / max(pl.value, 1)) / 1024.0
        traffic_src = ent['source']
    except Exception:
        pass
    prec_names = {'tc': 'f16 split-precision tcgen05 (f32 accumulate)', 'tc_fast': 'f16 split-precision tcgen05, 1-pass layer 2 (f32 accumulate)',
                  'fp32': 'f32 FFMA'}
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64 spectral step + %s CNN' % prec_names.get(chosen, chosen),
        'data': 'synthetic',
        'config': {'workload': 'nx=64 eddy + CGAN closure, %d members per GPU (configs[2])' % count, 'nx': NX, 'dt': DT,
                   'members_per_gpu': count, 'closure': 'gan', 'sampling': 'constant/1', 'precision': chosen,
                   'precision_requested': args.precision,
                   'precision_calibration': dict(calib, tolerance='rel-L2 <= 1e-3 of the fp32 path (north_star); tc_fast taken when <= 7e-4'),
                   'l2': 'inputs larger than L2 (state + activations > 5 GB)', 'state_healthy': healthy},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': nbytes, 'd2h_bytes_per_step': nbytes,
                'steps': e2e_steps, 'groups': G, 'state_healthy': e2e_healthy,
                'call': 'EnsembleQGModel.step_host -> qgb_step_host_async: pinned host q in, 1 step, host q out, per group of '
                        '%d members on its own stream' % gm},
        'e2e_f32_snapshot': {'value': e2e32_value, 'unit': UNIT, 'h2d_bytes_per_step': nbytes, 'd2h_bytes_per_step': nbytes // 2,
                             'call': 'step_host(q_in float64, no q out) + real32(q): the float32 snapshot the reference stores, '
                                     'converted on the device'},
        'gpu_launches': int(launches),
        'clocks': clk,
        'roofline': {'bound': 'tensor', 'kernel': 'conv layer 2 (128->64, 5x5), %s' % chosen,
                     'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
                     'peak_source': '%s %s (SM clock during the timed region: %s MHz)' % (which, peak_key, clk.get('sm_mhz')),
                     'traffic': traffic, 'traffic_source': traffic_src,
                     'issued_frac': (achieved * (1.5 if chosen == 'tc' else 1.0)) / peak,
                     'issued_note': 'tc runs layer 2 as a_hi x w_hi (fp16) + a_lo x w (e4m3, half cost): issued MMA work is 1.5x the '
                                    'algorithmic flops in bf16-equivalents; tc_fast issues the hi pass only',
                     'launch_ms': pms.value / max(pl.value, 1), 'launches': int(pl.value),
                     'share_of_step': pms.value / ms if ms else None},
        'roofline_step': {'bound': 'tensor', 'algorithmic_flops_per_member_step': 2.0 * sum(MAC_PER_PIXEL) * NX * NX,
                          'achieved': 2.0 * sum(MAC_PER_PIXEL) * NX * NX * (value / world) / 1e12, 'peak': peak, 'unit': 'TFLOP/s',
                          'frac': 2.0 * sum(MAC_PER_PIXEL) * NX * NX * (value / world) / 1e12 / peak},
        'kernels': kernels,
        'kernels_note': 'CUDA events around every launch in a separate %d-step pass (sum %.3f ms/step vs %.3f ms/step in the timed '
                        'region); fractions of %s peaks: %s / hbm_gbs' % (ksteps, kernel_ms_per_step, ms / args.steps, which, peak_key),
        'strong': strong,
        'cpu_baseline': cpu_obj,
    }
    if diag_obj is not None:
        line['diag_allreduce'] = diag_obj
    if diag_sample is not None:
        line['diag_sample'] = diag_sample
    if lib_base is not None:
        line['library_baseline'] = dict(lib_base, what='torch eager on the same GPU and workload: torch.fft.rfft2/irfft2 (cuFFT, fp64) spectral step + '
                                        'F.pad(circular) + conv2d (cuDNN) + batch_norm AndrewCNN, %d members, 3 steps' % count, unit=UNIT)
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', type=str, default='b200')
    ap.add_argument('--members', type=int, default=1024)
    ap.add_argument('--precision', type=str, default=os.environ.get('QGB_PRECISION', 'auto'),
                    help="auto (default: tc_fast if the loaded network measures <= 7e-4 against fp32, else tc), tc, tc_fast, fp32")
    ap.add_argument('--ref-members', type=int, default=16)
    ap.add_argument('--cpu-steps', type=int, default=8)
    ap.add_argument('--e2e-steps', type=int, default=10)
    ap.add_argument('--e2e-groups', type=int, default=4)
    ap.add_argument('--no-library-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    return run_b200(args)


if __name__ == '__main__':
    sys.exit(main())
