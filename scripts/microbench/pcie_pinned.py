"""Host <-> device bandwidth from pinned memory: per GPU and with all visible GPUs at once, H2D, D2H and both directions,
with 1 / 2 / 4 transfer streams per GPU.  Explains the end-to-end (host q in / out every step) scaling of bench.py: the
e2e path moves 2 x 67 MB per 1024-member step and GPU.
usage: python scripts/microbench/pcie_pinned.py [MiB per buffer = 64] > gpurun_out/pcie.json"""
import json
import sys
import threading
import time

import torch

MIB = int(sys.argv[1]) if len(sys.argv) > 1 else 64
REPS = 20


def run(devs, mode, streams_per_gpu):
    bufs = []
    for d in devs:
        for s in range(streams_per_gpu):
            h_in = torch.empty(MIB << 20, dtype=torch.uint8).pin_memory()
            h_out = torch.empty(MIB << 20, dtype=torch.uint8).pin_memory()
            g_in = torch.empty(MIB << 20, dtype=torch.uint8, device='cuda:%d' % d)
            g_out = torch.zeros(MIB << 20, dtype=torch.uint8, device='cuda:%d' % d)
            bufs.append((d, torch.cuda.Stream(device=d), h_in, h_out, g_in, g_out))

    def enqueue(reps):
        for _ in range(reps):
            for d, st, h_in, h_out, g_in, g_out in bufs:
                with torch.cuda.stream(st):
                    if mode in ('h2d', 'bidir'):
                        g_in.copy_(h_in, non_blocking=True)
                    if mode in ('d2h', 'bidir'):
                        h_out.copy_(g_out, non_blocking=True)
    enqueue(2)
    for d in devs:
        torch.cuda.synchronize(d)
    t0 = time.perf_counter()
    enqueue(REPS)
    for d in devs:
        torch.cuda.synchronize(d)
    dt = time.perf_counter() - t0
    nbytes = REPS * len(bufs) * (MIB << 20) * (2 if mode == 'bidir' else 1)
    return nbytes / dt / 1e9


def main():
    n = torch.cuda.device_count()
    out = {'gpus': n, 'mib_per_buffer': MIB, 'reps': REPS, 'rows': []}
    try:
        with open('/proc/cpuinfo') as f:
            out['host_cpus'] = sum(1 for line in f if line.startswith('processor'))
    except Exception:
        pass
    for devs in ([0], list(range(n))) if n > 1 else ([0],):
        for mode in ('h2d', 'd2h', 'bidir'):
            for spg in (1, 2, 4):
                gbs = run(devs, mode, spg)
                out['rows'].append({'gpus_active': len(devs), 'mode': mode, 'streams_per_gpu': spg, 'total_GBps': round(gbs, 2),
                                    'per_gpu_GBps': round(gbs / len(devs), 2)})
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
