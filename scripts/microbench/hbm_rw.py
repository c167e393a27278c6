"""HBM bandwidth on this B200 by direction: write-only (fill), read-only (sum), copy (read + write)."""
import torch
n = 1 << 30   # 4 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device='cuda')
b = torch.empty(n, dtype=torch.float32, device='cuda')
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e-3
tw = t(lambda: a.fill_(1.0)); print('write-only (fill_)  : %.2f TB/s' % (4 * n / tw / 1e12))
tm = t(lambda: torch.cuda.memset if False else a.zero_()); print('write-only (zero_)  : %.2f TB/s' % (4 * n / tm / 1e12))
tr = t(lambda: a.sum()); print('read-only (sum)     : %.2f TB/s' % (4 * n / tr / 1e12))
tc = t(lambda: b.copy_(a)); print('copy (read + write) : %.2f TB/s total' % (8 * n / tc / 1e12))
