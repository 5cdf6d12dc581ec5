"""GPU session helper: pure-write, pure-read and copy bandwidth of this B200 with library kernels (torch fill_ / sum /
copy_ over 4 GiB), as reference points for store-bound kernels (the driver's MEASURED_PEAKS.json has the copy figure)."""
import torch

n = 1 << 30
x = torch.empty(n, dtype=torch.float32, device='cuda')
y = torch.empty(n, dtype=torch.float32, device='cuda')


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


ms = timed(lambda: x.fill_(1.0))
print(f'fill_ (pure write) : {4 * n / ms / 1e6:.0f} GB/s')
ms = timed(lambda: x.sum())
print(f'sum   (pure read)  : {4 * n / ms / 1e6:.0f} GB/s')
ms = timed(lambda: y.copy_(x))
print(f'copy_ (read+write) : {8 * n / ms / 1e6:.0f} GB/s')
