"""Development aid: what does this B200 sustain for write-only / read-only / copy streams?  (torch kernels; the
roofline denominator stays MEASURED_PEAKS.json's copy figure -- this only tells how much of it a write-only kernel
such as the march can ever see.)"""
import torch
n = 1 << 30   # 4 GiB fp32
x = torch.empty(n, device="cuda"); y = torch.empty(n, device="cuda")
def t(fn, k=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
    e[0].record()
    for i in range(k):
        fn(); e[i + 1].record()
    torch.cuda.synchronize()
    return min(e[i].elapsed_time(e[i + 1]) for i in range(k))
for name, fn, nbytes in (("fill (write-only)", lambda: x.fill_(1.5), 4 * n), ("memset (zero_)", lambda: x.zero_(), 4 * n),
                         ("sum (read-only)", lambda: x.sum(), 4 * n), ("copy (read+write)", lambda: y.copy_(x), 8 * n)):
    ms = t(fn)
    print(f"{name:20s} {ms:8.3f} ms  {nbytes / ms / 1e6:8.0f} GB/s")
