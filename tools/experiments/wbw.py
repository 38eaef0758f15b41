import torch, time
n = 604078084
x = torch.empty(n, dtype=torch.float64, device='cuda')
y = torch.empty(n, dtype=torch.float64, device='cuda')
def t(label, f, bytes_, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print('%-28s %.3f ms  %.0f GB/s' % (label, ms, bytes_ / ms / 1e6))
t('fill_ (write 4.83 GB)', lambda: x.fill_(1.0), n * 8)
t('zero_ (memset)', lambda: x.zero_(), n * 8)
t('copy_ (read+write 9.66 GB)', lambda: y.copy_(x), 2 * n * 8)
t('sum (read 4.83 GB)', lambda: x.sum(), n * 8)
t('x.add_(1) (rw 9.66GB)', lambda: x.add_(1.0), 2 * n * 8)
