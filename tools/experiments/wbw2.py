"""Pure-write bandwidth of the device for the assembly's output size (4.8 GB of fp64)."""
import torch
n = 604078084
v = torch.empty(n, dtype=torch.float64, device="cuda")
def t(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ms = t(lambda: v.fill_(1.0)); print("fill_  %.3f ms  %.0f GB/s" % (ms, n * 8 / ms / 1e6))
ms = t(lambda: v.zero_());    print("zero_  %.3f ms  %.0f GB/s" % (ms, n * 8 / ms / 1e6))
w = torch.empty_like(v)
ms = t(lambda: w.copy_(v));   print("copy_  %.3f ms  %.0f GB/s (read+write)" % (ms, 2 * n * 8 / ms / 1e6))
