"""write / read / copy bandwidth probes (context for the store cost of the dominant kernel)."""
import torch
n = 300_000_000
a = torch.empty(n, dtype=torch.float64, device='cuda')
b = torch.empty(n, dtype=torch.float64, device='cuda')
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: a.fill_(1.0)); print('fill  2.4 GB: %.3f ms  %.0f GB/s written' % (ms, 8 * n / ms / 1e6))
ms = t(lambda: a.sum());      print('sum   2.4 GB: %.3f ms  %.0f GB/s read' % (ms, 8 * n / ms / 1e6))
ms = t(lambda: b.copy_(a));   print('copy  2.4 GB: %.3f ms  %.0f GB/s read+write' % (ms, 16 * n / ms / 1e6))
# instance-major scatter pattern: 4096 rows of 73242 doubles, written as [32 rows][256 B segments] (like the flush)
v = a[:4096 * 73242].view(4096, 73242)
src = torch.randn(32, 73242, dtype=torch.float64, device='cuda')
ms = t(lambda: [v[i:i + 32].copy_(src) for i in range(0, 4096, 32)]); print('row-block copy 2.4 GB: %.3f ms  %.0f GB/s written' % (ms, 8 * 4096 * 73242 / ms / 1e6))
