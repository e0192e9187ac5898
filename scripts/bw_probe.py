import torch, time
x = torch.empty(131072000, dtype=torch.float32, device='cuda')
y = torch.empty_like(x)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x.fill_(1.0)); print("fill 524MB: %.3f ms  %.0f GB/s" % (ms, 0.524288 / ms * 1e3))
ms = t(lambda: y.copy_(x)); print("copy 524MB: %.3f ms  %.0f GB/s (r+w)" % (ms, 2 * 0.524288 / ms * 1e3))
ms = t(lambda: x.sum()); print("read 524MB: %.3f ms  %.0f GB/s" % (ms, 0.524288 / ms * 1e3))
