import torch
x = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda").normal_()     # 2 GiB
y = torch.empty_like(x)
def timeit(f, n=10):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = timeit(lambda: y.copy_(x)); print(f"copy   : {2 * x.numel() * 2 / ms * 1e-6:.0f} GB/s (read+write)")
xf = x.view(torch.int32)
ms = timeit(lambda: xf.sum()); print(f"sum i32: {x.numel() * 2 / ms * 1e-6:.0f} GB/s (read only)")
ms = timeit(lambda: x.float().max() if False else torch.max(x)); print(f"max bf16: {x.numel() * 2 / ms * 1e-6:.0f} GB/s (read only)")
# 100 MB reads (the attention working set at 64 keys) from a cold region each time
z = x[: 50 * (1 << 20)]
ms = timeit(lambda: torch.max(z)); print(f"max over 100 MB: {z.numel() * 2 / ms * 1e-3:.1f} GB/s  ({ms*1e3:.1f} us)")
