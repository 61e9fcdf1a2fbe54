import torch, time
x = torch.empty(17_000_000_000 // 2, dtype=torch.float16, device='cuda')
y = torch.empty(3_500_000_000 // 2, dtype=torch.float16, device='cuda')
def t(f, n=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.zero_())
print(f"memset 17 GB: {ms:.3f} ms = {17e9/ms/1e6:.0f} GB/s")
ms = t(lambda: x.fill_(1.5))
print(f"fill 17 GB: {ms:.3f} ms = {17e9/ms/1e6:.0f} GB/s")
ms = t(lambda: y.sum())
print(f"read 3.5 GB (sum): {ms:.3f} ms = {3.5e9/ms/1e6:.0f} GB/s")
z = torch.empty_like(x)
ms = t(lambda: z.copy_(x))
print(f"copy 17 GB: {ms:.3f} ms = {34e9/ms/1e6:.0f} GB/s")
ms = t(lambda: x.sum())
print(f"read 17 GB (sum): {ms:.3f} ms = {17e9/ms/1e6:.0f} GB/s")
