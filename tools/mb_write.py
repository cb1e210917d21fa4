import torch
dev="cuda:0"
x = torch.empty(2*1024**3 // 4, device=dev)  # 2 GB fp32
y = torch.empty_like(x)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
ms = t(lambda: x.zero_()); print(f"memset 2GB: {ms:.3f} ms -> {2*1024**3/ms/1e6:.0f} GB/s write")
ms = t(lambda: x.fill_(1.5)); print(f"fill 2GB: {ms:.3f} ms -> {2*1024**3/ms/1e6:.0f} GB/s write")
ms = t(lambda: y.copy_(x)); print(f"copy 2GB: {ms:.3f} ms -> {2*2*1024**3/ms/1e6:.0f} GB/s read+write")
ms = t(lambda: x.sum()); print(f"sum 2GB: {ms:.3f} ms -> {2*1024**3/ms/1e6:.0f} GB/s read")
