"""Micro-benchmark: random-gather capacity of the B200 memory system (requests/s), to tell whether the
lookup kernel is bound by DRAM random-access efficiency or by its own structure."""
import torch, time
dev = "cuda:0"
buf = torch.randn(512 * 1024 * 1024, device=dev)  # 2 GB
def run(idx, label, bytes_per_req=64):
    for _ in range(2):
        out = torch.index_select(buf, 0, idx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = torch.index_select(buf, 0, idx)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    n = idx.numel()
    print(f"{label}: {ms*1e3:.1f} us for {n/1e6:.1f}M gathers -> {n/ms/1e6:.2f} G req/s, {n*bytes_per_req/ms/1e6:.0f} GB/s at {bytes_per_req}B/req")
n = 8 * 1024 * 1024
g = torch.Generator(device=dev); g.manual_seed(0)
idx = torch.randint(0, buf.numel(), (n,), device=dev, generator=g)
run(idx, "random 4B (1 per 64B granule)")
# 10-float runs at random starts (like a window row): 11 consecutive elements per start
starts = torch.randint(0, buf.numel() - 16, (n // 8,), device=dev, generator=g)
idx2 = (starts[:, None] + torch.arange(11, device=dev)[None, :]).reshape(-1)
run(idx2, "random 44B rows", bytes_per_req=1)
# window-like: 11 rows x 11 cols with row pitch 156 at random origins in random 7332-element maps
org = torch.randint(0, buf.numel() - 156 * 12, (n // 64,), device=dev, generator=g)
win = (org[:, None, None] + (torch.arange(11, device=dev) * 156)[None, :, None] + torch.arange(11, device=dev)[None, None, :]).reshape(-1)
run(win, "random 11x11 windows pitch 156", bytes_per_req=1)
# tiled: 3x3 tiles of 64B contiguous per tile-row (3*16 floats contiguous), 3 tile-rows far apart
tl = (org[:, None, None] + (torch.arange(3, device=dev) * 156 * 4)[None, :, None] + torch.arange(48, device=dev)[None, None, :]).reshape(-1)
run(tl, "tiled-like 3 x 192B chunks", bytes_per_req=1)
