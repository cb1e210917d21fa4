"""Break down one FocusRAFT forward (config 2) by CUDA kernel with torch.profiler."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from weights import synthetic_pair
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = True
dev = torch.device("cuda:0")
model = bench.make_model(dev, False); model.flow_net.update_channels_last = True
im1, im2, m1, _ = (t.to(dev) for t in synthetic_pair(8, 376, 1248, seed=1234))
with torch.no_grad():
    for _ in range(3):
        model(im1, im2, m1, None, raft_iters=12, test_mode=True)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        model(im1, im2, m1, None, raft_iters=12, test_mode=True)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
