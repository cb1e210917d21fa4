"""One PWC level forward + backward (for ncu): python tools/pwc_bwd_level.py B C H W"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import focusflow_official_b200 as ff  # noqa: E402

b, c, h, w = [int(v) for v in sys.argv[1:5]]
one = torch.randn(b, c, h, w, device="cuda", requires_grad=True)
two = torch.randn(b, c, h, w, device="cuda", requires_grad=True)
go = torch.randn(b, 81, h, w, device="cuda")
for _ in range(3):
    one.grad = two.grad = None
    ff.FunctionCorrelation(one, two).backward(go)
torch.cuda.synchronize()
print("ok", float(one.grad.abs().mean()), float(two.grad.abs().mean()))
