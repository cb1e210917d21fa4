"""One PWC level alone (for ncu): python tools/pwc_level.py B C H W [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import focusflow_official_b200 as ff  # noqa: E402

b, c, h, w = [int(v) for v in sys.argv[1:5]]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
one = torch.randn(b, c, h, w, device="cuda")
two = torch.randn(b, c, h, w, device="cuda")
for _ in range(reps):
    out = ff.FunctionCorrelation(one, two)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = ff.FunctionCorrelation(one, two)
e1.record()
torch.cuda.synchronize()
print(f"B={b} C={c} {h}x{w}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call (back to back)")
