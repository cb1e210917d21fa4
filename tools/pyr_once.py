import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from focusflow_official_b200 import _lib
q, h, w = [int(v) for v in sys.argv[1:4]]
l0 = torch.randn(q, 1, h, w, device="cuda")
lv = [l0] + [torch.empty(q, 1, h >> i, w >> i, device="cuda") for i in range(1, 4)]
_lib.check(_lib.lib().ffcorr_pyramid_f32(_lib.ptr_array(lv), 4, q, h, w, _lib.current_stream()), "pyr")
torch.cuda.synchronize()
cur = l0
for i in range(1, 4):
    cur = torch.nn.functional.avg_pool2d(cur, 2, stride=2)
    print(i, bool(torch.equal(cur, lv[i])))
