import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_host_model import make_model
from weights import synthetic_pair
from oracle.corr_torch_cpu import TorchCorrBlock
from focusflow_official_b200 import CorrBlock
g = np.load(os.path.join(ROOT, "tests/golden/ffraft_e2e.npz"))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda:0"
model = make_model(dev)
def epe(a, b): 
    e = torch.linalg.norm(a.cpu() - b.cpu(), dim=1); return float(e.mean()), float(e.max())
for tag in ("a", "b"):
    b, hh, ww, iters = [int(v) for v in g[f"{tag}_shape"]]
    im = [x.to(dev) for x in synthetic_pair(b, hh, ww, seed=1234 + b)]
    gold = torch.from_numpy(g[f"{tag}_flow_up"])
    for it in (1, 2, 4, iters):
        res = {}
        for name, blk, prec in (("torch", TorchCorrBlock, None), ("b200-fp32", CorrBlock, "fp32"), ("b200-fp16", CorrBlock, "fp16"), ("b200-tf32", CorrBlock, "tf32")):
            model.flow_net.corr_block = blk
            model.flow_net.corr_precision = prec
            with torch.no_grad():
                lo, up = model(*im, raft_iters=it, test_mode=True)
            res[name] = up
        msg = f"{tag} iters={it}: |flow| mean {float(res['torch'].abs().mean()):.2f}"
        for k in ("b200-fp32", "b200-fp16", "b200-tf32"):
            msg += f" | {k} vs torch-gpu EPE mean/max {epe(res[k], res['torch'])[0]:.2e}/{epe(res[k], res['torch'])[1]:.2e}"
        if it == iters:
            msg += f" | torch-gpu vs CPU golden {epe(res['torch'], gold)[0]:.2e}/{epe(res['torch'], gold)[1]:.2e}"
        print(msg)
# torch matmul with TF32 (the reference GPU path) vs fp32
torch.backends.cuda.matmul.allow_tf32 = True
model.flow_net.corr_block = TorchCorrBlock
with torch.no_grad():
    up_tf32 = model(*im, raft_iters=iters, test_mode=True)[1]
print("torch-gpu TF32-matmul vs torch-gpu fp32 EPE", epe(up_tf32, res["torch"]))
