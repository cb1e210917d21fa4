import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import focusflow_official_b200 as ff
shape = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (16, 32, 112, 256))]
one = torch.randn(*shape, device="cuda"); two = torch.randn(*shape, device="cuda")
for _ in range(2):
    out = ff.FunctionCorrelation(one, two)
torch.cuda.synchronize(); print(float(out.abs().mean()))
