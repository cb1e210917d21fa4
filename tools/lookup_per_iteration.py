"""Per-refinement-iteration duration of the lookup inside the FocusRAFT forward (config 2 shapes, CUDA events around every
launch).  Iteration 0 looks up at the integer grid (flow = 0): the normalise / un-normalise round trip of the reference
(utils.py:61-62 + ATen) lands a hair below many integers, so those taps deviate and the warp takes the exact per-tap path."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from weights import synthetic_pair  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    model = bench.make_model(dev, False)
    model.flow_net.update_block.to(memory_format=torch.channels_last)
    model.flow_net.update_channels_last = True
    if "--fuse-convc1" in sys.argv:
        model.flow_net.fuse_convc1 = True
    meter = bench.LaunchMeter()
    meter.install()
    im1, im2, m1, _ = (t.to(dev) for t in synthetic_pair(8, 376, 1248, seed=1234))
    iters = 12
    with torch.no_grad():
        for _ in range(3):
            model(im1, im2, m1, None, raft_iters=iters, test_mode=True)
        torch.cuda.synchronize()
        meter.enabled = True
        for _ in range(5):
            model(im1, im2, m1, None, raft_iters=iters, test_mode=True)
        torch.cuda.synchronize()
    per = [[] for _ in range(iters)]
    for i, (a, b) in enumerate(meter.lookup_events):
        per[i % iters].append(a.elapsed_time(b) * 1e3)
    print(json.dumps({"fused": "--fuse-convc1" in sys.argv, "lookup_us_by_iteration": [round(sorted(p)[len(p) // 2], 1) for p in per]}))


if __name__ == "__main__":
    main()
