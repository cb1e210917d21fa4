#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- records the outputs of the REFERENCE's own PWC CUDA kernels as tests/golden/pwc_ref_cuda.npz.

Run ON A GPU after oracle/build_pwc_ref_cuda.py (in the build container, where /root/reference exists):
    gpurun -- 'python oracle/make_golden_pwc.py gpurun_out/pwc_ref_cuda.npz'
then copy the file to tests/golden/.  Inputs are seeded, so the fixture is reproducible."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pwc_ref_cuda  # noqa: E402

KEEP = (0, 1, 3)   # small shapes only: the fixture stays ~300 KB


def inputs(idx, shape):
    rng = np.random.default_rng(1000 + idx)
    b, c, h, w = shape
    return (rng.standard_normal(shape).astype(np.float32), rng.standard_normal(shape).astype(np.float32),
            rng.standard_normal((b, 81, h, w)).astype(np.float32))


def main(path):
    data = {}
    for idx, shape in enumerate(pwc_ref_cuda.shapes()):
        if idx not in KEEP:
            continue
        one, two, gout = inputs(idx, shape)
        out, g1, g2 = pwc_ref_cuda.run(idx, *(torch.from_numpy(a).cuda() for a in (one, two, gout)))
        data[f"shape_{idx}"] = np.array(shape)
        data[f"out_{idx}"] = out.cpu().numpy()
        data[f"gone_{idx}"] = g1.cpu().numpy()
        data[f"gtwo_{idx}"] = g2.cpu().numpy()
    np.savez_compressed(path, **data)
    print("wrote", path, {k: v.shape for k, v in data.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "pwc_ref_cuda.npz")
