#!/usr/bin/env python
"""Measurement script (lives under tests/ because it runs oracle/_ref): the REFERENCE's own PWC CUDA kernels
(oracle/_ref/libpwc_ref_cuda.so, built from correlation.py) timed next to this repo's kernels at the five config-3
level shapes on the current GPU.  Forward = rearrange x2 + updateOutput (+ the three zero-fills the reference does);
backward = per-sample updateGradOne / updateGradTwo launches.  One JSON line per level."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import focusflow_official_b200 as ff  # noqa: E402
from oracle import pwc_ref_cuda as ref  # noqa: E402


def wall_ms(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / reps


def main():
    if not ref.available():
        print(json.dumps({"unavailable": "oracle/_ref/libpwc_ref_cuda.so not built"}))
        return
    tot_ref = tot_ours = 0.0
    for idx, shape in enumerate(ref.shapes()):
        if shape[0] != 16:
            continue
        b, c, h, w = shape
        one = torch.randn(shape, device="cuda")
        two = torch.randn(shape, device="cuda")
        gout = torch.randn(b, 81, h, w, device="cuda")
        t_ref = wall_ms(lambda: ref.run(idx, one, two, gout))          # forward + backward, host-synchronised
        t_ref_fwd = wall_ms(lambda: ref.run(idx, one, two, gout, backward=False))
        a, bb = one.clone().requires_grad_(True), two.clone().requires_grad_(True)

        def ours():
            a.grad = bb.grad = None
            ff.FunctionCorrelation(a, bb).backward(gout)

        t_ours = wall_ms(ours)
        with torch.no_grad():
            t_ours_fwd = wall_ms(lambda: ff.FunctionCorrelation(one, two))
        tot_ref += t_ref
        tot_ours += t_ours
        print(json.dumps({"level": f"C={c} {h}x{w} B={b}", "reference_kernels_fwd_bwd_ms": round(t_ref, 3),
                          "reference_kernels_fwd_ms": round(t_ref_fwd, 3), "this_repo_fwd_bwd_ms": round(t_ours, 3), "this_repo_fwd_ms": round(t_ours_fwd, 3)}), flush=True)
    print(json.dumps({"level": "all five", "reference_kernels_fwd_bwd_ms": round(tot_ref, 3), "this_repo_fwd_bwd_ms": round(tot_ours, 3)}))


if __name__ == "__main__":
    main()
