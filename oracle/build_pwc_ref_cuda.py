#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- builds the REFERENCE's own PWC correlation CUDA kernels into oracle/_ref/libpwc_ref_cuda.so.

The reference (core/models/ff-pwcnet/PWCNet_Core/correlation.py) keeps its kernels as CUDA-C *strings* that CuPy
JIT-compiles after a textual SIZE_n(tensor) substitution (cupy_kernel, :234-269).  CuPy is not installed here, but nvcc
is: this recipe
  1. reads correlation.py WHERE IT LIES under /root/reference (nothing is copied into the repository),
  2. reads the four kernel strings out of its syntax tree AS DATA (string constants of the `kernel_Correlation_* = ...`
     assignments; no reference code is executed -- the module could not be imported anyway: `import cupy`) and
     specialises them for a fixed list of shapes with a local restatement of the only macro they use, SIZE_n(tensor)
     -> the n-th extent of that tensor (what the reference's cupy_kernel(), :234-269, substitutes textually),
  3. appends a small host launcher restating the launch geometry of _FunctionCorrelation.forward/backward
     (:278-380: grid / block / shared-memory sizes, per-sample backward launches),
  4. compiles everything for sm_100a into oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).
tests/ then run the reference's kernels on the GPU next to ours, and oracle/make_golden_pwc.py records their outputs
as tests/golden/pwc_ref_cuda.npz so that the CPU suite can pin oracle/pwc_ref.c to reference-executed results.
"""
from __future__ import annotations

import ast
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
SRC = os.path.join(REF, "core/models/ff-pwcnet/PWCNet_Core/correlation.py")
OUT_DIR = os.path.join(HERE, "_ref")

# (B, C, H, W): small cases incl. C not a multiple of 32, ragged sizes, one config-3 level shape
SHAPES = [(2, 20, 11, 14), (1, 32, 8, 32), (2, 70, 19, 40), (1, 196, 7, 16), (1, 64, 28, 64),
          # the five config-3 level shapes (B = 16), for tests/pwc_reference_timing.py
          (16, 32, 112, 256), (16, 64, 56, 128), (16, 96, 28, 64), (16, 128, 14, 32), (16, 196, 7, 16)]


def load_reference_pieces():
    """{name: CUDA-C text} of the reference's kernel strings, taken from the syntax tree as constants (nothing runs)."""
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)      # the reference's regex literals are not raw strings
        tree = ast.parse(open(SRC).read(), SRC)
    kernels = {}
    for node in tree.body:
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name) \
                and node.targets[0].id.startswith("kernel_Correlation_") and isinstance(node.value, ast.Constant) \
                and isinstance(node.value.value, str):
            kernels[node.targets[0].id] = node.value.value
    want = {"kernel_Correlation_rearrange", "kernel_Correlation_updateOutput", "kernel_Correlation_updateGradOne",
            "kernel_Correlation_updateGradTwo"}
    if set(kernels) != want:
        raise RuntimeError(f"expected the kernel strings {sorted(want)} in {SRC}, found {sorted(kernels)}")
    return kernels


def substitute_sizes(text: str, sizes: dict) -> str:
    """SIZE_n(name) -> str(sizes[name][n]): the shape specialisation the reference applies before NVRTC."""
    def rep(m):
        return str(sizes[m.group(2)][int(m.group(1))])
    out = re.sub(r"SIZE_([0-4])\(([^\)]*)\)", rep, text)
    if "VALUE_" in out:
        raise RuntimeError("the kernel strings use VALUE_n(), which this recipe does not restate")
    return out


def specialise(ns, idx, shape):
    b, c, h, w = shape
    one, rbot = (b, c, h, w), (b, h + 8, w + 8, c)
    out = (b, 81, h, w)
    grad = (b, c, h, w)
    srcs = []

    def emit(name, variables, suffix):
        s = substitute_sizes(ns[name], {k: v for k, v in variables.items() if v is not None})
        s = s.replace("#define ROUND_OFF 50000", "")          # defined once at the top of the generated file
        return s.replace(name + "(", f"{name}_{suffix}_s{idx}(")

    srcs.append(emit("kernel_Correlation_rearrange", {"input": one, "output": rbot}, "r"))
    srcs.append(emit("kernel_Correlation_updateOutput", {"rbot0": rbot, "rbot1": rbot, "top": out}, "f"))
    srcs.append(emit("kernel_Correlation_updateGradOne", {"rbot0": rbot, "rbot1": rbot, "gradOutput": out, "gradOne": grad, "gradTwo": None}, "g1"))
    srcs.append(emit("kernel_Correlation_updateGradTwo", {"rbot0": rbot, "rbot1": rbot, "gradOutput": out, "gradOne": None, "gradTwo": grad}, "g2"))
    return "\n".join(srcs)


LAUNCHER = r'''
// ---- host launcher: the launch geometry of _FunctionCorrelation.forward / backward (correlation.py:278-380) ----
#include <cuda_runtime.h>
extern "C" int pwc_ref_num_shapes() { return NUM_SHAPES; }
extern "C" int pwc_ref_shape(int idx, int* b, int* c, int* h, int* w) {
    static const int S[NUM_SHAPES][4] = { SHAPE_TABLE };
    if (idx < 0 || idx >= NUM_SHAPES) return -1;
    *b = S[idx][0]; *c = S[idx][1]; *h = S[idx][2]; *w = S[idx][3];
    return 0;
}
// rbot0 / rbot1: zero-initialised [B, H+8, W+8, C]; out: zero-initialised [B, 81, H, W] (the reference uses new_zeros)
extern "C" int pwc_ref_forward(int idx, const float* one, const float* two, float* rbot0, float* rbot1, float* out) {
    int b, c, h, w;
    if (pwc_ref_shape(idx, &b, &c, &h, &w)) return -1;
    const int n = h * w;
    const dim3 rgrid((n + 16 - 1) / 16, c, b), fgrid(w, h, b);
    const int nout = 81 * h * w;
    switch (idx) {
FORWARD_CASES
        default: return -1;
    }
    return (int)cudaDeviceSynchronize();
}
extern "C" int pwc_ref_backward(int idx, const float* rbot0, const float* rbot1, const float* gout, float* gone, float* gtwo) {
    int b, c, h, w;
    if (pwc_ref_shape(idx, &b, &c, &h, &w)) return -1;
    const int n = c * h * w;
    const int grid = (n + 512 - 1) / 512;
    for (int s = 0; s < b; ++s) {
        switch (idx) {
BACKWARD_CASES
            default: return -1;
        }
    }
    return (int)cudaDeviceSynchronize();
}
'''


def main():
    if not os.path.exists(SRC):
        print(f"reference not present at {SRC}: keeping whatever oracle/_ref holds", file=sys.stderr)
        return 0
    ns = load_reference_pieces()
    os.makedirs(OUT_DIR, exist_ok=True)
    body = ["// GENERATED by oracle/build_pwc_ref_cuda.py from the reference's kernel strings -- do not commit",
            "#define ROUND_OFF 50000"]
    fwd, bwd = [], []
    for i, shp in enumerate(SHAPES):
        body.append(specialise(ns, i, shp))
        fwd.append(f"        case {i}:\n"
                   f"            kernel_Correlation_rearrange_r_s{i}<<<rgrid, 16>>>(n, one, rbot0);\n"
                   f"            kernel_Correlation_rearrange_r_s{i}<<<rgrid, 16>>>(n, two, rbot1);\n"
                   f"            kernel_Correlation_updateOutput_f_s{i}<<<fgrid, 32, c * 4>>>(nout, rbot0, rbot1, out);\n"
                   f"            break;")
        bwd.append(f"            case {i}:\n"
                   f"                if (gone) kernel_Correlation_updateGradOne_g1_s{i}<<<grid, 512>>>(n, s, rbot0, rbot1, gout, gone, nullptr);\n"
                   f"                if (gtwo) kernel_Correlation_updateGradTwo_g2_s{i}<<<grid, 512>>>(n, s, rbot0, rbot1, gout, nullptr, gtwo);\n"
                   f"                break;")
    launcher = (LAUNCHER.replace("NUM_SHAPES", str(len(SHAPES)))
                .replace("SHAPE_TABLE", ", ".join("{%d, %d, %d, %d}" % s for s in SHAPES))
                .replace("FORWARD_CASES", "\n".join(fwd)).replace("BACKWARD_CASES", "\n".join(bwd)))
    cu = os.path.join(OUT_DIR, "pwc_ref_kernels.cu")
    open(cu, "w").write("\n".join(body) + launcher)
    so = os.path.join(OUT_DIR, "libpwc_ref_cuda.so")
    cmd = ["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
           "-cudart", "static", "-o", so, cu]
    try:
        subprocess.check_call(cmd)
    finally:
        os.remove(cu)          # the specialised kernel text is the reference's source: keep only the binary
    print("built", so)
    return 0


if __name__ == "__main__":
    sys.exit(main())
