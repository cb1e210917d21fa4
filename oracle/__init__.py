"""CPU oracle for the FocusFlow correlation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``focusflow_official_b200/`` may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker or the timed CPU arm -- never as the product path.

Pinning status: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the reference itself, generated in
the build container by ``oracle/make_golden.py`` (imports the unmodified
``/root/reference`` code on CPU) and committed under ``tests/golden/``.
The PWC path of the reference has no CPU branch (``PWCNet_Core/correlation.py:320-321``
raises) and needs CuPy, which is absent -- but its kernels are plain CUDA-C strings, so
``oracle/build_pwc_ref_cuda.py`` specialises them with the reference's own ``cupy_kernel()``
and compiles them with nvcc into ``oracle/_ref/libpwc_ref_cuda.so``; ``oracle/make_golden_pwc.py``
ran them on a B200 and committed ``tests/golden/pwc_ref_cuda.npz``.  The C restatement
(``oracle/pwc_ref.c``) and the numpy closed form both reproduce those outputs (2e-7 / 2e-6),
and the GPU suite additionally runs the reference kernels live next to ours.
"""
