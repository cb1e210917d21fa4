"""CPU oracle for the FocusFlow correlation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``focusflow_official_b200/`` may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker or the timed CPU arm -- never as the product path.

Pinning status: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the reference itself, generated in
the build container by ``oracle/make_golden.py`` (imports the unmodified
``/root/reference`` code on CPU) and committed under ``tests/golden/``.
The PWC path of the reference cannot execute anywhere without CuPy + a GPU
(``PWCNet_Core/correlation.py:320-321`` raises on CPU), so for it the oracle is
a literal C restatement of the CUDA-C strings (``oracle/pwc_ref.c``) checked
against an independent numpy formulation: parity for PWC is "restated, not
reference-executed".
"""
