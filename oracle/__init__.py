"""CPU oracle for the LivelySpeaker RAG sampling path.  TEST INFRASTRUCTURE ONLY.

This package is a plain CPU restatement (numpy fp64 for the schedule / index
path, torch-CPU fp32 for the denoiser and the sampler update) of the reference
algorithm.  It exists so that the CUDA product path in ``livelyspeaker_b200``
can be checked against something that was itself pinned to the reference.

Rules (enforced by tests/test_host_logic.py::test_product_never_imports_oracle_or_reference):
  * only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
    ``--impl reference`` legs of ``bench.py`` may import anything from here;
  * nothing under ``livelyspeaker_b200/`` imports it - the product path fails
    loudly when its CUDA library is missing instead of falling back to this.

Pinning: the reference (zyhbili/LivelySpeaker @ 7f6ccd1) ships NO golden
vectors or tests for this path, so the oracle is pinned
against outputs of the reference itself, imported from /root/reference in the
build container by ``tests/golden/make_golden.py`` (committed together with the
fixtures it wrote under ``tests/golden/``).  ``tests/test_oracle_golden.py``
replays those fixtures against this package on every CPU test run.
"""
