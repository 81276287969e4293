"""CPU oracle for the guided-sampling hot path — TEST INFRASTRUCTURE ONLY.

This package is a plain torch-CPU / numpy fp32 restatement of the reference algorithm
(xypeng9903/k-diffusion-inverse-problems @ d5ae606).  Every function cites the reference
file:line it follows.  It is imported only by ``tests/`` (including the developer scripts under
``tests/tools/``), ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` — never by the product package (``k-diffusion-inverse-problems_b200/``) nor by the measurement
scripts under ``tools/``; the product must fail loudly when the CUDA library is missing
(``tests/test_host_cpu.py::test_no_cpu_fallback`` enforces both).

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned
against outputs of the reference's own code executed in the build container
(``tests/golden/make_golden.py`` imports ``/root/reference`` through ``oracle/refshim.py`` and
commits the vectors under ``tests/golden/``).  ``tests/test_oracle_golden.py`` checks every oracle
function against those vectors.  One piece stays *parity unpinned*: the PyWavelets packed layout of
the level-3 Haar DWT (pywt is neither vendored nor installed; see ``transforms_ref.py``).
"""
