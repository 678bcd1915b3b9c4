"""CPU oracle for the Helmholtz FEM hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement (numpy / scipy-LAPACK) of the
algorithm in anishasahu/fem-helmholtz-eqn.  It exists to *check* the CUDA path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product package
``fem_helmholtz_eqn_b200`` never imports, links or executes anything in here.

Parity pin: the reference ships no tests or golden vectors, so
this oracle is pinned against outputs of the reference itself, executed
unmodified in the build container by ``oracle/make_golden.py`` and committed under
``tests/golden/`` (plus the six convergence rates stored in the notebook output,
nb:439-444).  ``tests/test_oracle_golden.py`` holds the pin.
"""
