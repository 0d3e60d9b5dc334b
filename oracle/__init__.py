"""CPU oracle for the WavesAndEigenvalues.jl hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy/scipy restatement of the reference's algorithm for the
path  Helmholtz.discretize -> LinearOperatorFamily -> householder/mslp/beyn.
It is the checker for the CUDA product in ``wavesandeigenvalues.jl_b200/`` and
the timed ``cpu_baseline`` in ``bench.py``; nothing in the product path may
import it.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it.

Parity pinning: the reference (pure Julia) cannot run in this image, so the
oracle is pinned against the reference's own stored outputs -- golden vectors
G1-G6, G8-G10 of ``examples/tutorials/tutorial_04_perturbation_theory.ipynb`` and
``docs/src/tutorial_04_perturbation_theory.md``, the prose values and the
self-consistency sequences of tutorials 00, 01, 07 and 08 (tests/test_oracle_golden.py,
tests/test_bloch.py) -- and against element tables obtained by evaluating the reference's own
closed-form expressions in ``src/FEM/FEM.jl`` (tests/golden/fem_tables.npz,
made by tests/golden/make_fem_tables.py).  Third-party arithmetic the reference
delegates to and that is not vendored (Arpack.jl 0.4.0 / Arpack_jll 3.5.0,
SuiteSparse UMFPACK via Julia stdlib, FastGaussQuadrature 0.4.7) is replaced by
scipy's ARPACK + SuperLU and numpy's ``leggauss``; Beyn end-to-end results are
"parity unpinned" beyond the tutorial's prose values (272 / 695 Hz); so are the
shape sensitivity (pinned to eigenvalue finite differences instead) and the forced
response (pinned to the closed-form duct solution): the reference stores no output
of either.  ``tests/host_standin.py`` reuses the element routines of this package
as a test double of the device context for the host-mirror CPU tests.
"""
