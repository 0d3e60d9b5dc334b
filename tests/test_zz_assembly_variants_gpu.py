"""GPU check of the opt-in variants of the M/K assembly kernel (WAE_ASM_VARIANT = 1, 2, 3: unrolled summation pass, element pass
split into three parts per P2 element; csrc/assembly_kernels.cu).  The default (variant 0) is the measured and parity-tested kernel;
its SASS is unchanged by the introduction of the variants.

NOTE (round 1): the variants were derived from the phase shares of profiles/r01_ncu_assembly_phase_shares.txt after the round's GPU
budget was spent -- they have not run on a B200 yet.  The file sorts last so that `pytest -x` reaches it after everything else."""
import os

import numpy as np
import pytest

from cases import load_raw_mesh, rijke_dscrp, speedofsound

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order", ["lin", "quad"])
def test_variants_reproduce_the_default_kernel(order):
    import wae_b200 as W
    meshes = [W.Mesh("m", scale=0.001, raw=load_raw_mesh("rijke_mm")),
              W.kuhn_box((12, 10, 14), (0, 0, 0), (1, 1, 1), jitter=0.1, seed=5)]
    old = os.environ.pop("WAE_ASM_VARIANT", None)
    try:
        for mesh in meshes:
            c = mesh.generate_field(speedofsound) if "Flame" in mesh.domains else np.full(len(mesh.tetrahedra), 340.0)
            L = W.discretize(mesh, {"Interior": ("interior", ())}, c, order=order)
            ref = [t.coeff.csc()[2].copy() for t in L.terms]
            for var in ("1", "2", "3"):
                os.environ["WAE_ASM_VARIANT"] = var
                L.discretization.reassemble(c)
                for t, r in zip(L.terms, ref):
                    v = t.coeff.csc()[2]
                    assert np.abs(v - r).max() <= 1e-13 * np.abs(r).max(), (var, t.operator)
            os.environ.pop("WAE_ASM_VARIANT", None)
            L.discretization.reassemble(c)
            for t, r in zip(L.terms, ref):
                assert np.array_equal(t.coeff.csc()[2], r)  # the default kernel is bit-reproducible
    finally:
        os.environ.pop("WAE_ASM_VARIANT", None)
        if old is not None:
            os.environ["WAE_ASM_VARIANT"] = old
