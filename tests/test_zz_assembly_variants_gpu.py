"""GPU check of the M/K assembly kernel generations against each other: the star kernel (generation 3, the default;
csrc/assembly_star.cu) against the owner-computes pair program (generation 2, WAE_ASM_GEN=2; csrc/assembly_kernels.cu) and its opt-in
variants (WAE_ASM_VARIANT = 1, 2, 3: unrolled summation pass, element pass split into three parts per P2 element), and the star
kernel under the patch sizes / CTA shapes bench.py sweeps."""
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
    knobs = ("WAE_ASM_VARIANT", "WAE_ASM_GEN", "WAE_STAR_SMEM", "WAE_STAR_THREADS", "WAE_STAR_CTAS")
    old = {k: os.environ.pop(k, None) for k in knobs}
    try:
        for mesh in meshes:
            c = mesh.generate_field(speedofsound) if "Flame" in mesh.domains else np.full(len(mesh.tetrahedra), 340.0)
            os.environ["WAE_ASM_GEN"] = "3"
            L = W.discretize(mesh, {"Interior": ("interior", ())}, c, order=order)
            ref = [t.coeff.csc()[2].copy() for t in L.terms]  # star kernel, default layout
            os.environ["WAE_ASM_GEN"] = "2"
            for var in ("0", "1", "2", "3"):
                os.environ["WAE_ASM_VARIANT"] = var
                L.discretization.reassemble(c)
                for t, r in zip(L.terms, ref):
                    v = t.coeff.csc()[2]
                    assert np.abs(v - r).max() <= 1e-12 * np.abs(r).max(), (var, t.operator)
            os.environ.pop("WAE_ASM_VARIANT", None)
            os.environ["WAE_ASM_GEN"] = "3"
            for smem, thr in (("230400", "1024"), ("76800", "256"), ("57344", "256"), ("114688", "512")):
                os.environ["WAE_STAR_SMEM"], os.environ["WAE_STAR_THREADS"] = smem, thr
                L.discretization.reassemble(c)  # a changed budget rebuilds the star program
                for t, r in zip(L.terms, ref):
                    v = t.coeff.csc()[2]
                    assert np.abs(v - r).max() <= 1e-12 * np.abs(r).max(), (smem, thr, t.operator)
            os.environ.pop("WAE_STAR_SMEM", None)
            os.environ.pop("WAE_STAR_THREADS", None)
            L.discretization.reassemble(c)
            for t, r in zip(L.terms, ref):
                assert np.array_equal(t.coeff.csc()[2], r)  # the star kernel is bit-reproducible
            os.environ.pop("WAE_ASM_GEN", None)  # the default generation of this element order (P2: stars, P1: pairs), twice
            L.discretization.reassemble(c)
            first = [t.coeff.csc()[2].copy() for t in L.terms]
            L.discretization.reassemble(c)
            for t, r in zip(L.terms, first):
                assert np.array_equal(t.coeff.csc()[2], r)
    finally:
        for k in knobs:
            os.environ.pop(k, None)
            if old[k] is not None:
                os.environ[k] = old[k]
