"""wae-b200: B200-native (sm_100a) implementation of the data-parallel hot path of WavesAndEigenvalues.jl --
tetrahedral P1/P2 FEM assembly of the Helmholtz operator family, L(z) = sum f_i(z) A_i on a shared pattern and the
shifted sparse solves inside householder / mslp / beyn -- behind the reference's own
Helmholtz.discretize -> LinearOperatorFamily -> householder/beyn/mslp interface.

All numerics run in libwae_b200.so (CUDA, C ABI: include/wae_b200.h).  There is no CPU fallback: importing works
anywhere, but creating a context without the built library or without a GPU raises.
"""
from . import _lib  # noqa: F401
from .helmholtz import discretize  # noqa: F401
from .meshutils import Mesh, SymInfo, aggregate_elements, extend_mesh, kuhn_box, kuhn_unit_cell, octosplit  # noqa: F401
from .nlevp import (bloch_expand, conv_radius, LinearOperatorFamily, VectorFamily, Solution, Term, beyn, compute_moment_matrices, exp_az, exp_delay,  # noqa: F401
                    get_context, householder, inpoly, moments2eigs, mslp, pade, pade_bang, poly_roots, polyval, perturb_bang, perturb_fast_bang, perturb_norm_bang, pow0,
                    pow1, pow2, pow_a,
                    reset_context, wn)
from .shape import (bound_mass_normalize, discrete_adjoint_shape_sensitivity, get_normal_vectors, get_surface_points,  # noqa: F401
                    normal_sensitivity, normalize_sensitivity)

__all__ = ["Mesh", "discretize", "LinearOperatorFamily", "VectorFamily", "Term", "Solution", "householder", "mslp", "beyn", "pow0", "pow1",
           "pow2", "exp_delay", "kuhn_box", "aggregate_elements", "get_context"]
