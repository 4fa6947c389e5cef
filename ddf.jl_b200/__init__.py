"""ddf.jl_b200 -- B200-native (sm_100a) implementation of DDF.jl's discrete
exterior calculus operator hot path, behind the reference's own call surface.

Import as ``import ddf.jl_b200 as ddf`` (the ``ddf/`` shim at the repo root maps
the dotted name onto this directory).  All compute runs in libddf_b200.so
(hand-written CUDA, C ABI in include/ddf_b200.h); there is no CPU fallback.
"""
from ._lib import DDFError, LIB_PATH, load  # noqa: F401
from .core import (Context, Dl, Fun, Manifold, Op, Pr, default_context, op_from_csc, op_from_diag,  # noqa: F401
                   rand_fun, set_default_context, unit_fun, zero_fun)
from .manifoldops import (boundary, boundary_manifold, boundary_of, boundary_simplex_manifold, boundary_tables, coderiv,  # noqa: F401
                          coderiv_of, delaunay_hypercube_manifold, delaunay_hypercube_tables, delaunay_mesh, deriv, deriv_format, deriv_index_bytes, deriv_of,
                          hodge, hodge_of, hypercube_manifold, hypercube_tables, integral, invhodge, invhodge_of, isboundary,
                          large_delaunay_hypercube_manifold, large_delaunay_hypercube_tables, large_hypercube_manifold, laplace, laplace_of, one, orthogonal_simplex_manifold,
                          orthogonal_simplex_tables, refined_manifold,
                          refined_simplex_manifold, regular_simplex, sample, simplex_manifold, zero)
from .wave import (lincomb, wave, wave_dt, wave_format, wave_leapfrog, wave_rhs, wave_step_bytes,  # noqa: F401
                   wave_tsit5)  # noqa: F401
from .poisson import bicgstabl, poisson, poisson_jacobi, poisson_operator, poisson_solve  # noqa: F401
from .vtkout import write_vtu  # noqa: F401
from . import dist  # noqa: F401
