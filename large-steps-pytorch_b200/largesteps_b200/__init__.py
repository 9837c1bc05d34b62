"""largesteps_b200 -- B200-native drop-in for the hot path of rgl-epfl/large-steps-pytorch.

Same operator surface as the reference package `largesteps` (version 0.2.2, largesteps/__init__.py:1-9):

    from largesteps_b200.geometry import compute_matrix, laplacian_uniform, laplacian_cot
    from largesteps_b200.parameterize import to_differential, from_differential
    from largesteps_b200.solvers import Solver, CholeskySolver, ConjugateGradientSolver, PCGSolver, solve
    from largesteps_b200.optimize import AdamUniform

All device work is hand-written sm_100a CUDA in `libls_b200.so` (sources in ../csrc, C ABI in
../../include/largesteps_b200.h) loaded with ctypes.  There is no CPU or torch fallback.
The sibling package `largesteps` re-exports these modules under the reference's import names.
"""
__version__ = "0.1.0"
reference_version = "0.2.2"
