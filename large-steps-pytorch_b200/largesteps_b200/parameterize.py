"""Differential parameterization -- same surface as the reference's largesteps/parameterize.py.

    to_differential(L, v)                      u = M v            (parameterize.py:19-30)
    from_differential(L, u, method='Cholesky') v = M^-1 u         (parameterize.py:32-61), differentiable w.r.t. u

The solver cache keeps the reference's semantics (parameterize.py:5-17): keyed by (id(L), method), dropped by a
weakref callback when the matrix is garbage collected.
"""
import weakref

import torch

from . import _native as N
from .geometry import csr_of
from .solvers import CholeskySolver, ConjugateGradientSolver, PCGSolver, solve

# Cache for the system solvers
_cache = {}


def cache_put(key, value, A):
    # Called when 'A' is garbage collected
    def cleanup_callback(wr):
        _cache.pop(key, None)

    wr = weakref.ref(A, cleanup_callback)
    _cache[key] = (value, wr)


class _SpMM(torch.autograd.Function):
    """y = M x through the library's CSR SpMM.  The backward applies M itself: every matrix this package builds (system
    matrices, both Laplacians) is symmetric, like the reference's (geometry.py:56,94); a non-symmetric foreign matrix would
    need M^T and must go through torch's own `L @ v`."""

    @staticmethod
    def forward(ctx, L, v):
        ctx.L = L
        return spmm(L, v)

    @staticmethod
    def backward(ctx, g):
        # d/dv (L v) = L^T g; system matrices here are symmetric (M = M^T), as are both Laplacians
        return None, spmm(ctx.L, g.contiguous())


def spmm(L, v):
    """Non-differentiable y = L @ v on the device (ls_spmm_csr_f32). v: (V,k) or (V,) float32 CUDA."""
    rowptr, col, val = csr_of(L)
    N.require_cuda(v, "v")
    if v.device != val.device:
        raise RuntimeError(f"v is on {v.device} but the matrix is on {val.device}")
    if v.dtype != torch.float32:
        raise TypeError(f"v must be float32, got {v.dtype}")
    squeeze = v.dim() == 1
    x = (v.unsqueeze(1) if squeeze else v).detach().contiguous()
    if x.dim() != 2 or x.shape[0] != L.shape[1]:
        raise ValueError(f"shape mismatch: L is {tuple(L.shape)}, v is {tuple(v.shape)}")
    y = torch.empty_like(x)
    k = x.shape[1]
    with torch.cuda.device(x.device):
        N.check(N.lib().ls_spmm_csr_f32(L.shape[0], N.ptr(rowptr), N.ptr(col), N.ptr(val), N.ptr(x), k, N.ptr(y), k, k,
                                        N.stream_ptr(x.device)), "ls_spmm_csr_f32")
    return y.squeeze(1) if squeeze else y


def to_differential(L, v):
    """Convert vertex coordinates to the differential parameterization: u = L @ v (parameterize.py:30).

    L : torch.sparse.Tensor   (I + l*L) matrix;   v : torch.Tensor   vertex coordinates (V,3) float32 CUDA.
    """
    if v.requires_grad:
        return _SpMM.apply(L, v)
    return spmm(L, v)


def from_differential(L, u, method='Cholesky'):
    """Convert differential coordinates back to Cartesian: solve L v = u (parameterize.py:32-61).

    If this is the first time we call this function on a given matrix L, the solver is cached. It will be destroyed
    once the matrix is garbage collected.

    method : {'Cholesky', 'CG', 'PCG'}
        'Cholesky' and 'CG' are the reference's names and keep their contracts (cold-start direct-solve accuracy;
        warm-started CG).  'PCG' is the native solver with its defaults.  All three run csrc/ls_pcg.cu.
    """
    key = (id(L), method)
    if key not in _cache.keys():
        if method == 'Cholesky':
            solver = CholeskySolver(L)
        elif method == 'CG':
            solver = ConjugateGradientSolver(L)
        elif method == 'PCG':
            solver = PCGSolver(L)
        else:
            raise ValueError(f"Unknown solver type '{method}'.")
        cache_put(key, solver, L)
    else:
        solver = _cache[key][0]
    return solve(solver, u)
