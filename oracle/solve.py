"""Oracle: CPU solves for M x = b (TEST INFRASTRUCTURE, see oracle/__init__.py).

  * DirectSolver          fp64 (or fp32) sparse direct solve -- the stand-in for the reference's
                          CholeskySolver (solvers.py:26-39 -> cholespy/CHOLMOD, absent here).
                          SuperLU in symmetric mode: factor once, solve many (V,k) right-hand sides.
  * dense_cholesky_solve  numpy LL^T for V <= ~3K, second opinion on DirectSolver.
  * reference_cg / ReferenceCG   restatement of ConjugateGradientSolver (solvers.py:41-126):
                          plain CG per axis, ABSOLUTE tolerance 1e-5, warm start kept for fwd/bwd.
  * to_differential       parameterize.py:30  (u = M @ v)
  * jacobi_pcg_f32        numpy model of the device algorithm (fp32 vectors, fp64 dot products,
                          per-column alpha/beta/convergence) used to sanity-check iteration counts
                          and attainable accuracy on the CPU before spending GPU time.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

f32 = np.float32


def _csr(rows, cols, vals, V, dtype):
    return sp.csr_matrix((np.asarray(vals).astype(dtype), (np.asarray(rows), np.asarray(cols))), shape=(V, V))


class DirectSolver:
    """Factor once (SuperLU, symmetric mode, MMD(A^T+A) ordering), then x = M^{-1} b for (V,k)."""

    def __init__(self, rows, cols, vals, V, dtype=np.float64):
        self.V = V
        self.dtype = dtype
        A = _csr(rows, cols, vals, V, dtype).tocsc()
        self.lu = spla.splu(A, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                            options={"SymmetricMode": True})
        self.factor_nnz = int(self.lu.L.nnz + self.lu.U.nnz)

    def solve(self, b, backward=False):   # same signature as solvers.py:36 (M symmetric: bwd == fwd)
        return self.lu.solve(np.ascontiguousarray(b, dtype=self.dtype))


def dense_cholesky_solve(rows, cols, vals, V, b):
    A = _csr(rows, cols, vals, V, np.float64).toarray()
    Lc = np.linalg.cholesky(A)
    y = np.linalg.solve(Lc, np.asarray(b, dtype=np.float64))
    return np.linalg.solve(Lc.T, y)


def to_differential(rows, cols, vals, V, v):
    """parameterize.py:30 in fp32."""
    return (_csr(rows, cols, vals, V, f32) @ np.asarray(v, dtype=f32)).astype(f32)


def reference_cg(A, b, x0, tol=1e-5, maxit=100000):
    """solvers.py:58-84 (solve_axis), fp32.  `maxit` is a safety net the reference lacks."""
    x = x0.astype(f32).copy()
    r = (A @ x - b).astype(f32)                # solvers.py:70
    p = -r                                     # solvers.py:71
    r_norm = f32(np.sqrt(np.dot(r, r)))        # solvers.py:72
    it = 0
    while r_norm > tol and it < maxit:         # solvers.py:73 (absolute)
        Ap = (A @ p).astype(f32)               # solvers.py:74
        r2 = f32(r_norm * r_norm)              # solvers.py:75
        alpha = f32(r2 / f32(np.dot(p, Ap)))   # solvers.py:76
        x = (x + alpha * p).astype(f32)        # solvers.py:77
        r = (r + alpha * Ap).astype(f32)       # solvers.py:80
        r_norm = f32(np.sqrt(np.dot(r, r)))    # solvers.py:81
        beta = f32(f32(r_norm * r_norm) / r2)  # solvers.py:82
        p = (-r + beta * p).astype(f32)        # solvers.py:83
        it += 1
    return x, it


class ReferenceCG:
    """solvers.py:41-126: per-axis CG with separate fwd/bwd warm starts."""

    def __init__(self, rows, cols, vals, V):
        self.A = _csr(rows, cols, vals, V, f32)
        self.guess_fwd = None
        self.guess_bwd = None
        self.iters = []

    def solve(self, b, backward=False):
        b = np.asarray(b, dtype=f32)
        if self.guess_fwd is None:                              # solvers.py:102-105
            self.guess_bwd = np.zeros_like(b)
            self.guess_fwd = np.zeros_like(b)
        x0 = self.guess_bwd if backward else self.guess_fwd     # solvers.py:107-110
        if b.ndim != 2:                                         # solvers.py:112-113
            raise ValueError(f"Invalid array shape {b.shape} for ConjugateGradientSolver.solve: expected shape (a, b)")
        x = np.zeros_like(b)
        self.iters = []
        for axis in range(b.shape[1]):                          # solvers.py:115-118
            x[:, axis], it = reference_cg(self.A, b[:, axis], x0[:, axis])
            self.iters.append(it)
        if backward:                                            # solvers.py:120-124
            self.guess_bwd = x
        else:
            self.guess_fwd = x
        return x


def jacobi_pcg_f32(rows, cols, vals, V, b, x0=None, rtol=1e-7, maxit=10000, precond=True):
    """Model of the device PCG: all k columns in lock-step with per-column alpha/beta/freeze,
    fp32 vectors, fp64 dot products, relative residual test ||r||_2 <= rtol ||b||_2 per column."""
    A = _csr(rows, cols, vals, V, f32)
    b = np.asarray(b, dtype=f32)
    k = b.shape[1]
    dinv = (f32(1.0) / A.diagonal().astype(f32)) if precond else np.ones(V, dtype=f32)
    x = np.zeros_like(b) if x0 is None else np.asarray(x0, dtype=f32).copy()
    r = b.copy() if x0 is None else (b - (A @ x).astype(f32)).astype(f32)
    z = (dinv[:, None] * r).astype(f32)
    p = z.copy()
    d = lambda u, w: np.einsum("ij,ij->j", u.astype(np.float64), w.astype(np.float64))
    rz = d(r, z)
    bb = d(b, b)
    rr = d(r, r)
    active = rr > (rtol * rtol) * bb
    it = 0
    while active.any() and it < maxit:
        Ap = (A @ p).astype(f32)
        pAp = d(p, Ap)
        alpha = np.where(active & (pAp > 0), rz / np.where(pAp == 0, 1, pAp), 0.0).astype(f32)
        x = (x + alpha[None, :] * p).astype(f32)
        r = (r - alpha[None, :] * Ap).astype(f32)
        z = (dinv[:, None] * r).astype(f32)
        rz_new = d(r, z)
        rr = d(r, r)
        beta = np.where(active, rz_new / np.where(rz == 0, 1, rz), 0.0).astype(f32)
        p = (z + beta[None, :] * p).astype(f32)
        rz = rz_new
        it += 1
        active = active & (rr > (rtol * rtol) * bb)
    relres = np.sqrt(rr / np.where(bb == 0, 1, bb))
    return x, it, relres


def _bf16(x):
    """round-to-nearest-even to bfloat16, returned as float32 (what cvt.rn.bf16.f32 + a 16-bit shift give on the device)"""
    u = np.ascontiguousarray(x, dtype=f32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7fff
    return ((u + r) & np.uint32(0xffff0000)).view(f32)


def fused_pcg_f32(rows, cols, vals, V, b, rtol=1e-7, maxit=10000, bf16_rows=True, refine=1, theta=3.0):
    """Model of ls_pcg_fused.cuh.  Per iteration:
         phase B  r -= alpha s;  z = rnd(D^-1 r);  gamma' = r.z, rr = r.r        -> reduction 1 (beta, convergence)
         phase A  w = A z;  x += alpha_prev p;  p = z + beta p;  s = w + beta s;  delta = p.s   -> reduction 2 (alpha)
       and at convergence the true residual b - A x (fp64) decides whether to restart.
       Returns (x, iterations, restarts)."""
    A = _csr(rows, cols, vals, V, f32)
    A64, Aabs = A.astype(np.float64), abs(A.astype(np.float64))
    b = np.asarray(b, dtype=f32)
    dinv = (f32(1.0) / A.diagonal().astype(f32))
    d = lambda u, w: np.einsum("ij,ij->j", u.astype(np.float64), w.astype(np.float64))
    rnd = _bf16 if bf16_rows else (lambda t: t)
    bb = d(b, b)
    x = np.zeros_like(b)
    r = b.copy()
    active = bb > 0
    it = restarts = checks = 0
    while True:
        z = rnd((dinv[:, None] * r).astype(f32))
        gam = d(r, z)
        p = np.zeros_like(b)
        s = np.zeros_like(b)
        alpha = np.zeros(b.shape[1], dtype=f32)
        beta = np.zeros(b.shape[1], dtype=f32)
        while active.any() and it < maxit:
            w = (A @ z).astype(f32)                                   # phase A
            x = (x + alpha[None, :] * p).astype(f32)
            p = (z + beta[None, :] * p).astype(f32)
            s = (w + beta[None, :] * s).astype(f32)
            dl = d(p, s)
            alpha = np.where(active & (dl > 0), gam / np.where(dl == 0, 1, dl), 0.0).astype(f32)
            r = (r - alpha[None, :] * s).astype(f32)                   # phase B
            z = rnd((dinv[:, None] * r).astype(f32))
            gam_new, rr = d(r, z), d(r, r)
            it += 1
            conv = rr <= (rtol * rtol) * bb
            beta = np.where(active & ~conv, gam_new / np.where(gam == 0, 1, gam), 0.0).astype(f32)
            gam = gam_new
            active = active & ~conv
        x = (x + alpha[None, :] * p).astype(f32)                      # pending update
        if refine <= 0 or checks > refine or it == 0 or it >= maxit:
            break
        rt = b.astype(np.float64) - A64 @ x.astype(np.float64)
        floor = Aabs @ np.abs(x.astype(np.float64))
        rrt, fl2 = d(rt, rt), d(floor, floor)
        checks += 1
        need = (rrt > (rtol * rtol) * bb) & (rrt > (theta * 2.0 ** -24) ** 2 * fl2) & (bb > 0) & (restarts < refine)
        if not need.any():
            break
        restarts += 1
        r = rt.astype(f32)
        active = need
    return x, it, restarts
