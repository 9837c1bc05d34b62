"""Oracle: numpy restatement of largesteps/geometry.py (TEST INFRASTRUCTURE, see oracle/__init__.py).

All arithmetic is done in float32 in the same order of operations as the reference so that the
uniform-Laplacian system matrix is reproduced bit-exactly and the cotangent one to fp32 rounding
(the reference's diagonal is a torch.sparse.sum whose summation order is unspecified).

Matrices are returned as coalesced COO triplets (rows int64, cols int64, vals float32), sorted
row-major -- the layout `compute_matrix(...).coalesce()` has (geometry.py:133).
"""
import numpy as np

f32 = np.float32


def _coalesce(rows, cols, vals, V):
    """torch .coalesce(): sort by (row, col), sum duplicates (fp32, in sorted-stable order)."""
    key = rows.astype(np.int64) * V + cols.astype(np.int64)
    order = np.argsort(key, kind="stable")
    key = key[order]
    vals = vals[order].astype(f32)
    ukey, start = np.unique(key, return_index=True)
    out = np.add.reduceat(vals, start).astype(f32) if len(vals) else vals
    return ukey // V, ukey % V, out


def laplacian_uniform(V, faces):
    """geometry.py:65-94.  L = D - A (combinatorial), coalesced."""
    faces = np.asarray(faces, dtype=np.int64)
    ii = faces[:, [1, 2, 0]].ravel()                      # geometry.py:80
    jj = faces[:, [2, 0, 1]].ravel()                      # geometry.py:81
    r = np.concatenate([ii, jj])
    c = np.concatenate([jj, ii])
    key = np.unique(r * V + c)                            # .unique(dim=1), geometry.py:82
    ar, ac = key // V, key % V
    ones = np.ones(len(key), dtype=f32)                   # geometry.py:83
    rows = np.concatenate([ar, ar])                       # geometry.py:86-89 (diag_idx = adj[0])
    cols = np.concatenate([ac, ar])
    vals = np.concatenate([-ones, ones])                  # geometry.py:90
    return _coalesce(rows, cols, vals, V)                 # geometry.py:94


def cot_weights(verts, faces):
    """geometry.py:20-42: per-face cotangents (cota, cotb, cotc)/4 in fp32, shape (F,3)."""
    verts = np.asarray(verts, dtype=f32)
    fv = verts[faces]                                     # geometry.py:20
    v0, v1, v2 = fv[:, 0], fv[:, 1], fv[:, 2]
    nrm = lambda d: np.sqrt((d * d).sum(axis=1, dtype=f32)).astype(f32)
    A = nrm(v1 - v2)                                      # geometry.py:25-27
    B = nrm(v0 - v2)
    C = nrm(v0 - v1)
    s = (f32(0.5) * (A + B + C)).astype(f32)              # geometry.py:30
    area = np.sqrt(np.maximum(s * (s - A) * (s - B) * (s - C), f32(1e-12))).astype(f32)   # geometry.py:33
    A2, B2, C2 = A * A, B * B, C * C                      # geometry.py:36
    cota = (B2 + C2 - A2) / area                          # geometry.py:37-39
    cotb = (A2 + C2 - B2) / area
    cotc = (A2 + B2 - C2) / area
    cot = np.stack([cota, cotb, cotc], axis=1).astype(f32)
    cot /= f32(4.0)                                       # geometry.py:41
    return cot


def laplacian_cot(verts, faces):
    """geometry.py:3-63.  PSD cotangent Laplacian without the 1/2 factor, *coalesced* here
    (the reference leaves it uncoalesced; only compute_matrix:133 coalesces)."""
    rows, cols, vals = _laplacian_cot_uncoalesced(verts, faces)
    return _coalesce(rows, cols, vals, np.asarray(verts).shape[0])


def _laplacian_cot_uncoalesced(verts, faces):
    faces = np.asarray(faces, dtype=np.int64)
    V = np.asarray(verts).shape[0]
    cot = cot_weights(verts, faces)
    ii = faces[:, [1, 2, 0]].ravel()                      # geometry.py:47
    jj = faces[:, [2, 0, 1]].ravel()                      # geometry.py:48
    w = cot.ravel()                                       # L[v1,v2]=cota, L[v2,v0]=cotb, L[v0,v1]=cotc
    # L += L.t()  (geometry.py:56): entries (ii,jj,w) and (jj,ii,w)
    r = np.concatenate([ii, jj])
    c = np.concatenate([jj, ii])
    ww = np.concatenate([w, w]).astype(f32)
    # diagonal = column sums of the symmetric matrix (geometry.py:59), fp32 accumulation
    diag = np.zeros(V, dtype=f32)
    np.add.at(diag, c, ww)
    idx = np.arange(V, dtype=np.int64)                    # geometry.py:60-62: diag - L
    return (np.concatenate([idx, r]), np.concatenate([idx, c]),
            np.concatenate([diag, -ww]).astype(f32))


def compute_matrix(verts, faces, lambda_, alpha=None, cotan=False):
    """geometry.py:96-133.  Returns coalesced (rows, cols, vals, V)."""
    verts = np.asarray(verts)
    V = verts.shape[0]
    if cotan:
        lr, lc, lv = _laplacian_cot_uncoalesced(verts, faces)     # geometry.py:120
    else:
        lr, lc, lv = laplacian_uniform(V, faces)                  # geometry.py:122
    idx = np.arange(V, dtype=np.int64)                            # geometry.py:124-125
    ones = np.ones(V, dtype=f32)
    if alpha is None:
        ev = ones                                                 # geometry.py:128  M = I + lambda L
        sv = (f32(lambda_) * lv).astype(f32)
    else:
        if alpha < 0.0 or alpha >= 1.0:                           # geometry.py:130-131
            raise ValueError(f"Invalid value for alpha: {alpha} : it should take values between 0 (included) and 1 (excluded)")
        ev = (f32(1 - alpha) * ones).astype(f32)                  # geometry.py:132
        sv = (f32(alpha) * lv).astype(f32)
    rows = np.concatenate([idx, lr])
    cols = np.concatenate([idx, lc])
    vals = np.concatenate([ev, sv])
    r, c, v = _coalesce(rows, cols, vals, V)                      # geometry.py:133
    return r, c, v, V


def coo_to_scipy(rows, cols, vals, V, dtype=np.float64):
    import scipy.sparse as sp
    return sp.csr_matrix((np.asarray(vals, dtype=dtype), (rows, cols)), shape=(V, V))
