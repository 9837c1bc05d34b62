"""Synthetic mesh workloads for tests and bench (numpy, host side, deterministic).

These generate the *inputs* of the hot path (verts, faces) for the configurations
named in BASELINE.json / SURVEY.md section 8(d).  Nothing here is on the solve path.

  * icosphere(level)           config 1: level 4 -> V=2562, F=5120
  * subdivide(verts, faces)    config 2: bunny (3301 V) subdivided x2 -> V=52786
  * plane(n, seed)             config 3: n=1000 -> V=1e6, F=1996002, nnz(M)=6992002
                               config 4: n=500 (one per GPU, seed = rank)
  * load_obj(path)             minimal OBJ reader (v / f lines, triangles only)
"""
import numpy as np


def icosahedron():
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = np.array([
        [-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0],
        [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
        [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([
        [0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
        [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
        [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
        [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    return v, f


def subdivide(verts, faces):
    """One level of midpoint (1-to-4) subdivision. Returns (verts float64, faces int64)."""
    verts = np.asarray(verts, dtype=np.float64)
    faces = np.asarray(faces, dtype=np.int64)
    V = verts.shape[0]
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], axis=0)
    e = np.sort(e, axis=1)
    key = e[:, 0] * V + e[:, 1]
    ukey, inv = np.unique(key, return_inverse=True)
    mid = 0.5 * (verts[ukey // V] + verts[ukey % V])
    F = faces.shape[0]
    m01 = V + inv[0:F]
    m12 = V + inv[F:2 * F]
    m20 = V + inv[2 * F:3 * F]
    v0, v1, v2 = faces[:, 0], faces[:, 1], faces[:, 2]
    nf = np.concatenate([
        np.stack([v0, m01, m20], 1),
        np.stack([v1, m12, m01], 1),
        np.stack([v2, m20, m12], 1),
        np.stack([m01, m12, m20], 1)], axis=0)
    return np.concatenate([verts, mid], axis=0), nf


def icosphere(level=4):
    """Icosahedron, `level` midpoint subdivisions re-projected to the unit sphere."""
    v, f = icosahedron()
    for _ in range(level):
        v, f = subdivide(v, f)
        v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v.astype(np.float32), f


def plane(n, seed=0):
    """n x n vertex grid on [0,1]^2, z = 0.05 sin(6 pi x) cos(4 pi y) + N(0, 1e-3^2).

    Faces (a,b,d),(a,d,c) per cell with a=(i,j), b=(i,j+1), c=(i+1,j), d=(i+1,j+1);
    vertex id = i*n + j.  V = n^2, F = 2 (n-1)^2, nnz(M) = V + 2E = 7n^2 - 8n + 2... (=6992002 at n=1000).
    """
    lin = np.linspace(0.0, 1.0, n)
    y, x = np.meshgrid(lin, lin, indexing="ij")
    rng = np.random.default_rng(seed)
    z = 0.05 * np.sin(6 * np.pi * x) * np.cos(4 * np.pi * y) + rng.normal(0.0, 1e-3, size=x.shape)
    verts = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1).astype(np.float32)
    i, j = np.meshgrid(np.arange(n - 1), np.arange(n - 1), indexing="ij")
    a = (i * n + j).ravel()
    b = a + 1
    c = a + n
    d = c + 1
    faces = np.concatenate([np.stack([a, b, d], 1), np.stack([a, d, c], 1)], axis=0).astype(np.int64)
    return verts, faces


def load_obj(path):
    vs, fs = [], []
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith("v "):
                vs.append([float(t) for t in line.split()[1:4]])
            elif line.startswith("f "):
                idx = [int(t.split("/")[0]) - 1 for t in line.split()[1:]]
                for k in range(1, len(idx) - 1):
                    fs.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(vs, dtype=np.float64), np.asarray(fs, dtype=np.int64)


def shuffle_vertices(verts, faces, seed=0):
    """Random relabelling of the vertices (worst-case ordering for the gather)."""
    rng = np.random.default_rng(seed)
    V = verts.shape[0]
    perm = rng.permutation(V)           # new id of old vertex i is perm[i]
    nv = np.empty_like(verts)
    nv[perm] = verts
    return nv, perm[faces]


def rhs_recipe(M_apply, verts, seed0=0, seed1=1, seed2=2):
    """SURVEY 8(d) RHS recipe: v = verts + N(0,.01^2); u = M v; b = u + N(0,.01^2); g ~ N(0,1)."""
    V = verts.shape[0]
    v = verts.astype(np.float64) + np.random.default_rng(seed0).normal(0, 0.01, size=(V, 3))
    u = M_apply(v)
    b = u + np.random.default_rng(seed1).normal(0, 0.01, size=(V, 3))
    g = np.random.default_rng(seed2).normal(0, 1.0, size=(V, 3))
    return v.astype(np.float32), b.astype(np.float32), g.astype(np.float32)
