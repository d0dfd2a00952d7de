"""Deterministic synthetic tetrahedral meshes and scenes.

All arrays follow the reference's conventions (sim/vbd/Data.h:167-176): ``X`` is 3 x nV float64,
``T`` is 4 x nT int64, ``F`` is 3 x nF int64 with outward orientation, ``V`` the collision
vertices, ``B`` the per-vertex body ids.
"""
from __future__ import annotations

import numpy as np

# The reference's own test cube (sim/vbd/Integrator.cpp:251-264): local vertex = x + 2y + 4z.
CUBE_P = np.array([[0., 1., 0., 1., 0., 1., 0., 1.],
                   [0., 0., 1., 1., 0., 0., 1., 1.],
                   [0., 0., 0., 0., 1., 1., 1., 1.]])
CUBE_T = np.array([[0, 3, 5, 6, 0],
                   [1, 2, 4, 7, 5],
                   [3, 0, 6, 5, 3],
                   [5, 6, 0, 3, 6]], dtype=np.int64)
CUBE_F = np.array([[0, 1, 1, 3, 3, 2, 2, 0, 0, 0, 4, 5],
                   [1, 5, 3, 7, 2, 6, 0, 4, 3, 2, 5, 7],
                   [4, 4, 5, 5, 7, 7, 6, 6, 1, 3, 6, 6]], dtype=np.int64)

# odd-parity cell: the even split mirrored in x (local index bit 0 flipped); swapping two
# vertices of every tet restores positive orientation.
_ODD_T = (CUBE_T ^ 1)[[1, 0, 2, 3], :]


def tet_grid(nx: int, ny: int, nz: int, h: float = 1.0, origin=(0.0, 0.0, 0.0)):
    """``nx*ny*nz`` cubes of edge ``h``, each split into 5 tets with alternating parity so that
    faces match between neighbours.  Vertex id = (i*(ny+1)+j)*(nz+1)+k.  Returns (X, T)."""
    i, j, k = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    X = np.stack([i.ravel(), j.ravel(), k.ravel()]).astype(np.float64) * h
    X += np.asarray(origin, dtype=np.float64)[:, None]
    ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ci, cj, ck = ci.ravel(), cj.ravel(), ck.ravel()
    vid = lambda a, b, c: (a * (ny + 1) + b) * (nz + 1) + c  # noqa: E731
    corners = np.stack([vid(ci + (l & 1), cj + ((l >> 1) & 1), ck + ((l >> 2) & 1))
                        for l in range(8)])  # 8 x nCells
    odd = ((ci + cj + ck) & 1).astype(bool)
    T = np.empty((4, 5, ci.size), dtype=np.int64)
    for t in range(5):
        for a in range(4):
            T[a, t] = np.where(odd, corners[_ODD_T[a, t], np.arange(ci.size)],
                               corners[CUBE_T[a, t], np.arange(ci.size)])
    T = T.transpose(0, 2, 1).reshape(4, -1)  # cell-major, 5 tets per cell
    return X, np.ascontiguousarray(T)


def boundary_facets(T: np.ndarray):
    """Boundary triangles of a tet mesh, oriented outward (normal points away from the tet).
    Returns F (3 x nF)."""
    # faces opposite each local vertex, outward for a positively oriented tet
    loc = np.array([[1, 2, 3], [0, 3, 2], [0, 1, 3], [0, 2, 1]])
    faces = np.concatenate([T[loc[a]] for a in range(4)], axis=1)  # 3 x 4nT
    key = np.sort(faces, axis=0).astype(np.int64)
    n = int(key.max()) + 1 if key.size else 1
    if n < 2_000_000:       # the sorted triple as one 63-bit key: a 1-D sort instead of a lexicographic one (12 s -> 1 s at 2 M tets)
        key = (key[0] * n + key[1]) * n + key[2]
        _, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    else:
        _, inv, cnt = np.unique(key, axis=1, return_inverse=True, return_counts=True)
    return np.ascontiguousarray(faces[:, cnt[inv.ravel()] == 1])


def tet_volumes(X, T):
    a, b, c, d = (X[:, T[i]] for i in range(4))
    return np.einsum("ij,ij->j", np.cross(b - a, c - a, axis=0), d - a) / 6.0


def stack_bodies(X, T, n: int, axis: int = 2, gap_frac: float = 0.1):
    """n copies of (X,T) stacked along ``axis`` with a gap of ``gap_frac`` x extent, as the
    reference example does (python/examples/vbd.py:237-241).  Returns X, T, B."""
    ext = X[axis].max() - X[axis].min()
    Xs, Ts, Bs = [], [], []
    for b in range(n):
        Xb = X.copy()
        Xb[axis] += b * ext * (1.0 + gap_frac)
        Xs.append(Xb)
        Ts.append(T + b * X.shape[1])
        Bs.append(np.full(X.shape[1], b, dtype=np.int64))
    return np.concatenate(Xs, 1), np.concatenate(Ts, 1), np.concatenate(Bs)


def batch_scenes(X, T, n: int, perturb: float = 0.0):
    """n independent copies of one scene concatenated into a single (disconnected) mesh: the
    throughput mode of SURVEY.md section 8e.  Scene s is perturbed with rng seed s."""
    nV = X.shape[1]
    Xs = np.tile(X, (1, n))
    if perturb > 0:
        for s in range(n):
            Xs[:, s * nV:(s + 1) * nV] += perturb * np.random.default_rng(s).uniform(-1, 1, (3, nV))
    Ts = (T[:, None, :] + (np.arange(n) * nV)[None, :, None]).reshape(4, -1)
    return Xs, np.ascontiguousarray(Ts)
