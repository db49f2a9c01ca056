"""Problem generator for the surface-multigrid hot path (host side, numpy/scipy).

This is *input production* for tests and benchmarks, not part of the GPU product
path: it builds the ``(A, P_l, known, b, z0)`` tuples that the reference's examples
hand to ``min_quad_with_fixed_mg_precompute/_solve``.

Restated reference behaviour (file:line under /root/reference):
  * ``normalize_unit_area``      src/normalize_unit_area.cpp:3-25
  * ``cotmatrix``                libigl/include/igl/cotmatrix.cpp:19-78 (+ cotmatrix_entries.cpp:21-58)
  * ``massmatrix`` (voronoi/barycentric) libigl/include/igl/massmatrix_intrinsic.cpp:30-119
  * ``boundary_loop`` (longest)  libigl/include/igl/boundary_loop.cpp:93-128
  * ``upsample``                 libigl/include/igl/upsample.cpp:17-103
  * Poisson problem of 03/04     03_mg_solver/main.cpp:44-65, 04_mg_solver_nobd/main.cpp:73-94
  * MCF step of 05               05_example_mean_curvature_flow/main.cpp:57-79

The reference's own hierarchy builder (mg_precompute -> SSP decimation + joint
LSCM) needs Eigen and is out of scope; hierarchies here are
midpoint-subdivision hierarchies, whose subdivision matrix ``S`` is a barycentric
prolongation with the structure the hot path relies on (row-stochastic, <= 3
entries per row, coarse-triangle support).
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional, Sequence

import numpy as np
import scipy.sparse as sp


# --------------------------------------------------------------------------- #
# mesh IO / basic geometry
# --------------------------------------------------------------------------- #
def read_obj(path: str):
    """Minimal OBJ reader (v / f lines, triangles, 1-based, ``a/b/c`` tolerated)."""
    vs, fs = [], []
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith("v "):
                p = line.split()
                vs.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith("f "):
                p = line.split()[1:]
                idx = [int(t.split("/")[0]) - 1 for t in p]
                for t in range(1, len(idx) - 1):
                    fs.append((idx[0], idx[t], idx[t + 1]))
    return np.asarray(vs, dtype=np.float64), np.asarray(fs, dtype=np.int32)


def doublearea(V: np.ndarray, F: np.ndarray) -> np.ndarray:
    e1 = V[F[:, 1]] - V[F[:, 0]]
    e2 = V[F[:, 2]] - V[F[:, 0]]
    return np.linalg.norm(np.cross(e1, e2), axis=1)


def normalize_unit_area(V: np.ndarray, F: np.ndarray) -> np.ndarray:
    """src/normalize_unit_area.cpp:3-25 : scale to unit area, recentre x,y, z-min to 0."""
    V = np.array(V, dtype=np.float64, copy=True)
    scale = np.sqrt(doublearea(V, F).sum() / 2.0)
    V /= scale
    V[:, 0] -= V[:, 0].mean()
    V[:, 1] -= V[:, 1].mean()
    V[:, 2] -= V[:, 2].min()
    return V


def edge_lengths(V, F):
    """igl::edge_lengths: column i is the edge opposite corner i."""
    l0 = np.linalg.norm(V[F[:, 1]] - V[F[:, 2]], axis=1)
    l1 = np.linalg.norm(V[F[:, 2]] - V[F[:, 0]], axis=1)
    l2 = np.linalg.norm(V[F[:, 0]] - V[F[:, 1]], axis=1)
    return np.stack([l0, l1, l2], axis=1)


def cotmatrix(V: np.ndarray, F: np.ndarray) -> sp.csc_matrix:
    """Cotangent Laplacian L (negative semi-definite), igl::cotmatrix semantics:
    L(i,j) = 1/2 (cot a_ij + cot b_ij), L(i,i) = -sum_j L(i,j)."""
    n = V.shape[0]
    l = edge_lengths(V, F)
    l2 = l * l
    # Heron, as igl::doublearea(l) does
    s = l.sum(axis=1) * 0.5
    dblA = 2.0 * np.sqrt(np.maximum(s * (s - l[:, 0]) * (s - l[:, 1]) * (s - l[:, 2]), 0.0))
    C = np.empty_like(l)
    # cotmatrix_entries.cpp:41-55 : C(:,i) = cot(angle at corner i) / 2
    C[:, 0] = (l2[:, 1] + l2[:, 2] - l2[:, 0]) / dblA / 4.0
    C[:, 1] = (l2[:, 2] + l2[:, 0] - l2[:, 1]) / dblA / 4.0
    C[:, 2] = (l2[:, 0] + l2[:, 1] - l2[:, 2]) / dblA / 4.0
    # edge opposite corner i is (i+1, i+2)
    I, J, X = [], [], []
    for i in range(3):
        a = F[:, (i + 1) % 3]
        b = F[:, (i + 2) % 3]
        c = C[:, i]
        I += [a, b, a, b]
        J += [b, a, a, b]
        X += [c, c, -c, -c]
    L = sp.coo_matrix(
        (np.concatenate(X), (np.concatenate(I).astype(np.int64), np.concatenate(J).astype(np.int64))),
        shape=(n, n),
    ).tocsc()
    L.sum_duplicates()
    L.sort_indices()
    return L


def massmatrix_diag(V: np.ndarray, F: np.ndarray, kind: str = "voronoi") -> np.ndarray:
    """Diagonal of igl::massmatrix (massmatrix_intrinsic.cpp:50-112)."""
    n = V.shape[0]
    l = edge_lengths(V, F)
    s = l.sum(axis=1) * 0.5
    dblA = 2.0 * np.sqrt(np.maximum(s * (s - l[:, 0]) * (s - l[:, 1]) * (s - l[:, 2]), 0.0))
    if kind == "barycentric":
        q = np.repeat((dblA / 6.0)[:, None], 3, axis=1)
    elif kind == "voronoi":
        l0, l1, l2 = l[:, 0], l[:, 1], l[:, 2]
        cos = np.stack(
            [
                (l2**2 + l1**2 - l0**2) / (l1 * l2 * 2.0),
                (l0**2 + l2**2 - l1**2) / (l2 * l0 * 2.0),
                (l1**2 + l0**2 - l2**2) / (l0 * l1 * 2.0),
            ],
            axis=1,
        )
        bary = cos * l
        bary = bary / bary.sum(axis=1, keepdims=True)
        partial = bary * (dblA * 0.5)[:, None]
        q = np.stack(
            [
                (partial[:, 1] + partial[:, 2]) * 0.5,
                (partial[:, 2] + partial[:, 0]) * 0.5,
                (partial[:, 0] + partial[:, 1]) * 0.5,
            ],
            axis=1,
        )
        for c in range(3):
            ob = cos[:, c] < 0
            for d in range(3):
                q[:, d] = np.where(ob, (0.25 if d == c else 0.125) * dblA, q[:, d])
    else:
        raise ValueError(kind)
    m = np.zeros(n)
    for c in range(3):
        np.add.at(m, F[:, c], q[:, c])
    return m


def boundary_loops(F: np.ndarray) -> List[np.ndarray]:
    """All boundary loops (vertex index sequences following boundary half-edges)."""
    he = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], axis=0).astype(np.int64)
    n = int(F.max()) + 1
    key = he[:, 0] * n + he[:, 1]
    rkey = he[:, 1] * n + he[:, 0]
    is_b = ~np.isin(key, rkey)
    nxt = {int(a): int(b) for a, b in he[is_b]}
    loops, seen = [], set()
    for start in sorted(nxt):
        if start in seen:
            continue
        loop, v = [], start
        while v not in seen:
            seen.add(v)
            loop.append(v)
            v = nxt[v]
        loops.append(np.asarray(loop, dtype=np.int32))
    return loops


def boundary_loop(F: np.ndarray) -> np.ndarray:
    """igl::boundary_loop(F,b): the longest loop (boundary_loop.cpp:93-128)."""
    loops = boundary_loops(F)
    if not loops:
        return np.zeros(0, dtype=np.int32)
    return max(loops, key=len)


# --------------------------------------------------------------------------- #
# midpoint subdivision (igl::upsample)
# --------------------------------------------------------------------------- #
def upsample(n_verts: int, F: np.ndarray):
    """igl::upsample(n_verts, F, S, NF) (upsample.cpp:17-103).

    New vertex ``n_verts + e`` sits on undirected edge ``e``; edges are numbered in
    order of first appearance while walking faces row-major over (face, corner j)
    with edge j = (F[i,j], F[i,(j+1)%3]).  Returns (S, NF): S is
    (n_verts+#E) x n_verts CSC with rows [I; 1/2 1/2].
    """
    F = np.asarray(F, dtype=np.int64)
    m = F.shape[0]
    a = F.reshape(-1)  # (i,j) -> F[i,j]
    b = np.roll(F, -1, axis=1).reshape(-1)  # F[i,(j+1)%3]
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    key = lo * n_verts + hi
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")  # rank edges by first appearance
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    NI = rank[inv].reshape(m, 3)
    ne = uniq.size
    e_lo = (uniq // n_verts)[order]
    e_hi = (uniq % n_verts)[order]
    # the first half-edge that created edge e decides (F(i,j), F(i,j+1)) order of the
    # two triplets; values are equal so only the pattern matters.
    rows = np.concatenate([np.arange(n_verts), n_verts + np.arange(ne), n_verts + np.arange(ne)])
    cols = np.concatenate([np.arange(n_verts), e_lo, e_hi])
    vals = np.concatenate([np.ones(n_verts), np.full(ne, 0.5), np.full(ne, 0.5)])
    S = sp.coo_matrix((vals, (rows, cols)), shape=(n_verts + ne, n_verts)).tocsc()
    S.sort_indices()
    e = NI + n_verts
    NF = np.empty((m * 4, 3), dtype=np.int64)
    NF[0::4] = np.stack([F[:, 0], e[:, 0], e[:, 2]], axis=1)
    NF[1::4] = np.stack([F[:, 1], e[:, 1], e[:, 0]], axis=1)
    NF[2::4] = np.stack([e[:, 0], e[:, 1], e[:, 2]], axis=1)
    NF[3::4] = np.stack([e[:, 1], F[:, 2], e[:, 2]], axis=1)
    return S, NF.astype(np.int32)


def pad_prolongation_to_three(S: sp.csc_matrix, F_coarse: np.ndarray) -> sp.csc_matrix:
    """Give every row of a subdivision matrix exactly three *stored* entries
    (explicit zeros on the remaining corners of a coarse triangle containing the
    point) - the storage layout get_prolong produces (src/get_prolong.cpp:45-56)."""
    S = S.tocsr()
    S.sort_indices()
    nf, nc = S.shape
    F = np.asarray(F_coarse, dtype=np.int64)
    # one incident coarse face per coarse vertex and per coarse edge
    face_of_vertex = np.full(nc, -1, dtype=np.int64)
    for c in (2, 1, 0):
        face_of_vertex[F[:, c]] = np.arange(F.shape[0])
    a = F.reshape(-1)
    b = np.roll(F, -1, axis=1).reshape(-1)
    ekey = np.minimum(a, b) * nc + np.maximum(a, b)
    eface = np.repeat(np.arange(F.shape[0]), 3)
    order = np.argsort(ekey, kind="stable")
    ekey_s, eface_s = ekey[order], eface[order]
    rows, cols, vals = [], [], []
    indptr, indices, data = S.indptr.astype(np.int64), S.indices.astype(np.int64), S.data
    cnt = np.diff(indptr)
    # rows with one entry (coarse vertex copies)
    r1 = np.nonzero(cnt == 1)[0]
    v = indices[indptr[r1]]
    f = F[face_of_vertex[v]]
    for c in range(3):
        rows.append(r1)
        cols.append(f[:, c])
        vals.append(np.where(f[:, c] == v, data[indptr[r1]], 0.0))
    # rows with two entries (edge midpoints)
    r2 = np.nonzero(cnt == 2)[0]
    v0, v1 = indices[indptr[r2]], indices[indptr[r2] + 1]
    k = np.minimum(v0, v1) * nc + np.maximum(v0, v1)
    pos = np.searchsorted(ekey_s, k)
    f = F[eface_s[pos]]
    w0, w1 = data[indptr[r2]], data[indptr[r2] + 1]
    for c in range(3):
        rows.append(r2)
        cols.append(f[:, c])
        vals.append(np.where(f[:, c] == v0, w0, np.where(f[:, c] == v1, w1, 0.0)))
    r3 = np.nonzero(cnt >= 3)[0]
    for r in r3:
        for p in range(indptr[r], indptr[r + 1]):
            rows.append(np.array([r]))
            cols.append(np.array([indices[p]]))
            vals.append(np.array([data[p]]))
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    vals = np.concatenate(vals)
    return csc_keep_zeros(rows, cols, vals, (nf, nc))


def csc_keep_zeros(rows, cols, vals, shape) -> sp.csc_matrix:
    """CSC from unique triplets, explicit zeros kept, row indices sorted."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    vals = np.asarray(vals, dtype=np.float64)
    order = np.lexsort((rows, cols))
    rows, cols, vals = rows[order], cols[order], vals[order]
    indptr = np.zeros(shape[1] + 1, dtype=np.int32)
    np.add.at(indptr, cols + 1, 1)
    indptr = np.cumsum(indptr).astype(np.int32)
    m = sp.csc_matrix(shape, dtype=np.float64)
    m.indptr, m.indices, m.data = indptr, rows.astype(np.int32), vals
    m.has_sorted_indices = True
    return m


# --------------------------------------------------------------------------- #
# problems
# --------------------------------------------------------------------------- #
@dataclasses.dataclass
class Problem:
    """Everything the reference's examples pass to the solver API."""

    name: str
    A: sp.csc_matrix  # n x n symmetric (e.g. -cotmatrix)
    P: List[sp.csc_matrix]  # P[l-1] = mg[l].P_full : n_{l-1} x n_l, l = 1..nlev-1
    known: Optional[np.ndarray]  # int32 or None (free variant)
    known_val: Optional[np.ndarray]  # nknown x k
    rhs: np.ndarray  # n x k  (col-major when k > 1)
    z0: np.ndarray  # n x k
    tol: float = 1e-3
    max_iter: int = 20
    V: Optional[np.ndarray] = None
    F: Optional[np.ndarray] = None

    @property
    def n(self):
        return self.A.shape[0]

    @property
    def k(self):
        return 1 if self.rhs.ndim == 1 else self.rhs.shape[1]

    @property
    def nlev(self):
        return len(self.P) + 1


def octahedron():
    V = np.array(
        [[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=np.float64
    )
    F = np.array(
        [[0, 2, 4], [2, 1, 4], [1, 3, 4], [3, 0, 4], [2, 0, 5], [1, 2, 5], [3, 1, 5], [0, 3, 5]],
        dtype=np.int32,
    )
    return V, F


def subdivision_hierarchy(V0, F0, n_sub: int, n_levels: int, project_sphere: bool = False,
                          pad_three: bool = False):
    """Subdivide (V0,F0) ``n_sub`` times; keep the last ``n_levels`` meshes as the
    multigrid levels. Returns (V_fine, F_fine, [P_1..P_{n_levels-1}]) with P_l the
    subdivision matrix from level l to level l-1 (fine index first)."""
    assert n_levels >= 1 and n_levels <= n_sub + 1
    meshes = [(np.asarray(V0, dtype=np.float64), np.asarray(F0, dtype=np.int32))]
    Ss = []
    for _ in range(n_sub):
        V, F = meshes[-1]
        S, NF = upsample(V.shape[0], F)
        NV = S @ V
        if project_sphere:
            NV = NV / np.linalg.norm(NV, axis=1, keepdims=True)
        Sx = pad_prolongation_to_three(S, F) if pad_three else S
        Ss.append(Sx)
        meshes.append((NV, NF))
    V, F = meshes[-1]
    P = [Ss[n_sub - l] for l in range(1, n_levels)]  # fine -> coarse order
    for l, p in enumerate(P):
        p.sort_indices()
        P[l] = sp.csc_matrix(p)
        P[l].indices = P[l].indices.astype(np.int32)
        P[l].indptr = P[l].indptr.astype(np.int32)
    return V, F, P


def poisson_problem(name, V, F, P, known, tol=1e-3, max_iter=20, z0=None, known_val=None):
    """The toy Poisson problem of 03_mg_solver/main.cpp:44-65 (and 04):
    A = -cotmatrix, B = M_voronoi * 1, B(known) = known_val, z0 = 0."""
    A = (-cotmatrix(V, F)).tocsc()
    A.sort_indices()
    A.indices = A.indices.astype(np.int32)
    A.indptr = A.indptr.astype(np.int32)
    B = massmatrix_diag(V, F, "voronoi")
    if known is not None:
        known = np.asarray(known, dtype=np.int32)
        if known_val is None:
            known_val = np.zeros(known.shape[0])
        B = B.copy()
        B[known] = known_val if np.ndim(known_val) == 1 else known_val[:, 0]
    if z0 is None:
        z0 = np.zeros(V.shape[0])
    return Problem(name, A, P, known, known_val, B, z0, tol, max_iter, V, F)


def sphere_problem(n_sub: int, n_levels: int, tol: float = 1e-10, max_iter: int = 20,
                   pad_three: bool = False, random_z0: bool = False) -> Problem:
    """BASELINE config 3 family: octahedron subdivided ``n_sub`` times,
    projected to the sphere, unit area; Poisson with the 6 octahedron vertices pinned
    to 0 (they keep indices 0..5 at every level). n_sub=9, n_levels=5 is '1M'."""
    V0, F0 = octahedron()
    V, F, P = subdivision_hierarchy(V0, F0, n_sub, n_levels, project_sphere=True,
                                    pad_three=pad_three)
    V = normalize_unit_area(V, F)
    z0 = None
    if random_z0:
        z0 = np.random.default_rng(0).uniform(-1.0, 1.0, V.shape[0])
    known = np.arange(6, dtype=np.int32)
    if z0 is not None:
        pass
    return poisson_problem(f"sphere_s{n_sub}_l{n_levels}", V, F, P, known, tol, max_iter, z0)


def mesh_subdivided_problem(name, V0, F0, n_sub, n_levels, known_fn=None, tol=1e-10,
                            max_iter=20, pad_three=False) -> Problem:
    """Poisson problem on an input mesh subdivided ``n_sub`` times (flat midpoint
    subdivision). ``known_fn(V,F) -> known`` picks constraints on the fine mesh;
    default: longest boundary loop if any, else vertex 0."""
    V, F, P = subdivision_hierarchy(V0, F0, n_sub, n_levels, pad_three=pad_three)
    V = normalize_unit_area(V, F)
    if known_fn is not None:
        known = known_fn(V, F)
    else:
        known = boundary_loop(F)
        if known.size == 0:
            known = np.array([0], dtype=np.int32)
    return poisson_problem(name, V, F, P, known, tol, max_iter)


def upsampled_mesh_problem(name: str, V0, F0, known, P_coarse, n_sub: int, tol: float = 1e-10,
                           max_iter: int = 40, pad_three: bool = True) -> Problem:
    """04_mg_solver_nobd/main.cpp on an input mesh upsampled ``n_sub`` times (BASELINE config 5:
    hilbert_cube.obj x 3 = 4 028 672 vertices): unit-area normalisation of the fine mesh,
    A = -cotmatrix, ``known`` (indices on the INPUT mesh; igl::upsample keeps them) pinned to 0,
    b = voronoi mass, z0 = 0.  Hierarchy = the ``n_sub`` subdivision prolongations followed by
    the given prolongations below the input mesh (``P_coarse``, fine -> coarse order)."""
    V, F, P = subdivision_hierarchy(V0, F0, n_sub, n_sub + 1, pad_three=pad_three)
    V = normalize_unit_area(V, F)
    Pall = list(P)
    for p in P_coarse:
        p = sp.csc_matrix(p)
        p.indices = p.indices.astype(np.int32)
        p.indptr = p.indptr.astype(np.int32)
        Pall.append(p)
    return poisson_problem(name, V, F, Pall, np.asarray(known, dtype=np.int32), tol, max_iter)


def load_csc_keep_zeros(d, key: str) -> sp.csc_matrix:
    """CSC matrix stored as <key>_{indptr,indices,data,shape} arrays (explicit zeros kept)."""
    shp = tuple(int(x) for x in d[f"{key}_shape"])
    m = sp.csc_matrix(shp, dtype=np.float64)
    m.indptr, m.indices, m.data = d[f"{key}_indptr"], d[f"{key}_indices"], d[f"{key}_data"]
    return m


def mcf_step_problem(V, F, P, U=None, delta: float = 0.01, tol: float = 5e-7,
                     max_iter: int = 20, L0: Optional[sp.csc_matrix] = None) -> Problem:
    """One mean-curvature-flow step of 05_example_mean_curvature_flow/main.cpp:57-79:
    LHS = M(U) - delta*L0 (free variant, SPD), RHS = M(U)*U (n x 3), z0 = U."""
    if U is None:
        U = V
    if L0 is None:
        L0 = cotmatrix(V, F)
    m = massmatrix_diag(U, F, "barycentric")
    A = (sp.diags(m) - delta * L0).tocsc()
    A.sort_indices()
    A.indices = A.indices.astype(np.int32)
    A.indptr = A.indptr.astype(np.int32)
    rhs = np.asfortranarray(m[:, None] * U)
    z0 = np.asfortranarray(U.copy())
    return Problem("mcf_step", A, P, None, None, rhs, z0, tol, max_iter, V, F)


def block_prolongation(P: sp.csc_matrix) -> sp.csc_matrix:
    """get_prolong_block (src/get_prolong.cpp:59-115): the scalar prolongation applied to each of
    the three coordinates of a vertex, interleaved xyz: P_block(3r+d, 3c+d) = P(r, c).  Explicit
    zeros of P stay stored entries (setFromTriplets keeps them)."""
    P = sp.csc_matrix(P)
    rows, cols = [], []
    vals = []
    coo_c = np.repeat(np.arange(P.shape[1]), np.diff(P.indptr))
    for d in range(3):
        rows.append(3 * P.indices.astype(np.int64) + d)
        cols.append(3 * coo_c.astype(np.int64) + d)
        vals.append(P.data)
    return csc_keep_zeros(np.concatenate(rows), np.concatenate(cols), np.concatenate(vals),
                          (3 * P.shape[0], 3 * P.shape[1]))


def balloon_step_problem(V, F, P, dt: float = 0.02, stiffness: float = 50.0, seed: int = 0, tol: float = 1e-8,
                         max_iter: int = 20) -> Problem:
    """The linear system of one Newton step of the balloon example
    (06_example_balloon_sim/sim_utils/implicit_euler_mg_balloon.h:62-76): H = M + dt^2 K on the
    3n interleaved-xyz degrees of freedom, free variant (no fixed values), one right-hand side,
    zero initial guess, with the block hierarchy of mg_precompute_block.  K stands in for the
    libshell membrane Hessian (physics out of scope): an edge-spring Hessian with the same
    structure -- one dense SPD 3x3 block per mesh edge and vertex,
    K_ij = -w_ij (a I + b d d^T), K_ii = -sum_j K_ij, d = unit edge direction -- so H is SPD with
    the sparsity pattern of cotmatrix (x) ones(3, 3)."""
    rng = np.random.default_rng(seed)
    n = V.shape[0]
    ij = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], axis=0).astype(np.int64)
    ij = np.unique(np.sort(ij, axis=1), axis=0)
    d = V[ij[:, 1]] - V[ij[:, 0]]
    length = np.linalg.norm(d, axis=1)
    d = d / length[:, None]
    w = stiffness * (1.0 + 0.3 * rng.random(ij.shape[0])) / np.maximum(length, 1e-12)
    blocks = w[:, None, None] * (0.2 * np.eye(3)[None] + d[:, :, None] * d[:, None, :])  # SPD 3x3 per edge
    rows, cols, vals = [], [], []
    for a in range(3):
        for b in range(3):
            i3, j3 = 3 * ij[:, 0] + a, 3 * ij[:, 1] + b
            i3b, j3b = 3 * ij[:, 0] + b, 3 * ij[:, 1] + a
            rows += [i3, j3b, 3 * ij[:, 0] + a, 3 * ij[:, 1] + a]
            cols += [j3, i3b, 3 * ij[:, 0] + b, 3 * ij[:, 1] + b]
            vals += [-blocks[:, a, b], -blocks[:, a, b], blocks[:, a, b], blocks[:, a, b]]
    K = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(3 * n, 3 * n))
    m = np.repeat(massmatrix_diag(V, F, "barycentric"), 3)
    H = (sp.diags(m) + dt * dt * K).tocsc()
    H.sum_duplicates()
    H.sort_indices()
    H.indices = H.indices.astype(np.int32)
    H.indptr = H.indptr.astype(np.int32)
    Pb = []
    for p in P:
        q = block_prolongation(p)
        q.indices = q.indices.astype(np.int32)
        q.indptr = q.indptr.astype(np.int32)
        Pb.append(q)
    g = rng.standard_normal(3 * n) * np.repeat(massmatrix_diag(V, F, "barycentric"), 3)  # -(M dq + dt G + dt f)
    return Problem("balloon_step", H, Pb, None, None, g, np.zeros(3 * n), tol, max_iter, V, F)


def grid_mesh(nx: int, ny: int, jitter: float = 0.0, seed: int = 0):
    """Small open (boundary-carrying) triangulated grid for unit tests."""
    xs, ys = np.meshgrid(np.arange(nx, dtype=np.float64), np.arange(ny, dtype=np.float64))
    V = np.stack([xs.reshape(-1), ys.reshape(-1), np.zeros(nx * ny)], axis=1)
    if jitter:
        rng = np.random.default_rng(seed)
        interior = (V[:, 0] > 0) & (V[:, 0] < nx - 1) & (V[:, 1] > 0) & (V[:, 1] < ny - 1)
        V[interior, :2] += rng.uniform(-jitter, jitter, (int(interior.sum()), 2))
        V[:, 2] = 0.1 * np.sin(V[:, 0]) * np.cos(V[:, 1])
    fs = []
    for j in range(ny - 1):
        for i in range(nx - 1):
            a = j * nx + i
            fs.append((a, a + 1, a + nx + 1))
            fs.append((a, a + nx + 1, a + nx))
    return V, np.asarray(fs, dtype=np.int32)


# --------------------------------------------------------------------------- #
# stand-in hierarchy for arbitrary meshes (NOT the reference's SSP decimation)
# --------------------------------------------------------------------------- #
def mis_hierarchy(V: np.ndarray, F: np.ndarray, n_levels: int, pad_three: bool = True):
    """Prolongations for an arbitrary triangle mesh without the reference's hierarchy
    builder (mg_precompute needs Eigen and stays out of scope).

    Coarse vertices = a greedy maximal independent set of the level's graph; a fine vertex
    interpolates from its (up to three) nearest coarse neighbours with inverse-distance
    weights.  The result has the layout the hot path relies on (get_prolong.cpp:45-56):
    non-negative rows summing to one, at most -- with ``pad_three`` exactly, where three
    coarse vertices are in reach -- three stored entries per row, explicit zeros kept.
    The coarse graph is the pattern of P^T G P.  Returns [P_1, ..., P_{n_levels-1}].
    """
    n = V.shape[0]
    ij = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], axis=0).astype(np.int64)
    G = sp.coo_matrix((np.ones(ij.shape[0]), (ij[:, 0], ij[:, 1])), shape=(n, n)).tocsr()
    G = ((G + G.T) > 0).astype(np.float64).tocsr()
    X = np.asarray(V, dtype=np.float64)
    Ps = []
    for _ in range(n_levels - 1):
        nf = G.shape[0]
        indptr, indices = G.indptr, G.indices
        state = np.zeros(nf, dtype=np.int8)  # 0 undecided, 1 coarse, 2 fine
        for i in np.argsort(-np.diff(indptr), kind="stable"):  # high degree first
            if state[i] == 0:
                state[i] = 1
                nb = indices[indptr[i]:indptr[i + 1]]
                state[nb[state[nb] == 0]] = 2
        coarse = np.nonzero(state == 1)[0]
        cid = np.full(nf, -1, dtype=np.int64)
        cid[coarse] = np.arange(coarse.size)
        rows, cols, vals = [], [], []
        agg = np.zeros(nf, dtype=np.int64)  # nearest coarse vertex: the aggregate of i
        for i in range(nf):
            if state[i] == 1:
                cand = [(0.0, i)]
            else:
                cand = []
            nb = indices[indptr[i]:indptr[i + 1]]
            seen = {int(i)}
            for j in nb:  # coarse vertices in the 1-ring and 2-ring
                for jj in [j] + list(indices[indptr[j]:indptr[j + 1]]):
                    jj = int(jj)
                    if state[jj] == 1 and jj not in seen:
                        seen.add(jj)
                        cand.append((float(np.linalg.norm(X[i] - X[jj])), jj))
            cand.sort()
            cand = cand[:3]
            agg[i] = cid[cand[0][1]]
            if state[i] == 1:
                w = [1.0] + [0.0] * (len(cand) - 1)
            else:
                inv = np.array([1.0 / max(d, 1e-300) for d, _ in cand])
                w = list(inv / inv.sum())
            if not pad_three:
                keep = [t for t in range(len(cand)) if w[t] != 0.0]
                cand, w = [cand[t] for t in keep], [w[t] for t in keep]
            for (d, j), wt in zip(cand, w):
                rows.append(i)
                cols.append(cid[j])
                vals.append(wt)
        P = csc_keep_zeros(rows, cols, vals, (nf, coarse.size))
        Ps.append(P)
        # coarse graph: two aggregates are adjacent when a fine edge joins them
        Pb = sp.csr_matrix((np.ones(nf), (np.arange(nf), agg)), shape=(nf, coarse.size))
        Gc = (Pb.T @ G @ Pb).tocsr()
        Gc.setdiag(0)
        Gc.eliminate_zeros()
        G = (Gc > 0).astype(np.float64).tocsr()
        X = X[coarse]
    return Ps


def mesh_problem(name: str, V, F, n_levels: int, tol=1e-3, max_iter=20, pad_three=True) -> Problem:
    """03_mg_solver/main.cpp on an arbitrary mesh: unit-area normalisation, A = -cotmatrix,
    longest boundary loop pinned to 0 (vertex 0 if closed), b = voronoi mass, z0 = 0, with
    the stand-in hierarchy of ``mis_hierarchy``."""
    V = normalize_unit_area(V, F)
    known = boundary_loop(F)
    if known.size == 0:
        known = np.array([0], dtype=np.int32)
    P = mis_hierarchy(V, F, n_levels, pad_three=pad_three)
    return poisson_problem(name, V, F, P, known, tol, max_iter)


def write_problem_file(path: str, pr: Problem):
    """Flat binary hand-over of a (k = 1, fixed-variant) problem to the headless C++
    examples: int32 header {magic 'SMG1', levels, n, nknown}; A then every P_full as
    {rows, cols, nnz, colptr[cols+1], rowidx[nnz], val[nnz]}; known[nknown] (int32),
    known_val, rhs, z0 (float64)."""
    assert pr.k == 1 and pr.known is not None
    with open(path, "wb") as fh:
        np.array([0x534D4731, pr.nlev, pr.n, pr.known.size], dtype=np.int32).tofile(fh)
        for m in [pr.A] + list(pr.P):
            m = m.tocsc()
            np.array([m.shape[0], m.shape[1], m.nnz], dtype=np.int32).tofile(fh)
            m.indptr.astype(np.int32).tofile(fh)
            m.indices.astype(np.int32).tofile(fh)
            m.data.astype(np.float64).tofile(fh)
        pr.known.astype(np.int32).tofile(fh)
        np.asarray(pr.known_val, dtype=np.float64).tofile(fh)
        np.asarray(pr.rhs, dtype=np.float64).tofile(fh)
        np.asarray(pr.z0, dtype=np.float64).tofile(fh)
