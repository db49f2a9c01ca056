"""Problem-generator sanity checks, using the reference tree's adjacent known-answer
tests as the spec (libigl/tests/include/igl/upsample.cpp:6-28, cotmatrix.cpp:31-128)."""
import numpy as np

from surface_multigrid_code_b200 import meshgen as mg


def test_upsample_single_triangle_known_answer():
    S, NF = mg.upsample(3, np.array([[0, 1, 2]]))
    assert np.array_equal(NF, [[0, 3, 5], [1, 4, 3], [3, 4, 5], [4, 2, 5]])
    assert np.array_equal(S.toarray(), [[1, 0, 0], [0, 1, 0], [0, 0, 1], [.5, .5, 0], [0, .5, .5], [.5, 0, .5]])


def test_cotmatrix_invariants():
    V, F = mg.octahedron()
    Vs, Fs, _ = mg.subdivision_hierarchy(V, F, 3, 1, project_sphere=True)
    L = mg.cotmatrix(Vs, Fs)
    assert abs(L @ np.ones(Vs.shape[0])).max() < 1e-12
    assert abs(L - L.T).max() < 1e-14
    L2 = mg.cotmatrix(Vs * 1e8, Fs)  # scale invariance
    assert abs(L - L2).max() < 1e-6
    m = mg.massmatrix_diag(Vs, Fs, "voronoi")
    assert abs(m.sum() - mg.doublearea(Vs, Fs).sum() / 2) < 1e-12


def test_padded_prolongation_layout():
    """get_prolong.cpp:45-56: exactly three stored entries per row, non-negative,
    rows sum to one, explicit zeros kept."""
    pr = mg.sphere_problem(8, 2, pad_three=True)  # 262146 vertices: int64-sized edge keys
    P = pr.P[0].tocsr()
    assert np.all(np.diff(P.indptr) == 3)
    assert P.data.min() == 0.0 and np.all(np.asarray(P.sum(axis=1)).ravel() == 1.0)
    assert P.shape == (262146, 65538)


def test_boundary_loop_and_normalisation():
    V, F = mg.grid_mesh(5, 4)
    b = mg.boundary_loop(F)
    assert len(b) == 2 * (5 + 4) - 4
    Vn = mg.normalize_unit_area(V, F)
    assert abs(mg.doublearea(Vn, F).sum() / 2 - 1.0) < 1e-12
