"""CPU test of the arithmetic of the device-side mean-curvature-flow assembly
(surface_multigrid_code_b200/csrc/mcf_core.hpp, the code the CUDA kernels run): compiled for
the host (tests/native/mcf_host.cpp) and compared with a numpy restatement of what
05_example_mean_curvature_flow/main.cpp:66-69 computes through libigl
(squared_edge_lengths.cpp:39-41, doublearea.cpp Kahan/Heron from sorted lengths,
massmatrix_intrinsic.cpp:56-66 barycentric, setFromTriplets summation order)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from surface_multigrid_code_b200 import meshgen as mg

HERE = os.path.dirname(os.path.abspath(__file__))


def igl_barycentric_mass(U, F):
    """numpy restatement, same operation order as libigl"""
    def length(p, q):
        d = U[p] - U[q]
        return np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])

    l = np.stack([length(F[:, 1], F[:, 2]), length(F[:, 2], F[:, 0]), length(F[:, 0], F[:, 1])], axis=1)
    l = -np.sort(-l, axis=1)  # descending
    l0, l1, l2 = l[:, 0], l[:, 1], l[:, 2]
    arg = (l0 + (l1 + l2)) * (l2 - (l0 - l1)) * (l2 + (l0 - l1)) * (l0 + (l1 - l2))
    with np.errstate(invalid="ignore"):
        dblA = 2.0 * 0.25 * np.sqrt(arg)
    dblA = np.where(np.isnan(dblA), 0.0, dblA)
    mv = dblA / 6.0
    m = np.zeros(U.shape[0])
    for c in range(3):  # setFromTriplets: corner 0 of every face, then corner 1, then corner 2
        np.add.at(m, F[:, c], mv)
    return dblA, m


@pytest.fixture(scope="module")
def host_lib():
    so = os.path.join(HERE, "native", "libmcf_host.so")
    src = os.path.join(HERE, "native", "mcf_host.cpp")
    hdr = os.path.join(os.path.dirname(HERE), "surface_multigrid_code_b200", "csrc", "mcf_core.hpp")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src])
    L = C.CDLL(so)
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.mcf_host_assemble.argtypes = [C.c_int, C.c_int, ip, dp, ip, ip, dp, dp]
    L.mcf_host_lhs_entry.restype = C.c_double
    L.mcf_host_lhs_entry.argtypes = [C.c_double, C.c_double, C.c_double]
    return L


def _vertex_faces(F, nV):
    """incident faces per vertex in assembly order (corner-major, faces ascending)"""
    nF = F.shape[0]
    flat = np.concatenate([F[:, 0], F[:, 1], F[:, 2]])
    faces = np.concatenate([np.arange(nF)] * 3)
    order = np.argsort(flat, kind="stable")
    ptr = np.zeros(nV + 1, dtype=np.int32)
    np.add.at(ptr, flat + 1, 1)
    return np.cumsum(ptr).astype(np.int32), faces[order].astype(np.int32)


@pytest.mark.parametrize("noise", [0.0, 0.05])
def test_mass_and_system_match_the_libigl_restatement(host_lib, noise):
    V0, F0 = mg.octahedron()
    Vs, Fs, P = mg.subdivision_hierarchy(V0, F0, 4, 3, project_sphere=True, pad_three=True)
    Vs = mg.normalize_unit_area(Vs, Fs)
    rng = np.random.default_rng(4)
    U = np.asfortranarray(Vs * (1.0 + noise * rng.standard_normal((Vs.shape[0], 1))))
    F = np.ascontiguousarray(Fs, dtype=np.int32)
    nV, nF = U.shape[0], F.shape[0]
    ptr, faces = _vertex_faces(F, nV)
    Fcm = np.ascontiguousarray(F.T).reshape(-1)
    Ucm = np.ascontiguousarray(U.T).reshape(-1)
    dblA, mass = np.empty(nF), np.empty(nV)
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    host_lib.mcf_host_assemble(nV, nF, Fcm.ctypes.data_as(ip), Ucm.ctypes.data_as(dp), ptr.ctypes.data_as(ip),
                               faces.ctypes.data_as(ip), dblA.ctypes.data_as(dp), mass.ctypes.data_as(dp))
    dblA_ref, mass_ref = igl_barycentric_mass(np.asarray(U), F)
    assert np.array_equal(dblA, dblA_ref)
    assert np.array_equal(mass, mass_ref)
    # the textbook formulas agree to rounding (cross product area, plain Heron)
    assert np.allclose(dblA, mg.doublearea(np.asarray(U), F), rtol=1e-12)
    assert np.allclose(mass, mg.massmatrix_diag(np.asarray(U), F, "barycentric"), rtol=1e-12)
    assert abs(mass.sum() - 0.5 * dblA.sum()) <= 1e-12 * mass.sum()  # the masses partition the area
    # LHS = M - delta * L entry by entry, as scipy (and Eigen) evaluate it
    delta = 0.01
    L0 = mg.cotmatrix(Vs, Fs).tocsc()
    L0.sort_indices()
    A = (sp.diags(mass_ref) - delta * L0).tocsc()
    A.sort_indices()
    assert np.array_equal(A.indices, L0.indices) and np.array_equal(A.indptr, L0.indptr)
    cols = np.repeat(np.arange(nV), np.diff(L0.indptr))
    got = np.array([host_lib.mcf_host_lhs_entry(mass[c] if r == c else 0.0, delta, l)
                    for r, c, l in zip(L0.indices, cols, L0.data)])
    assert np.array_equal(got, A.data)


def test_degenerate_triangle_gives_zero_area(host_lib):
    U = np.asfortranarray(np.array([[0.0, 0, 0], [1, 0, 0], [2, 0, 0], [0, 1, 0]]))
    F = np.array([[0, 1, 2], [0, 1, 3]], dtype=np.int32)  # first face is collinear
    ptr, faces = _vertex_faces(F, 4)
    dblA, mass = np.empty(2), np.empty(4)
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    Fcm, Ucm = np.ascontiguousarray(F.T).reshape(-1), np.ascontiguousarray(U.T).reshape(-1)
    host_lib.mcf_host_assemble(4, 2, Fcm.ctypes.data_as(ip), Ucm.ctypes.data_as(dp), ptr.ctypes.data_as(ip),
                               faces.ctypes.data_as(ip), dblA.ctypes.data_as(dp), mass.ctypes.data_as(dp))
    assert dblA[0] == 0.0 and dblA[1] == 1.0
    assert np.array_equal(mass, igl_barycentric_mass(np.asarray(U), F)[1])
