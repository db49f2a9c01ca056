"""The Eigen stand-in (oracle/ref_shim) checked on its own against scipy / numpy: it is what
the reference's sources run on in oracle/_ref, so its arithmetic and - just as important - its
PATTERN semantics (structural zeros kept by products and transposes, duplicates summed by
setFromTriplets, coeffRef inserting) have to be Eigen's, not scipy's defaults."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.join(os.path.dirname(HERE), "oracle")
ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def shim():
    so = os.path.join(ORACLE, "_ref", "libref_shim_selftest.so")
    srcs = [os.path.join(ORACLE, "ref_shim_selftest.cpp"), os.path.join(ORACLE, "ref_shim", "Eigen", "Sparse"),
            os.path.join(ORACLE, "ref_shim", "Eigen", "Core")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE, "shimtest"], stdout=subprocess.DEVNULL)
    return C.CDLL(so)


def _csc(m):
    m = m.tocsc()
    return (np.ascontiguousarray(m.indptr, dtype=np.int32), np.ascontiguousarray(m.indices, dtype=np.int32),
            np.ascontiguousarray(m.data, dtype=np.float64))


def _args(m):
    p, i, v = _csc(m)
    return (p.ctypes.data_as(ip), i.ctypes.data_as(ip), v.ctypes.data_as(dp)), (p, i, v)


def _result(shim):
    r, c = C.c_int(), C.c_int()
    shim.shim_result_dims(C.byref(r), C.byref(c))
    nnz = shim.shim_result_nnz()
    p = np.empty(c.value + 1, dtype=np.int32)
    i = np.empty(max(nnz, 1), dtype=np.int32)
    v = np.empty(max(nnz, 1), dtype=np.float64)
    shim.shim_result_copy(p.ctypes.data_as(ip), i.ctypes.data_as(ip), v.ctypes.data_as(dp))
    m = sp.csc_matrix((r.value, c.value))
    m.indptr, m.indices, m.data = p, i[:nnz], v[:nnz]
    return m


def _random_sparse(rng, rows, cols, density, zero_fraction=0.2):
    m = sp.random(rows, cols, density=density, random_state=np.random.RandomState(rng.integers(1 << 30)), format="csc")
    m.sort_indices()
    z = rng.random(m.nnz) < zero_fraction  # explicit zeros, like the padded prolongations
    m.data[z] = 0.0
    return m


def _structural_product_pattern(A, B):
    a, b = A.copy(), B.copy()
    a.data = np.ones_like(a.data)
    b.data = np.ones_like(b.data)
    c = (a @ b).tocsc()
    c.sort_indices()
    return c


def test_sparse_products_keep_structural_zeros(shim):
    rng = np.random.default_rng(0)
    A, B, D = _random_sparse(rng, 40, 30, 0.15), _random_sparse(rng, 30, 50, 0.12), _random_sparse(rng, 50, 20, 0.2)
    (aa, ka), (bb, kb), (dd, kd) = _args(A), _args(B), _args(D)
    shim.shim_spgemm(40, 30, *aa, 50, *bb)
    got = _result(shim)
    want = _structural_product_pattern(A, B)
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    assert got.nnz > (A @ B).tocsc().count_nonzero()  # some stored entries are exactly zero
    assert np.allclose(got.toarray(), (A @ B).toarray(), rtol=1e-14, atol=1e-15)
    shim.shim_triple(40, 30, *aa, 50, *bb, 20, *dd)
    got3 = _result(shim)
    want3 = _structural_product_pattern(want, D)
    assert np.array_equal(got3.indptr, want3.indptr) and np.array_equal(got3.indices, want3.indices)
    assert np.allclose(got3.toarray(), (A @ B @ D).toarray(), rtol=1e-13, atol=1e-15)


def test_product_accumulates_in_storage_order(shim):
    """entry (i, j) = sum over k ascending of A(i,k) B(k,j), first term assigned: reproduce it
    term by term (this order is what makes the Galerkin values bit-comparable)"""
    rng = np.random.default_rng(1)
    A, B = _random_sparse(rng, 25, 18, 0.3, 0.0), _random_sparse(rng, 18, 22, 0.3, 0.0)
    (aa, ka), (bb, kb) = _args(A), _args(B)
    shim.shim_spgemm(25, 18, *aa, 22, *bb)
    got = _result(shim)
    Ad, Bd = A.toarray(), B.toarray()
    for j in range(22):
        for p in range(got.indptr[j], got.indptr[j + 1]):
            i = got.indices[p]
            acc = None
            for k in B.indices[B.indptr[j]:B.indptr[j + 1]]:
                if i in A.indices[A.indptr[k]:A.indptr[k + 1]]:
                    term = Ad[i, k] * Bd[k, j]
                    acc = term if acc is None else acc + term
            assert got.data[p] == acc


def test_transpose_triplets_coeffref_diagonal(shim):
    rng = np.random.default_rng(2)
    A = _random_sparse(rng, 30, 45, 0.2)
    aa, keep = _args(A)
    shim.shim_transpose(30, 45, *aa)
    T = _result(shim)
    want = A.T.tocsc()
    want.sort_indices()
    assert T.shape == (45, 30) and np.array_equal(T.indptr, want.indptr) and np.array_equal(T.indices, want.indices)
    assert np.array_equal(T.data, want.data) and T.nnz == A.nnz  # explicit zeros survive
    # setFromTriplets: duplicates summed in order of appearance, explicit zeros kept, sorted
    r = np.array([3, 1, 3, 0, 3, 2, 1], dtype=np.int32)
    c = np.array([2, 0, 2, 1, 2, 2, 0], dtype=np.int32)
    v = np.array([1e16, 0.0, 1.0, 5.0, -1e16, 0.0, 0.0])
    shim.shim_from_triplets(4, 3, 7, r.ctypes.data_as(ip), c.ctypes.data_as(ip), v.ctypes.data_as(dp))
    M = _result(shim)
    assert M.nnz == 4  # (1,0) (0,1) (2,2) (3,2): zeros stored, duplicates merged
    assert M[3, 2] == (1e16 + 1.0) - 1e16 and M[0, 1] == 5.0
    assert np.array_equal(M.indices[M.indptr[2]:M.indptr[3]], [2, 3])
    # coeffRef(i, i) += d on existing and on missing entries; diagonal() reads 0 where absent
    B = sp.csc_matrix(np.array([[2.0, 1, 0], [1, 0, 0], [0, 0, 0.0]]))
    bb, keepb = _args(B)
    rr = np.array([0, 2], dtype=np.int32)
    dd = np.array([1e-12, 4.0])
    diag = np.empty(3)
    shim.shim_coeffref_add(3, 3, *bb, 2, rr.ctypes.data_as(ip), rr.ctypes.data_as(ip), dd.ctypes.data_as(dp),
                           diag.ctypes.data_as(dp))
    assert np.array_equal(diag, [2.0 + 1e-12, 0.0, 4.0])
    assert _result(shim).nnz == B.nnz + 1  # the missing (2,2) was inserted


def test_sparse_times_dense_and_ldlt(shim):
    rng = np.random.default_rng(3)
    A = _random_sparse(rng, 35, 28, 0.2)
    X = np.asfortranarray(rng.standard_normal((28, 3)))
    Y = np.empty((35, 3), order="F")
    aa, keep = _args(A)
    shim.shim_spmm(35, 28, *aa, X.ctypes.data_as(dp), 3, Y.ctypes.data_as(dp))
    assert np.allclose(Y, A @ X, rtol=1e-14, atol=1e-15)
    # column-major scatter order: y(i) accumulates its terms in ascending column order
    Ad = A.toarray()
    for i in range(35):
        acc = 0.0
        for j in range(28):
            if i in A.indices[A.indptr[j]:A.indptr[j + 1]]:
                acc += Ad[i, j] * (1.0 * X[j, 0])
        assert Y[i, 0] == acc
    # SPD system: the stand-in for SimplicialLDLT solves it to rounding
    n = 60
    S = sp.diags([-np.ones(n - 1), 2.5 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]).tocsc()
    S = (S + sp.csc_matrix(([-0.5, -0.5], ([0, n - 1], [n - 1, 0])), shape=(n, n))).tocsc()
    S.sort_indices()
    B = np.asfortranarray(rng.standard_normal((n, 2)))
    Xs = np.empty((n, 2), order="F")
    ss, keeps = _args(S)
    assert shim.shim_ldlt_solve(n, *ss, B.ctypes.data_as(dp), 2, Xs.ctypes.data_as(dp)) == 0
    assert np.allclose(S @ Xs, B, rtol=1e-12, atol=1e-12)
    # not positive definite -> reported
    Nn = sp.csc_matrix(np.array([[1.0, 2.0], [2.0, 1.0]]))
    nn, keepn = _args(Nn)
    assert shim.shim_ldlt_solve(2, *nn, B.ctypes.data_as(dp), 1, Xs.ctypes.data_as(dp)) == -1
