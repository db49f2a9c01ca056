"""ctypes front end of the CPU oracle (``oracle/smg_oracle.c``) and, with ``impl="ref"``, of
``oracle/_ref/libsmg_ref.so``: the reference's own two hot-path source files compiled
unmodified against the Eigen stand-in of ``oracle/ref_shim`` (``make -C oracle ref``).  Both
libraries export the same ``orc_*`` entry points.

TEST INFRASTRUCTURE ONLY.  Importable from ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs; the product package
``surface_multigrid_code_b200`` never imports it.  Pinning: the reference has no golden
vectors for this path and genuine Eigen is not in this image, so the C restatement is pinned
against (a) the reference sources on the stand-in Eigen (``tests/test_reference_sources.py``)
and (b) an independent scipy restatement; what remains unpinned is Eigen 3.3.7's own
arithmetic order, restated from its sources in ``oracle/ref_shim/Eigen/Sparse``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsmg_oracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libsmg_ref.so")
# the same orc_* entry points over adapter/smg_eigen_adapter.cpp + libsmg.so (GPU): the drop-in
# under test, reachable here only so that tests can drive it exactly like the reference code
_ADAPTER_PATH = os.path.join(_HERE, "_ref", "libsmg_adapter.so")
_adapter_lib = None
_lib = None
_ref_lib = None

_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "smg_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libsmg_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def ref_available(build_if_possible: bool = True) -> bool:
    """oracle/_ref/libsmg_ref.so exists (it is built where /root/reference is mounted and
    travels to the GPU box as a built file)."""
    if not os.path.exists(_REF_PATH) and build_if_possible and os.path.isdir("/root/reference/src"):
        subprocess.call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return os.path.exists(_REF_PATH)


def adapter_available() -> bool:
    return ref_available() and os.path.exists(_ADAPTER_PATH)


def _declare(L):
    if True:
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_prolongation.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp]
        L.orc_precompute.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _dp, _ip, C.c_int]
        L.orc_solve.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_int, C.c_double, C.c_int, _dp, _dp, _ip]
        L.orc_iterate.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_int, _dp]
        L.orc_vcycle.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int]
        L.orc_relax.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, C.c_int]
        for f in (L.orc_apply_A, L.orc_restrict, L.orc_prolong):
            f.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int]
        L.orc_coarse_solve.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
        L.orc_num_unknown.argtypes = [C.c_void_p]
        L.orc_get_unknown.argtypes = [C.c_void_p, _ip]
        L.orc_get_keep.argtypes = [C.c_void_p, C.c_int, _ip]
        L.orc_matrix_dims.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip, _ip, _ip]
        L.orc_matrix_copy.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip, _ip, _dp]
        L.orc_get_diag.argtypes = [C.c_void_p, C.c_int, _dp]
        L.orc_coarse_bandwidth.argtypes = [C.c_void_p]
        if hasattr(L, "orc_solve_twice_same_rhis"):
            L.orc_solve_twice_same_rhis.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_double, C.c_int, _ip, _ip]
        if hasattr(L, "orc_iterate_mt"):
            L.orc_iterate_mt.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_int, _dp, C.c_int]
    return L


def lib(impl: str = "port"):
    """impl: "port" = oracle/smg_oracle.c, "ref" = the reference sources on the Eigen stand-in."""
    global _lib, _ref_lib, _adapter_lib
    if impl == "adapter":
        if _adapter_lib is None:
            if not (ref_available() and os.path.exists(_ADAPTER_PATH)):
                raise OSError("oracle/_ref/libsmg_adapter.so is not built (make -C oracle ref)")
            _adapter_lib = _declare(C.CDLL(_ADAPTER_PATH))
        return _adapter_lib
    if impl == "ref":
        if _ref_lib is None:
            if not ref_available():
                raise OSError("oracle/_ref/libsmg_ref.so is not built (make -C oracle ref needs /root/reference)")
            _ref_lib = _declare(C.CDLL(_REF_PATH))
        return _ref_lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = _declare(C.CDLL(_LIB_PATH))
    return _lib


def _i(a):
    return a.ctypes.data_as(_ip)


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _colmajor(a):
    """n x k array -> flat col-major float64 buffer (the reference's Eigen layout)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        return np.ascontiguousarray(a), 1
    return np.ascontiguousarray(a.T).reshape(-1), a.shape[1]


def _from_colmajor(buf, n, k, ndim):
    if ndim == 1:
        return buf.copy()
    return buf.reshape(k, n).T.copy()


class Oracle:
    """CPU restatement of min_quad_with_fixed_mg_{precompute,solve} + mg_VCycle."""

    def __init__(self, P: List, impl: str = "port"):
        self.nlev = len(P) + 1
        self.impl = impl
        self._L = lib(impl)
        self._h = C.c_void_p(self._L.orc_create(self.nlev))
        for l, p in enumerate(P, start=1):
            p = p.tocsc()
            ip = np.ascontiguousarray(p.indptr, dtype=np.int32)
            ix = np.ascontiguousarray(p.indices, dtype=np.int32)
            vv = _f64(p.data)
            rc = self._L.orc_set_prolongation(self._h, l, p.shape[0], p.shape[1], _i(ip), _i(ix), _d(vv))
            assert rc == 0
        self.n = None

    def iterate_mt(self, bu, zu, cycles, threads=0):
        """NOT the reference algorithm: OpenMP multicolour Gauss-Seidel + row-parallel products,
        `cycles` x (residual norm + V(2,2)); -> (zu, r_his, threads used).  C port only."""
        b, k = _colmajor(bu)
        x, _ = _colmajor(zu)
        x = x.copy()
        r = np.zeros(cycles)
        used = self._L.orc_iterate_mt(self._h, _d(b), _d(x), k, cycles, _d(r), int(threads))
        return _from_colmajor(x, self.level_rows(0), k, np.ndim(zu)), r, int(used)

    def refresh_count(self) -> int:
        """precompute calls served by a numeric-only refresh (adapter build of the harness only)"""
        return int(self._L.orc_refresh_count()) if hasattr(self._L, "orc_refresh_count") else 0

    def __del__(self):
        try:
            if self._h:
                self._L.orc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def precompute(self, A, known: Optional[np.ndarray] = None):
        A = A.tocsc()
        ip = np.ascontiguousarray(A.indptr, dtype=np.int32)
        ix = np.ascontiguousarray(A.indices, dtype=np.int32)
        vv = _f64(A.data)
        self.n = A.shape[0]
        if known is None:
            rc = self._L.orc_precompute(self._h, self.n, _i(ip), _i(ix), _d(vv), None, -1)
            self.nknown = 0
        else:
            kn = np.ascontiguousarray(known, dtype=np.int32)
            rc = self._L.orc_precompute(self._h, self.n, _i(ip), _i(ix), _d(vv), _i(kn), kn.size)
            self.nknown = kn.size
        if rc != 0:
            raise RuntimeError(f"orc_precompute failed: {rc}")
        return self

    # ---- index / structure outputs -------------------------------------- #
    @property
    def unknown(self):
        nu = self._L.orc_num_unknown(self._h)
        out = np.empty(nu, dtype=np.int32)
        self._L.orc_get_unknown(self._h, _i(out))
        return out

    def keep(self, lv):
        n = self._L.orc_get_keep(self._h, lv, None)
        if n < 0:
            return None
        out = np.empty(n, dtype=np.int32)
        self._L.orc_get_keep(self._h, lv, _i(out))
        return out

    def matrix(self, lv, which="A"):
        import scipy.sparse as sp

        w = {"A": 0, "P": 1, "PT": 2, "LHS": 3, "Auk": 4}[which]
        r, c, z = C.c_int(), C.c_int(), C.c_int()
        rc = self._L.orc_matrix_dims(self._h, lv, w, C.byref(r), C.byref(c), C.byref(z))
        assert rc == 0
        ip = np.empty(c.value + 1, dtype=np.int32)
        ix = np.empty(max(z.value, 1), dtype=np.int32)
        vv = np.empty(max(z.value, 1), dtype=np.float64)
        self._L.orc_matrix_copy(self._h, lv, w, _i(ip), _i(ix), _d(vv))
        m = sp.csc_matrix((r.value, c.value), dtype=np.float64)
        m.indptr, m.indices, m.data = ip, ix[: z.value], vv[: z.value]
        return m

    def diag(self, lv):
        n = self.matrix(lv, "A").shape[0]
        out = np.empty(n)
        self._L.orc_get_diag(self._h, lv, _d(out))
        return out

    def level_rows(self, lv):
        r, c, z = C.c_int(), C.c_int(), C.c_int()
        self._L.orc_matrix_dims(self._h, lv, 0, C.byref(r), C.byref(c), C.byref(z))
        return r.value

    # ---- mg_VCycle.h pieces --------------------------------------------- #
    def relax(self, lv, iters, B, u):
        b, k = _colmajor(B)
        x, _ = _colmajor(u)
        self._L.orc_relax(self._h, lv, iters, _d(b), _d(x), k)
        return _from_colmajor(x, self.level_rows(lv), k, np.ndim(u))

    def _op(self, fn, lv, x, nout):
        xb, k = _colmajor(x)
        y = np.empty(nout * k)
        fn(self._h, lv, _d(xb), _d(y), k)
        return _from_colmajor(y, nout, k, np.ndim(x))

    def apply_A(self, lv, u):
        return self._op(self._L.orc_apply_A, lv, u, self.level_rows(lv))

    def restrict(self, lv, x):
        return self._op(self._L.orc_restrict, lv, x, self.level_rows(lv + 1))

    def prolong(self, lv, x):
        return self._op(self._L.orc_prolong, lv, x, self.level_rows(lv))

    def coarse_solve(self, B, u):
        b, k = _colmajor(B)
        x, _ = _colmajor(u)
        self._L.orc_coarse_solve(self._h, _d(b), _d(x), k)
        return _from_colmajor(x, self.level_rows(self.nlev - 1), k, np.ndim(u))

    def vcycle(self, lv, B, u, pre=2, post=2):
        b, k = _colmajor(B)
        x, _ = _colmajor(u)
        self._L.orc_vcycle(self._h, _d(b), pre, post, lv, _d(x), k)
        return _from_colmajor(x, self.level_rows(lv), k, np.ndim(u))

    # ---- min_quad_with_fixed_mg_solve ------------------------------------ #
    def solve(self, RHS, z0, known_val=None, tol=1e-3, max_iter=20):
        b, k = _colmajor(RHS)
        x0, _ = _colmajor(z0)
        z = np.empty(self.n * k)
        r_his = np.zeros(max(max_iter, 1))
        nh = C.c_int(0)
        kv = None
        if self.nknown > 0:
            if known_val is None:
                known_val = np.zeros((self.nknown,) if np.ndim(RHS) == 1 else (self.nknown, k))
            kv, _ = _colmajor(known_val)
        conv = self._L.orc_solve(self._h, _d(b), _d(kv) if kv is not None else None, _d(x0), k,
                                 float(tol), int(max_iter), _d(z), _d(r_his), C.byref(nh))
        return _from_colmajor(z, self.n, k, np.ndim(RHS)), r_his[: nh.value].copy(), bool(conv)

    def solve_twice_same_rhis(self, RHS, z0, known_val=None, tol=1e-3, max_iter=20):
        """Two solves that reuse ONE r_his vector (k = 1): its size after each (reference: cleared per solve)."""
        b, _ = _colmajor(RHS)
        x0, _ = _colmajor(z0)
        kv = None
        if self.nknown > 0:
            kv, _ = _colmajor(known_val if known_val is not None else np.zeros(self.nknown))
        n1, n2 = C.c_int(0), C.c_int(0)
        rc = self._L.orc_solve_twice_same_rhis(self._h, _d(b), _d(kv) if kv is not None else None, _d(x0),
                                               float(tol), int(max_iter), C.byref(n1), C.byref(n2))
        assert rc == 0
        return n1.value, n2.value

    def live_handles(self) -> int:
        return int(self._L.orc_live_handles()) if hasattr(self._L, "orc_live_handles") else 0

    def iterate(self, bu, zu, cycles):
        """`cycles` x (residual norm + V(2,2)) on the unknown-sized system (timed CPU baseline)."""
        b, k = _colmajor(bu)
        x, _ = _colmajor(zu)
        r = np.zeros(cycles)
        self._L.orc_iterate(self._h, _d(b), _d(x), k, cycles, _d(r))
        return _from_colmajor(x, self.level_rows(0), k, np.ndim(zu)), r
