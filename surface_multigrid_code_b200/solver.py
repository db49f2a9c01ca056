"""Host-side wrapper of one libsmg handle (``include/smg.h``).

``Solver`` owns an ``smg_handle``: the multigrid hierarchy resident in HBM plus the
precomputed Galerkin operators, and exposes the operators of the reference's
``mg_VCycle.h`` / ``min_quad_with_fixed_mg.h`` on numpy arrays (copied in and out
by the library) or on device pointers.  The reference-named free functions live in
``reference_api.py``.

Dense blocks follow Eigen: a 1-D array is a ``VectorXd``; a 2-D ``(n, k)`` array is a
``MatrixXd`` and is handed over column-major.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L


class SmgError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{L.STATUS.get(status, status)}: {message}")
        self.status = status


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ip(a):
    return a.ctypes.data_as(L._ip)


def _dp(a):
    return a.ctypes.data_as(L._dp)


def _colmajor(a):
    """(n,) or (n,k) -> (flat col-major float64 buffer, k, ndim)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        return np.ascontiguousarray(a), 1, 1
    if a.ndim != 2:
        raise ValueError("dense blocks must be 1-D or 2-D")
    return np.ascontiguousarray(a.T).reshape(-1), a.shape[1], 2


def _from_colmajor(buf, n, k, ndim):
    if ndim == 1:
        return buf
    return np.asfortranarray(buf.reshape(k, n).T)


class Solver:
    """One libsmg handle.  ``device=None`` -> current CUDA device; ``device='none'``
    -> plan-only handle (host index planning; every compute call raises)."""

    def __init__(self, smoother: str = "multicolour", device=None, use_graph: bool = True,
                 pre_relax: int = 2, post_relax: int = 2, verbose: bool = False,
                 locality_reorder: bool = True, sigma: int = 256, patch_rows: Optional[int] = None):
        self._lib = L.load()
        opt = L.smg_options()
        self._lib.smg_default_options(C.byref(opt))
        opt.smoother = {"wavefront": L.SMOOTHER_WAVEFRONT, "multicolour": L.SMOOTHER_MULTICOLOUR,
                        "multicolor": L.SMOOTHER_MULTICOLOUR}[smoother]
        if device is None:
            opt.device = L.SMG_DEVICE_CURRENT
        elif device == "none":
            opt.device = L.SMG_DEVICE_NONE
        else:
            opt.device = int(device)
        opt.use_graph = int(bool(use_graph))
        opt.pre_relax = int(pre_relax)
        opt.post_relax = int(post_relax)
        opt.verbose = int(bool(verbose))
        opt.locality_reorder = int(bool(locality_reorder))
        opt.sigma = int(sigma)
        if patch_rows is not None:  # 0 automatic, < 0 off
            opt.patch_rows = int(patch_rows)
        self.smoother = smoother
        self.plan_only = device == "none"
        self._h = L._vp()
        rc = self._lib.smg_create(C.byref(self._h), C.byref(opt))
        if rc != L.SMG_OK:
            self._h = None
            raise SmgError(rc, "smg_create failed (no CUDA device? there is no CPU fallback)")
        self.n = None
        self.nknown = 0

    # -- lifetime ---------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.smg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != L.SMG_OK:
            raise SmgError(rc, self._lib.smg_last_error(self._h).decode())

    # -- multi-GPU (include/smg.h "multi-GPU" block): one Solver per rank ----------------
    def dist_init(self, rank: int, world: int, comm_bytes: int = 0):
        self._check(self._lib.smg_dist_init(self._h, int(rank), int(world), int(comm_bytes)))
        self.rank, self.world = int(rank), int(world)
        return self

    def dist_handle(self) -> bytes:
        """This rank's export blob; all-gather the blobs and pass them to dist_connect."""
        buf = C.create_string_buffer(self._lib.smg_dist_handle_bytes())
        self._check(self._lib.smg_dist_get_handle(self._h, buf))
        return buf.raw

    def dist_connect(self, blobs: Sequence[bytes]):
        raw = b"".join(blobs)
        self._check(self._lib.smg_dist_connect(self._h, C.create_string_buffer(raw, len(raw))))
        return self

    def dist_connect_torch(self, group=None):
        """all-gather the blobs over torch.distributed (any backend) and connect."""
        import torch.distributed as dist

        blobs = [None] * dist.get_world_size(group)
        dist.all_gather_object(blobs, self.dist_handle(), group=group)
        return self.dist_connect(blobs)

    def dist_connect_files(self, directory: str, tag: str = "smg", timeout_ms: int = 60000):
        """all-gather the blobs through files in `directory` (no MPI / torch needed) and connect."""
        self._check(self._lib.smg_dist_connect_files(self._h, directory.encode(), tag.encode(),
                                                     int(timeout_ms)))
        return self

    def dist_options(self, exact=False, dist_levels: int = -1, dist_min_rows: int = 0):
        """exact: False/0 = halo exchange per sweep, True/1 = per colour, 2 = per relax call."""
        self._check(self._lib.smg_dist_set_options(self._h, int(exact), int(dist_levels),
                                                   int(dist_min_rows)))
        return self

    def dist_info(self) -> dict:
        out = (C.c_int64 * 8)()
        self._check(self._lib.smg_dist_info(self._h, out))
        keys = ("rank", "world", "dist_levels", "exchanges", "connected", "slot_doubles", "exact")
        return dict(zip(keys, [int(v) for v in out]))

    def dist_level_info(self, lv: int) -> dict:
        out = (C.c_int64 * 8)()
        self._check(self._lib.smg_dist_level_info(self._h, lv, out))
        keys = ("layout", "parts", "own_rows", "halo_u_recv", "halo_u_send", "halo_r_recv", "own_begin",
                "own_end")
        return dict(zip(keys, [int(v) for v in out]))

    def dist_part(self, lv: int) -> np.ndarray:
        out = np.empty(self.level_rows(lv), dtype=np.int32)
        self._check(self._lib.smg_dist_get_part(self._h, lv, _ip(out)))
        return out

    def dist_exchange(self, lv: int, which: str, src: int, dst: int) -> np.ndarray:
        w = {"halo_u": 0, "halo_r": 1, "halo_pu": 2, "gather": 3}[which]
        n = C.c_int(0)
        self._check(self._lib.smg_dist_get_exchange(self._h, lv, w, src, dst, None, C.byref(n)))
        out = np.empty(n.value, dtype=np.int32)
        if n.value:
            self._check(self._lib.smg_dist_get_exchange(self._h, lv, w, src, dst, _ip(out), C.byref(n)))
        return out

    # -- hierarchy / precompute ----------------------------------------------------
    def set_hierarchy(self, P: Sequence):
        """``P[l-1]`` = ``mg[l].P_full`` (scipy sparse, n_{l-1} x n_l), l = 1..nlev-1."""
        Ps = []
        for p in P:
            p = p.tocsc()
            if not p.has_sorted_indices:
                p = p.copy()
                p.sort_indices()
            Ps.append(p)
        nlev = len(Ps) + 1
        n_rows = _i32([Ps[0].shape[0]] + [p.shape[1] for p in Ps]) if Ps else _i32([0])
        keep = [(_i32(p.indptr), _i32(p.indices), _f64(p.data)) for p in Ps]
        cp = (L._ip * max(len(Ps), 1))(*[_ip(k[0]) for k in keep])
        ri = (L._ip * max(len(Ps), 1))(*[_ip(k[1]) for k in keep])
        vv = (L._dp * max(len(Ps), 1))(*[_dp(k[2]) for k in keep])
        self._check(self._lib.smg_set_hierarchy(self._h, nlev, _ip(n_rows), cp, ri, vv))
        self.nlev = nlev
        return self

    def precompute(self, A, known: Optional[np.ndarray] = None):
        """min_quad_with_fixed_mg_precompute; ``known=None`` selects the variant
        without fixed values (src/min_quad_with_fixed_mg.cpp:3-51)."""
        A = A.tocsc()
        if not A.has_sorted_indices:
            A = A.copy()
            A.sort_indices()
        cp, ri, vv = _i32(A.indptr), _i32(A.indices), _f64(A.data)
        if known is None:
            rc = self._lib.smg_precompute(self._h, A.shape[0], _ip(cp), _ip(ri), _dp(vv), None, -1)
            self.nknown = 0
        else:
            kn = _i32(known)
            rc = self._lib.smg_precompute(self._h, A.shape[0], _ip(cp), _ip(ri), _dp(vv), _ip(kn), kn.size)
            self.nknown = int(kn.size)
        self._check(rc)
        self.n = A.shape[0]
        self.nnz = int(cp[-1])
        return self

    def update_values(self, A_data):
        vv = _f64(A_data)
        if vv.size != self.nnz:
            raise ValueError("value array does not match the precomputed pattern")
        self._check(self._lib.smg_update_values(self._h, _dp(vv)))
        return self

    # -- solve ---------------------------------------------------------------------------
    def solve(self, RHS, z0, known_val=None, tol: float = 1e-3, max_iter: int = 20):
        """min_quad_with_fixed_mg_solve on host arrays -> (z, r_his, converged)."""
        b, k, nd = _colmajor(RHS)
        x0, k2, _ = _colmajor(z0)
        if b.size != self.n * k or x0.size != b.size or k2 != k:
            raise ValueError("RHS / z0 shape mismatch")
        kv = None
        if self.nknown > 0:
            if known_val is None:
                raise ValueError("known_val is required")
            kv, k3, _ = _colmajor(known_val)
            if kv.size != self.nknown * k or k3 != k:
                raise ValueError("known_val shape mismatch")
        z = np.empty(self.n * k)
        r_his = np.zeros(max(int(max_iter), 1))
        nh, conv = C.c_int(0), C.c_int(0)
        rc = self._lib.smg_solve(self._h, _dp(b), _dp(kv) if kv is not None else None, _dp(x0), k,
                                 float(tol), int(max_iter), _dp(z), _dp(r_his), C.byref(nh), C.byref(conv))
        self._check(rc)
        return _from_colmajor(z, self.n, k, nd), r_his[: nh.value].copy(), bool(conv.value)

    def solve_host_ptr(self, RHS_ptr: int, known_val_ptr: Optional[int], z0_ptr: int, z_ptr: int,
                       k: int = 1, tol: float = 1e-3, max_iter: int = 20):
        """smg_solve on raw HOST addresses (e.g. pinned buffers): no conversion copies.
        Buffers are n x k col-major float64 (known_val: nknown x k)."""
        r_his = np.zeros(max(int(max_iter), 1))
        nh, conv = C.c_int(0), C.c_int(0)
        cast = lambda p: C.cast(C.c_void_p(p), L._dp) if p else None
        rc = self._lib.smg_solve(self._h, cast(RHS_ptr), cast(known_val_ptr), cast(z0_ptr), k,
                                 float(tol), int(max_iter), cast(z_ptr), _dp(r_his), C.byref(nh),
                                 C.byref(conv))
        self._check(rc)
        return r_his[: nh.value].copy(), bool(conv.value)

    def solve_device(self, d_RHS: int, d_known_val: Optional[int], d_z0: int, d_z: int, k: int = 1,
                     tol: float = 1e-3, max_iter: int = 20):
        """Same on raw device pointers (ints, e.g. ``tensor.data_ptr()``), col-major."""
        r_his = np.zeros(max(int(max_iter), 1))
        nh, conv = C.c_int(0), C.c_int(0)
        rc = self._lib.smg_solve_device(self._h, d_RHS, d_known_val, d_z0, k, float(tol), int(max_iter),
                                        d_z, _dp(r_his), C.byref(nh), C.byref(conv))
        self._check(rc)
        return r_his[: nh.value].copy(), bool(conv.value)

    # -- mean-curvature-flow step with device-side assembly -------------------------------------
    def mcf_setup(self, F, L, delta: float = 0.01):
        """F: (nF, 3) int faces; L: scipy sparse cotangent matrix with the pattern the handle was
        precomputed with (05_example_mean_curvature_flow/main.cpp:41, computed once)."""
        F = np.asarray(F)
        Fb = np.ascontiguousarray(F.T, dtype=np.int32).reshape(-1)  # column-major nF x 3
        L = L.tocsc()
        if not L.has_sorted_indices:
            L = L.copy()
            L.sort_indices()
        vv = _f64(L.data)
        if vv.size != self.nnz:
            raise ValueError("L does not have the precomputed pattern")
        self._check(self._lib.smg_mcf_setup(self._h, self.n, F.shape[0], _ip(Fb), _dp(vv), float(delta)))
        return self

    def mcf_step(self, U, tol: float = 5e-7, max_iter: int = 20):
        """One flow step (main.cpp:66-76) -> (U_new, r_his, converged); U is (n, 3)."""
        u, k, nd = _colmajor(U)
        if k != 3 or u.size != self.n * 3:
            raise ValueError("U must be n x 3")
        out = np.empty(self.n * 3)
        r_his = np.zeros(max(int(max_iter), 1))
        nh, conv = C.c_int(0), C.c_int(0)
        self._check(self._lib.smg_mcf_step(self._h, _dp(u), float(tol), int(max_iter), _dp(out), _dp(r_his),
                                           C.byref(nh), C.byref(conv)))
        return _from_colmajor(out, self.n, 3, 2), r_his[: nh.value].copy(), bool(conv.value)

    # -- mg_VCycle.h operators -------------------------------------------------------------
    def level_rows(self, lv: int) -> int:
        return int(self._lib.smg_level_rows(self._h, lv))

    def num_levels(self) -> int:
        return int(self._lib.smg_num_levels(self._h))

    def vcycle(self, lv, B, u, pre=2, post=2):
        b, k, nd = _colmajor(B)
        x, _, _ = _colmajor(u)
        x = x.copy()
        self._check(self._lib.smg_vcycle(self._h, lv, pre, post, _dp(b), _dp(x), k))
        return _from_colmajor(x, self.level_rows(lv), k, nd)

    def relax(self, lv, iters, B, u):
        b, k, nd = _colmajor(B)
        x, _, _ = _colmajor(u)
        x = x.copy()
        self._check(self._lib.smg_relax(self._h, lv, iters, _dp(b), _dp(x), k))
        return _from_colmajor(x, self.level_rows(lv), k, nd)

    def apply_A(self, lv, u):
        x, k, nd = _colmajor(u)
        y = np.empty(self.level_rows(lv) * k)
        self._check(self._lib.smg_apply_A(self._h, lv, _dp(x), _dp(y), k))
        return _from_colmajor(y, self.level_rows(lv), k, nd)

    def residual(self, lv, B, u):
        b, k, nd = _colmajor(B)
        x, _, _ = _colmajor(u)
        r = np.empty(self.level_rows(lv) * k)
        self._check(self._lib.smg_residual(self._h, lv, _dp(b), _dp(x), _dp(r), k))
        return _from_colmajor(r, self.level_rows(lv), k, nd)

    def residual_norm(self, lv, B, u) -> float:
        b, k, _ = _colmajor(B)
        x, _, _ = _colmajor(u)
        out = C.c_double(0.0)
        self._check(self._lib.smg_residual_norm(self._h, lv, _dp(b), _dp(x), k, C.byref(out)))
        return float(out.value)

    def restrict(self, lv, x):
        xb, k, nd = _colmajor(x)
        y = np.empty(self.level_rows(lv + 1) * k)
        self._check(self._lib.smg_restrict(self._h, lv, _dp(xb), _dp(y), k))
        return _from_colmajor(y, self.level_rows(lv + 1), k, nd)

    def prolong(self, lv, x):
        xb, k, nd = _colmajor(x)
        y = np.empty(self.level_rows(lv) * k)
        self._check(self._lib.smg_prolong(self._h, lv, _dp(xb), _dp(y), k))
        return _from_colmajor(y, self.level_rows(lv), k, nd)

    def coarse_solve(self, B, u):
        b, k, nd = _colmajor(B)
        x, _, _ = _colmajor(u)
        x = x.copy()
        self._check(self._lib.smg_coarse_solve(self._h, _dp(b), _dp(x), k))
        return _from_colmajor(x, self.level_rows(self.num_levels() - 1), k, nd)

    # -- index / topology outputs --------------------------------------------------------------
    @property
    def unknown(self) -> np.ndarray:
        nu = self._lib.smg_num_unknown(self._h)
        out = np.empty(max(nu, 0), dtype=np.int32)
        self._check(self._lib.smg_get_unknown(self._h, _ip(out)))
        return out

    def keep(self, lv) -> Optional[np.ndarray]:
        n = C.c_int(0)
        self._check(self._lib.smg_get_keep(self._h, lv, None, C.byref(n)))
        if n.value < 0:
            return None
        out = np.empty(n.value, dtype=np.int32)
        self._check(self._lib.smg_get_keep(self._h, lv, _ip(out), C.byref(n)))
        return out

    def matrix(self, lv, which="A", values: bool = True):
        import scipy.sparse as sp

        w = L.MAT[which]
        r, c, z = C.c_int(), C.c_int(), C.c_int()
        self._check(self._lib.smg_matrix_dims(self._h, lv, w, C.byref(r), C.byref(c), C.byref(z)))
        ip = np.empty(c.value + 1, dtype=np.int32)
        ix = np.empty(max(z.value, 1), dtype=np.int32)
        vv = np.zeros(max(z.value, 1), dtype=np.float64)
        self._check(self._lib.smg_matrix_copy(self._h, lv, w, _ip(ip), _ip(ix), _dp(vv) if values else None))
        m = sp.csc_matrix((r.value, c.value), dtype=np.float64)
        m.indptr, m.indices, m.data = ip, ix[: z.value], vv[: z.value]
        return m

    def diag(self, lv) -> np.ndarray:
        out = np.empty(self.level_rows(lv))
        self._check(self._lib.smg_get_diag(self._h, lv, _dp(out)))
        return out

    def phases(self, lv):
        n = C.c_int(0)
        out = np.empty(self.level_rows(lv), dtype=np.int32)
        self._check(self._lib.smg_get_phases(self._h, lv, C.byref(n), _ip(out)))
        return n.value, out

    def row_order(self, lv):
        """-> (perm, group_ptr): perm[r] = reference row at position r of the library's numbering,
        group_ptr = row offsets of the (part / phase) groups."""
        ng = C.c_int(0)
        self._check(self._lib.smg_get_row_order(self._h, lv, None, C.byref(ng), None))
        perm = np.empty(self.level_rows(lv), dtype=np.int32)
        gp = np.empty(ng.value + 1, dtype=np.int32)
        self._check(self._lib.smg_get_row_order(self._h, lv, _ip(perm), C.byref(ng), _ip(gp)))
        return perm, gp

    def padded_nnz(self, lv) -> int:
        v = C.c_int64(0)
        self._check(self._lib.smg_level_padded_nnz(self._h, lv, C.byref(v)))
        return int(v.value)

    def level_stats(self, lv) -> dict:
        out = (C.c_int64 * 8)()
        self._check(self._lib.smg_level_stats(self._h, lv, out))
        keys = ("rows", "nnz_ref", "nnz", "padded", "p_nnz", "p_padded", "pt_padded", "phases")
        return dict(zip(keys, [int(v) for v in out]))

    def patch_plan(self, lv: int, kind: str = "down", iters: int = 2, target_rows: int = 256,
                   smem_limit: int = 0, verify: bool = True) -> dict:
        """Lay out (and verify symbolically) the patch schedule of level lv; host only."""
        out = (C.c_int64 * 10)()
        self._check(self._lib.smg_patch_plan(self._h, lv, {"down": 0, "up": 1}[kind], iters, target_rows,
                                             smem_limit, int(verify), out))
        keys = ("patches", "owned", "local", "rhs_rows", "updates", "max_blob_bytes", "max_vec", "blob_bytes",
                "max_passes", "entries")
        return dict(zip(keys, [int(v) for v in out]))

    @property
    def solved_on_device(self) -> bool:
        """The last solve ran its loop on the device (one graph launch)."""
        return bool(self._lib.smg_solve_on_device(self._h))

    def level_patched(self, lv: int) -> int:
        return int(self._lib.smg_level_patched(self._h, lv))

    # -- measurement --------------------------------------------------------------------------
    def time_kernel(self, which: str, lv: int = 0, k: int = 1, reps: int = 20, flush_l2: bool = False):
        """-> (mean ms per rep, kernel launches per rep); CUDA events on the handle's stream."""
        ms, nl = C.c_float(0.0), C.c_int(0)
        self._check(self._lib.smg_time_kernel(self._h, L.KERNEL[which], lv, k, reps, int(flush_l2),
                                              C.byref(ms), C.byref(nl)))
        return float(ms.value), int(nl.value)

    def trace_iteration(self, k: int = 1, max_events: int = 512):
        """Device timeline of one solve-loop iteration -> list of (label, start_us, end_us)."""
        names = C.create_string_buffer(max_events * 64)
        t0, t1 = np.zeros(max_events), np.zeros(max_events)
        n = C.c_int(0)
        self._check(self._lib.smg_trace_iteration(self._h, k, max_events, names, len(names), _dp(t0),
                                                  _dp(t1), C.byref(n)))
        labels = names.value.decode().split("\n")[: n.value]
        return [(labels[i], float(t0[i]), float(t1[i])) for i in range(n.value)]

    @property
    def launch_count(self) -> int:
        return int(self._lib.smg_launch_count(self._h))

    def timings(self):
        out = np.zeros(8)
        self._lib.smg_get_timings(self._h, _dp(out), 8)
        return {"h2d_ms": out[0], "solve_ms": out[1], "d2h_ms": out[2], "plan_ms": out[3],
                "precompute_device_ms": out[4]}

    @property
    def stream(self) -> int:
        return int(self._lib.smg_get_stream(self._h) or 0)
