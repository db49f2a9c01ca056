"""Size-independent properties of the hot-path operators (what can be
checked at BASELINE.json's full sizes without a second implementation): linearity, symmetry of
A, adjointness of restrict / prolong, consistency of the fused residual kernels, the
Gauss-Seidel fixed point, and that a solve returns something that satisfies the system.
`eng` is anything with the operator methods of surface_multigrid_code_b200.solver.Solver
(the CPU oracle gets a thin shim so the same checks pin the checker itself)."""
import numpy as np


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


def check_operator_properties(eng, nlev, rng, tol=1e-11):
    out = {}
    for lv in range(nlev):
        n = eng.level_rows(lv)
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        a, b = 0.75, -1.5
        Ax, Ay = eng.apply_A(lv, x), eng.apply_A(lv, y)
        # linearity and symmetry of A
        assert _rel(eng.apply_A(lv, a * x + b * y), a * Ax + b * Ay) < tol, lv
        assert abs(np.dot(Ax, y) - np.dot(x, Ay)) <= tol * np.linalg.norm(Ax) * np.linalg.norm(y), lv
        # the fused residual kernels agree with A
        rhs = rng.standard_normal(n)
        r = eng.residual(lv, rhs, x)
        assert _rel(r, rhs - Ax) < tol, lv
        assert abs(eng.residual_norm(lv, rhs, x) - np.linalg.norm(r)) <= tol * np.linalg.norm(r), lv
        # Gauss-Seidel leaves an exact solution alone: b = A x*  =>  relax(b, x*) == x*
        fixed = eng.relax(lv, 2, Ax, x.copy())
        assert _rel(fixed, x) < 1e-9, lv
        # ... and contracts the error of anything else in the A-norm
        e0 = y - x
        e1 = eng.relax(lv, 1, Ax, y.copy()) - x
        assert np.dot(e1, eng.apply_A(lv, e1)) < np.dot(e0, eng.apply_A(lv, e0)), lv
        if lv + 1 < nlev:
            nc = eng.level_rows(lv + 1)
            xc = rng.standard_normal(nc)
            Pxc, Ry = eng.prolong(lv, xc), eng.restrict(lv, y)
            # restrict is the transpose of prolong
            assert abs(np.dot(Pxc, y) - np.dot(xc, Ry)) <= tol * np.linalg.norm(Pxc) * np.linalg.norm(y), lv
            # rows of P sum to one: constants are prolonged exactly
            assert _rel(eng.prolong(lv, np.ones(nc)), np.ones(n)) < tol, lv
        out[lv] = n
    return out


def check_solve_properties(eng, pr, tol=1e-10, max_iter=40):
    z, r_his, ok = eng.solve(pr.rhs, pr.z0, pr.known_val, tol, max_iter)
    assert ok and r_his[-1] < tol and len(r_his) >= 2
    # monotone after the first cycle or two (the first cycle from z0 = 0 may
    # raise the residual), and geometric: at least 2x per cycle on these meshes
    tail = r_his[2:]
    assert np.all(tail[1:] < 0.5 * tail[:-1])
    # the returned vector satisfies the reference's own stopping test, recomputed on the host
    A = pr.A.tocsr()
    unknown = np.setdiff1d(np.arange(pr.n), pr.known) if pr.known is not None else np.arange(pr.n)
    res = (pr.rhs - A @ z)
    res = res[unknown] if res.ndim == 1 else res[unknown, :]
    assert np.linalg.norm(res) < 10 * tol
    if pr.known is not None:
        zk = z[pr.known] if z.ndim == 1 else z[pr.known, :]
        assert np.array_equal(zk, pr.known_val)
    return z, r_his
