"""An independent numpy/scipy restatement of the reference hot path, used ONLY to pin
the C oracle (tests/test_oracle.py) and to generate tests/golden/*.npz.

It is written from the behavioural description of the reference
(src/mg_VCycle.cpp:3-201, src/min_quad_with_fixed_mg.cpp:137-361), with scipy doing
the sparse algebra (different summation orders than the oracle, so comparisons are
to rounding, not bit-exact) and SuperLU doing the coarse solve.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def gauss_seidel(A_csr, diag, b, u, iters):
    """mg_VCycle.cpp:146-177 -- lexicographic forward sweeps, in place, per column."""
    u = np.array(u, dtype=np.float64, copy=True)
    cols = [u] if u.ndim == 1 else [u[:, c] for c in range(u.shape[1])]
    bs = [b] if u.ndim == 1 else [b[:, c] for c in range(u.shape[1])]
    indptr, indices, data = A_csr.indptr, A_csr.indices, A_csr.data
    for _ in range(iters):
        for x, rhs in zip(cols, bs):
            for i in range(A_csr.shape[0]):
                s = 0.0
                for p in range(indptr[i], indptr[i + 1]):
                    j = indices[p]
                    if j != i:
                        s += data[p] * x[j]
                x[i] = (rhs[i] - s) / diag[i]
    return u


class Hierarchy:
    """min_quad_with_fixed_mg_precompute (fixed variant when known is not None)."""

    def __init__(self, A, P_full, known):
        A = sp.csc_matrix(A)
        n = A.shape[0]
        self.n = n
        self.known = None if known is None else np.asarray(known, dtype=np.int64)
        if known is None:
            self.unknown = np.arange(n)
            LHS = A
            self.Auk = None
            P = [sp.csc_matrix(p) for p in P_full]
        else:
            mask = np.ones(n, dtype=bool)
            mask[self.known] = False
            self.unknown = np.nonzero(mask)[0]
            LHS = A[self.unknown][:, self.unknown]
            self.Auk = A[self.unknown][:, self.known]
            P = [sp.csc_matrix(p) for p in P_full]
            P[0] = sp.csc_matrix(P[0][self.unknown])
            self.keep = []
            for l in range(len(P)):
                Pl = sp.csc_matrix(P[l])
                # keep a column if any stored value exceeds 1e-15 (cpp:190-204)
                colmax = np.full(Pl.shape[1], -np.inf)
                for c in range(Pl.shape[1]):
                    seg = Pl.data[Pl.indptr[c]:Pl.indptr[c + 1]]
                    if seg.size:
                        colmax[c] = seg.max()
                keep = np.nonzero(colmax > 1e-15)[0]
                if keep.size < Pl.shape[1]:
                    P[l] = Pl[:, keep]
                    self.keep.append(keep)
                    if l + 1 < len(P):
                        P[l + 1] = sp.csc_matrix(P_full[l + 1])[keep]
                else:
                    self.keep.append(None)
                    break
        self.P = [sp.csr_matrix(p) for p in P]
        self.A = [sp.csr_matrix(LHS)]
        for p in self.P:
            self.A.append(sp.csr_matrix(p.T @ self.A[-1] @ p))
        Ac = self.A[-1].tolil()
        Ac.setdiag(Ac.diagonal() + 1e-12)
        self.A[-1] = sp.csr_matrix(Ac)
        self.diag = [a.diagonal() for a in self.A]
        self.coarse = spla.splu(sp.csc_matrix(self.A[-1]))

    def vcycle(self, b, u, lv=0, pre=2, post=2):
        last = len(self.A) - 1
        if lv == last:
            return u + self.coarse.solve(b)
        u = gauss_seidel(self.A[lv], self.diag[lv], b, u, pre)
        r = b - self.A[lv] @ u
        rc = self.P[lv].T @ r
        uc = self.vcycle(rc, np.zeros_like(rc), lv + 1, pre, post)
        u = u + self.P[lv] @ uc
        return gauss_seidel(self.A[lv], self.diag[lv], b, u, post)

    def solve(self, RHS, z0, known_val=None, tol=1e-3, max_iter=20):
        """min_quad_with_fixed_mg_solve incl. its stale-residual return value."""
        zu = np.array(z0[self.unknown], dtype=np.float64)
        bu = np.array(RHS[self.unknown], dtype=np.float64)
        if self.known is not None and self.known.size:
            bu = bu - self.Auk @ known_val
        r_his, residual = [], 0.0
        for _ in range(max_iter):
            residual = float(np.linalg.norm(bu - self.A[0] @ zu))
            r_his.append(residual)
            if residual < tol:
                break
            zu = self.vcycle(bu, zu)
        z = np.array(z0, dtype=np.float64, copy=True)
        z[self.unknown] = zu
        if self.known is not None:
            for i, idx in enumerate(self.known):  # sequential writes: last one wins
                z[idx] = known_val[i]
        return z, np.asarray(r_his), not (residual > tol)
