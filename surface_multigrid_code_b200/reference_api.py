"""The reference's operator interface for the hot path, by name, on numpy/scipy types.

Mirrors ``src/mg_data.h``, ``src/min_quad_with_fixed_mg.h`` and ``src/mg_VCycle.h`` of
HTDerekLiu/surface_multigrid_code (same names, argument order/meaning, defaults and
return quirks) so tests and examples read like the reference's callers
(03_mg_solver/main.cpp:69-75).  Output arguments of the C++ API are returned instead.
Every call runs on the GPU through libsmg.so; there is no CPU implementation here.
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional

import numpy as np

from .solver import Solver


@dataclasses.dataclass
class mg_data:
    """src/mg_data.h:11-44.  ``mg_precompute`` (CPU, the caller's job) fills V, F,
    P_full; ``min_quad_with_fixed_mg_precompute`` owns A, A_diag, P, PT, which live on
    the device -- they are materialised here only on request (``mirror=True``)."""

    V: Optional[np.ndarray] = None
    F: Optional[np.ndarray] = None
    P_full: object = None
    A: object = None
    A_diag: Optional[np.ndarray] = None
    P: object = None
    PT: object = None

    def reset(self):
        self.V = self.F = self.P_full = self.A = self.A_diag = self.P = self.PT = None


@dataclasses.dataclass
class min_quad_with_fixed_mg_data:
    """src/min_quad_with_fixed_mg.h:22-29 (+ the device handle that replaces the
    SimplicialLDLT argument)."""

    n: int = 0
    known: Optional[np.ndarray] = None
    unknown: Optional[np.ndarray] = None
    LHS: object = None
    Auk: object = None
    solver: Optional[Solver] = None


def hierarchy_from_prolongations(P: List) -> List[mg_data]:
    """What ``mg_precompute`` leaves behind (src/mg_precompute.cpp:71-77) given the
    prolongations: mg[0] is the finest level, mg[lv].P_full maps level lv -> lv-1."""
    mg = [mg_data()]
    for p in P:
        mg.append(mg_data(P_full=p))
    return mg


def min_quad_with_fixed_mg_precompute(A, known, data: min_quad_with_fixed_mg_data, mg: List[mg_data],
                                      smoother: str = "multicolour", mirror: bool = False, **opts):
    """src/min_quad_with_fixed_mg.cpp:137-257 (``known`` given) / :3-51 (``known is None``)."""
    s = Solver(smoother=smoother, **opts)
    s.set_hierarchy([m.P_full for m in mg[1:]])
    s.precompute(A, known)
    data.n = A.shape[0]
    data.known = None if known is None else np.asarray(known, dtype=np.int32)
    data.unknown = s.unknown
    data.solver = s
    if mirror:
        data.LHS = s.matrix(0, "LHS")
        if known is not None:
            data.Auk = s.matrix(0, "Auk")
        for lv, m in enumerate(mg):
            m.A = s.matrix(lv, "A")
            m.A_diag = s.diag(lv)
            if lv >= 1:
                m.P, m.PT = s.matrix(lv, "P"), s.matrix(lv, "PT")
    return data


def min_quad_with_fixed_mg_solve(data: min_quad_with_fixed_mg_data, RHS, known_val, z0, mg=None,
                                 tolerance: float = 1e-3, maxIter: int = 20, r_his: Optional[list] = None):
    """src/min_quad_with_fixed_mg.cpp:288-361 / :80-135 -> (converged, z, r_his).
    Defaults tolerance 1e-3, maxIter 20 (cpp:63,77,270,285).  ``r_his`` is appended to,
    like the reference's push_back."""
    z, his, ok = data.solver.solve(RHS, z0, known_val, tolerance, maxIter)
    if r_his is None:
        r_his = []
    r_his.extend(float(v) for v in his)
    return ok, z, r_his


def _solver_of(mg_or_data) -> Solver:
    return mg_or_data.solver if isinstance(mg_or_data, min_quad_with_fixed_mg_data) else mg_or_data


def mg_VCycle(data, B, preRelaxIter, postRelaxIter, lv, u):
    """src/mg_VCycle.cpp:3-59 -> u"""
    return _solver_of(data).vcycle(lv, B, u, preRelaxIter, postRelaxIter)


def A(u, data, lv):
    """src/mg_VCycle.cpp:62-70 -> Au"""
    return _solver_of(data).apply_A(lv, u)


def restrict(x, data, lv):
    """src/mg_VCycle.cpp:72-81 -> Rx"""
    return _solver_of(data).restrict(lv, x)


def prolong(x, data, lv):
    """src/mg_VCycle.cpp:83-92 -> Px"""
    return _solver_of(data).prolong(lv, x)


def relax(B, lv, iters, u, data):
    """src/mg_VCycle.cpp:113-178 -> u"""
    return _solver_of(data).relax(lv, iters, B, u)


def coarseSolve(data, B, lv, u):
    """src/mg_VCycle.cpp:181-201 -> u"""
    return _solver_of(data).coarse_solve(B, u)
