"""The size-independent property checks (tests/property_checks.py) applied to the CPU checkers:
the C restatement and, when built, the reference's own sources."""
import numpy as np
import pytest

import property_checks as pc
from oracle import cpu_oracle
from oracle.cpu_oracle import Oracle


class OracleEngine:
    """the Solver method names on top of an Oracle (residual kernels composed from A)"""

    def __init__(self, ora):
        self.o = ora

    def __getattr__(self, name):
        return getattr(self.o, name)

    def relax(self, lv, iters, B, u):
        return self.o.relax(lv, iters, B, u.copy())  # the oracle relaxes in place

    def residual(self, lv, B, u):
        return B - self.o.apply_A(lv, u)

    def residual_norm(self, lv, B, u):
        return float(np.linalg.norm(B - self.o.apply_A(lv, u)))


IMPLS = ["port"] + (["ref"] if cpu_oracle.ref_available() else [])


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("name", ["sphere_pad", "grid"])
def test_operator_and_solve_properties(problems, name, impl):
    pr = problems[name]
    eng = OracleEngine(Oracle(pr.P, impl=impl).precompute(pr.A, pr.known))
    pc.check_operator_properties(eng, pr.nlev, np.random.default_rng(5))
    pc.check_solve_properties(eng, pr)
