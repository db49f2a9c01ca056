"""GPU test of the mean-curvature-flow step with device-side assembly (smg_mcf_setup /
smg_mcf_step, include/smg.h); the arithmetic core is also covered on the CPU by
tests/test_mcf_core.py.

Checks, against the host path the reference takes (assemble M - delta L and M U on the host,
precompute, solve; 05_example_mean_curvature_flow/main.cpp:66-76): the assembled matrix values
bit-exact, the step result to 1e-12, several steps in a row, and the CPU checker's result.
"""

import numpy as np
import pytest
import scipy.sparse as sp

from oracle.cpu_oracle import Oracle
from surface_multigrid_code_b200 import meshgen as mg
from surface_multigrid_code_b200.solver import Solver
from test_mcf_core import igl_barycentric_mass

pytestmark = [pytest.mark.gpu]


def _mesh():
    V0, F0 = mg.octahedron()
    Vs, Fs, P = mg.subdivision_hierarchy(V0, F0, 5, 3, project_sphere=True, pad_three=True)
    Vs = mg.normalize_unit_area(Vs, Fs)
    rng = np.random.default_rng(12)
    U = np.asfortranarray(Vs * (1.0 + 0.05 * rng.standard_normal((Vs.shape[0], 1))))
    return Vs, np.ascontiguousarray(Fs, dtype=np.int32), P, U


def _host_system(U, F, L0, delta):
    _, m = igl_barycentric_mass(np.asarray(U), F)
    A = (sp.diags(m) - delta * L0).tocsc()
    A.sort_indices()
    return A, np.asfortranarray(m[:, None] * U)


def test_mcf_step_matches_host_assembly():
    V, F, P, U = _mesh()
    delta = 0.01
    L0 = mg.cotmatrix(V, F).tocsc()
    L0.sort_indices()
    A, rhs = _host_system(U, F, L0, delta)
    with Solver(smoother="wavefront", device=0) as host_path, Solver(smoother="wavefront", device=0) as dev_path:
        host_path.set_hierarchy(P).precompute(A, None)
        z_host, r_host, ok_host = host_path.solve(rhs, U, None, 5e-7, 20)
        # the device path only needs the PATTERN from precompute (values: any SPD matrix of that
        # pattern, here identity - L); every step assembles its own values on the device
        pattern_holder = (sp.identity(L0.shape[0], format="csc") - L0).tocsc()
        pattern_holder.sort_indices()
        assert np.array_equal(pattern_holder.indices, L0.indices)
        dev_path.set_hierarchy(P).precompute(pattern_holder, None)
        dev_path.mcf_setup(F, L0, delta)
        z_dev, r_dev, ok_dev = dev_path.mcf_step(U, 5e-7, 20)
        assert np.array_equal(dev_path.matrix(0, "A").data, A.data)
        assert ok_dev == ok_host and len(r_dev) == len(r_host)
        assert np.allclose(r_dev, r_host, rtol=1e-9, atol=1e-16)
        assert np.linalg.norm(z_dev - z_host) <= 1e-12 * np.linalg.norm(z_host)
        # the CPU checker on the host-assembled system
        ora = Oracle(P).precompute(A, None)
        z_ref, r_ref, ok_ref = ora.solve(rhs, U, None, 5e-7, 20)
        assert ok_ref == ok_dev and np.linalg.norm(z_dev - z_ref) <= 1e-9 * np.linalg.norm(z_ref)
        # three more steps, each against a fresh host assembly
        Ucur = z_dev
        for _ in range(3):
            Ucur = np.asfortranarray(mg.normalize_unit_area(Ucur, F))
            A, rhs = _host_system(Ucur, F, L0, delta)
            host_path.update_values(A.data)
            z_host, _, _ = host_path.solve(rhs, Ucur, None, 5e-7, 20)
            z_dev, _, ok = dev_path.mcf_step(Ucur, 5e-7, 20)
            assert ok and np.linalg.norm(z_dev - z_host) <= 1e-12 * np.linalg.norm(z_host)
            Ucur = z_dev
