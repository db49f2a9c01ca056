"""pytest configuration: the ``gpu`` marker and shared problem fixtures."""
import os
import sys

import numpy as np
import pytest

# The multi-GPU tests run several ranks as threads that share one device and one CUDA
# context.  A kernel spinning in a halo exchange must never sit in front of another rank's
# kernels in the same hardware work queue, so give the context the maximum number of queues
# (one process per GPU, the production layout, has no such coupling).  Read at context
# creation, hence set before anything touches CUDA.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("SMG_XCHG_TIMEOUT_MS", "5000")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    """Build the oracle and (if missing) libsmg.so once per session."""
    from oracle import cpu_oracle

    cpu_oracle.build()
    from surface_multigrid_code_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()


def _warm_up_gpu():
    """First CUDA use on a fresh box pages in cuSOLVER/cuBLAS (can take a minute): do it once,
    outside any test's own time limits."""
    try:
        import torch

        if not torch.cuda.is_available():
            return
    except Exception:
        return
    from surface_multigrid_code_b200 import meshgen as mg
    from surface_multigrid_code_b200.solver import Solver

    pr = mg.sphere_problem(3, 2, pad_three=True)
    with Solver(device=0) as s:
        s.set_hierarchy(pr.P).precompute(pr.A, pr.known)
        s.solve(pr.rhs, pr.z0, pr.known_val, 1e-8, 5)


@pytest.fixture(scope="session", autouse=True)
def built():
    _ensure_built()
    _warm_up_gpu()


@pytest.fixture(scope="session")
def problems():
    """Small seeded problems covering the reference's callers (03, 04, 05)."""
    from surface_multigrid_code_b200 import meshgen as mg

    out = {}
    # 04-style: closed surface, a few pinned vertices, explicit-zero padded P (3 per row)
    out["sphere_pad"] = mg.sphere_problem(4, 3, pad_three=True)
    out["sphere"] = mg.sphere_problem(5, 4, pad_three=False, random_z0=True)
    # 03-style: open surface, longest boundary loop pinned (drops coarse columns)
    V, F = mg.grid_mesh(9, 7, jitter=0.2, seed=1)
    out["grid"] = mg.mesh_subdivided_problem("grid", V, F, 3, 3, pad_three=True)
    # 05-style: free variant, k = 3
    V0, F0 = mg.octahedron()
    Vs, Fs, P = mg.subdivision_hierarchy(V0, F0, 4, 3, project_sphere=True, pad_three=True)
    Vs = mg.normalize_unit_area(Vs, Fs)
    rng = np.random.default_rng(3)
    U = Vs * (1.0 + 0.05 * rng.standard_normal((Vs.shape[0], 1)))
    out["mcf"] = mg.mcf_step_problem(Vs, Fs, P, U=U)
    # 06-style: block (3n x 3n, interleaved xyz) hierarchy of mg_precompute_block, free variant
    Vb, Fb, Pb = mg.subdivision_hierarchy(V0, F0, 3, 3, project_sphere=True, pad_three=True)
    Vb = mg.normalize_unit_area(Vb, Fb)
    out["block"] = mg.balloon_step_problem(Vb * (1.0 + 0.1 * rng.standard_normal((Vb.shape[0], 1))), Fb, Pb)
    return out
