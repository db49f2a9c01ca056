"""Generates tests/golden/*.npz.

PARITY UNPINNED BY THE REFERENCE: the reference ships no golden vectors for this path
and cannot be compiled or imported here (C++/Eigen 3.3.7, Eigen absent).  These
fixtures therefore come from tests/scipy_restatement.py -- an independent numpy/scipy
restatement of mg_VCycle / min_quad_with_fixed_mg_{precompute,solve} -- plus one
self-contained known answer taken from the reference tree's own test-suite:
libigl/tests/include/igl/upsample.cpp:6-28 (upsample of a single triangle).

    python tests/golden/make_golden.py        (run from the repo root)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from scipy_restatement import Hierarchy  # noqa: E402
from surface_multigrid_code_b200 import meshgen as mg  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def pack(pr, z, r_his, ok, H):
    d = {
        "A_indptr": pr.A.indptr.astype(np.int32), "A_indices": pr.A.indices.astype(np.int32), "A_data": pr.A.data,
        "n": np.int64(pr.n), "nlev": np.int64(pr.nlev), "rhs": pr.rhs, "z0": pr.z0,
        "known": np.zeros(0, np.int32) if pr.known is None else pr.known.astype(np.int32),
        "has_known": np.int64(pr.known is not None),
        "known_val": np.zeros(0) if pr.known_val is None else pr.known_val,
        "tol": np.float64(pr.tol), "max_iter": np.int64(pr.max_iter),
        "z": z, "r_his": r_his, "converged": np.int64(ok), "unknown": H.unknown.astype(np.int32),
    }
    for l, p in enumerate(pr.P):
        p = p.tocsc()
        d[f"P{l}_indptr"] = p.indptr.astype(np.int32)
        d[f"P{l}_indices"] = p.indices.astype(np.int32)
        d[f"P{l}_data"] = p.data
        d[f"P{l}_shape"] = np.asarray(p.shape, dtype=np.int64)
    for l, a in enumerate(H.A):
        d[f"Alev{l}_dense_diag"] = a.diagonal()
        d[f"Alev{l}_rowsum"] = np.asarray(a.sum(axis=1)).ravel()
    return d


def main():
    # 04-style: closed sphere, pinned vertices, 3-per-row padded P, tol 1e-10
    pr = mg.sphere_problem(3, 3, tol=1e-10, pad_three=True)
    H = Hierarchy(pr.A, pr.P, pr.known)
    z, r, ok = H.solve(pr.rhs, pr.z0, pr.known_val, pr.tol, pr.max_iter)
    np.savez_compressed(os.path.join(HERE, "sphere_s3_l3.npz"), **pack(pr, z, r, ok, H))
    # 03-style: open jittered grid, boundary loop pinned with non-zero values, default tol
    V, F = mg.grid_mesh(7, 6, jitter=0.2, seed=2)
    pr = mg.mesh_subdivided_problem("grid", V, F, 2, 3, tol=1e-3, pad_three=True)
    pr.known_val = np.linspace(-1.0, 1.0, pr.known.size)
    H = Hierarchy(pr.A, pr.P, pr.known)
    z, r, ok = H.solve(pr.rhs, pr.z0, pr.known_val, pr.tol, pr.max_iter)
    np.savez_compressed(os.path.join(HERE, "grid_s2_l3.npz"), **pack(pr, z, r, ok, H))
    # 05-style: free variant, k = 3 (one mean-curvature-flow step)
    V0, F0 = mg.octahedron()
    Vs, Fs, P = mg.subdivision_hierarchy(V0, F0, 3, 3, project_sphere=True, pad_three=True)
    Vs = mg.normalize_unit_area(Vs, Fs)
    U = Vs * (1.0 + 0.05 * np.random.default_rng(3).standard_normal((Vs.shape[0], 1)))
    pr = mg.mcf_step_problem(Vs, Fs, P, U=U)
    H = Hierarchy(pr.A, pr.P, None)
    z, r, ok = H.solve(pr.rhs, pr.z0, None, pr.tol, pr.max_iter)
    np.savez_compressed(os.path.join(HERE, "mcf_s3_l3.npz"), **pack(pr, z, r, ok, H))
    # BASELINE configs 1-2: 03_mg_solver Poisson on the reference's own meshes (bunny.obj,
    # ogre.obj) with the stand-in MIS hierarchy (the reference's SSP hierarchy needs Eigen).
    # Only possible where /root/reference is mounted; the .npz travels to the GPU box.
    mesh_dir = "/root/reference/meshes"
    if os.path.isdir(mesh_dir):
        for name, nlev in (("bunny", 3), ("ogre", 4)):
            V, F = mg.read_obj(os.path.join(mesh_dir, name + ".obj"))
            pr = mg.mesh_problem(name, V, F, nlev, tol=1e-3, max_iter=20)
            H = Hierarchy(pr.A, pr.P, pr.known)
            z, r, ok = H.solve(pr.rhs, pr.z0, pr.known_val, pr.tol, pr.max_iter)
            np.savez_compressed(os.path.join(HERE, f"{name}_l{nlev}.npz"), **pack(pr, z, r, ok, H))
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
