"""Loader of tests/golden/*.npz (see tests/golden/make_golden.py)."""
import os

import numpy as np
import scipy.sparse as sp

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["sphere_s3_l3", "grid_s2_l3", "mcf_s3_l3", "bunny_l3", "ogre_l4"]


def load(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n, nlev = int(d["n"]), int(d["nlev"])
    A = sp.csc_matrix((d["A_data"], d["A_indices"], d["A_indptr"]), shape=(n, n))
    P = []
    for l in range(nlev - 1):
        shp = tuple(int(x) for x in d[f"P{l}_shape"])
        m = sp.csc_matrix(shp, dtype=np.float64)
        # assign the arrays directly: explicit zeros are part of the fixture
        m.indptr, m.indices, m.data = d[f"P{l}_indptr"], d[f"P{l}_indices"], d[f"P{l}_data"]
        P.append(m)
    has_known = bool(int(d["has_known"]))
    return {
        "A": A, "P": P, "known": d["known"] if has_known else None,
        "known_val": d["known_val"] if has_known else None, "rhs": d["rhs"], "z0": d["z0"],
        "tol": float(d["tol"]), "max_iter": int(d["max_iter"]), "z": d["z"], "r_his": d["r_his"],
        "converged": bool(int(d["converged"])), "unknown": d["unknown"], "nlev": nlev,
        "diag": [d[f"Alev{l}_dense_diag"] for l in range(nlev)],
        "rowsum": [d[f"Alev{l}_rowsum"] for l in range(nlev)],
    }
