"""Generates tests/golden/hilbert_cube_base.npz: the input of BASELINE config 5
(04_mg_solver_nobd/main.cpp on meshes/hilbert_cube.obj).

  V, F      the reference's mesh as read by igl::read_triangle_mesh (62 948 vertices);
  known     the 346 constrained vertices: nearest mesh vertex of every vertex of
            meshes/hilbert_cube_known.obj, 04_mg_solver_nobd/main.cpp:39-56 (first minimum wins);
  Pc0, Pc1  two coarsening prolongations below the original mesh from the stand-in
            meshgen.mis_hierarchy (NOT the reference's SSP decimation, which needs Eigen and
            stays on the CPU side by north_star): 62 948 -> ~15.7 K -> ~3.9 K.

bench.py --workload hilbert upsamples (V, F) three times (igl::upsample restated in
meshgen.upsample; original vertices keep their indices, so `known` stays valid) to 4 028 672
vertices: 6 levels = 3 subdivision levels + the original mesh + the 2 coarsened levels.
Only possible where /root/reference is mounted; the .npz travels to the GPU box.

    python tests/golden/make_hilbert.py        (run from the repo root)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from surface_multigrid_code_b200 import meshgen as mg  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    mesh_dir = "/root/reference/meshes"
    V, F = mg.read_obj(os.path.join(mesh_dir, "hilbert_cube.obj"))
    Vk, _ = mg.read_obj(os.path.join(mesh_dir, "hilbert_cube_known.obj"))
    known = np.empty(Vk.shape[0], dtype=np.int32)
    for i, v in enumerate(Vk):  # main.cpp:41-56: strict '<', the first minimum wins
        known[i] = int(np.argmin(np.linalg.norm(V - v, axis=1)))
    Vn = mg.normalize_unit_area(V, F)
    Pc = mg.mis_hierarchy(Vn, F, 3, pad_three=True)
    d = {"V": V, "F": F.astype(np.int32), "known": known}
    for l, p in enumerate(Pc):
        p = p.tocsc()
        d[f"Pc{l}_indptr"] = p.indptr.astype(np.int32)
        d[f"Pc{l}_indices"] = p.indices.astype(np.int32)
        d[f"Pc{l}_data"] = p.data
        d[f"Pc{l}_shape"] = np.asarray(p.shape, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "hilbert_cube_base.npz"), **d)
    print("V", V.shape, "F", F.shape, "known", known.size, "distinct", np.unique(known).size,
          "coarse levels", [p.shape for p in Pc])


if __name__ == "__main__":
    main()
