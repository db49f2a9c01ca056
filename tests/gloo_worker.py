"""One rank of the CPU (gloo) test of the multi-rank host logic: every rank plans the
partition on its own (plan-only handle, no CUDA), the ranks cross-check over
torch.distributed that their views agree, and the bench's max-over-ranks aggregation runs."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from surface_multigrid_code_b200 import meshgen as mg
    from surface_multigrid_code_b200.solver import Solver

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    pr = mg.sphere_problem(5, 4, pad_three=True)
    s = Solver(device="none")
    s.dist_init(rank, world).dist_options(False, 2, 0)
    s.set_hierarchy(pr.P).precompute(pr.A, pr.known)
    view = {"rank": rank, "info": s.dist_info(), "levels": [s.dist_level_info(l) for l in range(pr.nlev)]}
    # what I send to q and what I expect from q, per exchange, as digests
    for which in ("halo_u", "halo_r", "gather"):
        for q in range(world):
            if q == rank:
                continue
            view[which, "send", q] = hashlib.sha1(s.dist_exchange(0, which, rank, q).tobytes()).hexdigest()
            view[which, "recv", q] = hashlib.sha1(s.dist_exchange(0, which, q, rank).tobytes()).hexdigest()
    views = [None] * world
    dist.all_gather_object(views, view)
    if rank == 0:
        n0 = s.level_rows(0)
        ranges = sorted((v["levels"][0]["own_begin"], v["levels"][0]["own_end"]) for v in views)
        assert ranges[0][0] == 0 and ranges[-1][1] == n0
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        assert sum(v["levels"][0]["own_rows"] for v in views) == n0
        for a in range(world):
            for b in range(world):
                if a == b:
                    continue
                for which in ("halo_u", "halo_r", "gather"):
                    assert views[a][which, "send", b] == views[b][which, "recv", a], (which, a, b)
                # a symmetric pattern gives symmetric halo sizes only in count of pairs, not rows;
                # what must hold: a sends to b iff b receives from a (checked above)
        assert all(v["info"]["dist_levels"] == 2 and v["info"]["world"] == world for v in views)
    # bench.py's aggregation: time = max over ranks, one partitioned job
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t[0]) == 10.0 + world - 1
    dist.barrier()
    if rank == 0:
        print("GLOO_DIST_OK", world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
