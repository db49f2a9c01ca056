"""One rank of a multi-process run of the row-partitioned solver (one process per GPU, CUDA
IPC between them).  Launch: python -m torch.distributed.run --nproc-per-node N
--master-addr 127.0.0.1 --master-port P tests/dist_worker.py [subdivisions]
Checks on every rank: solve to 1e-10 matches the CPU oracle (rel 1e-7), every rank returns
the same bits.  Prints one line per rank."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from oracle.cpu_oracle import Oracle
    from surface_multigrid_code_b200 import meshgen as mg
    from surface_multigrid_code_b200.solver import Solver

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    ndev = torch.cuda.device_count()
    dev = local % ndev
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo")
    sub = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    pr = mg.sphere_problem(sub, min(4, sub - 1), pad_three=True, tol=1e-10)
    s = Solver(device=dev)
    s.dist_init(rank, world, 64 << 20)
    halo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    s.dist_options(halo, -1, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    s.dist_connect_torch()
    s.set_hierarchy(pr.P).precompute(pr.A, pr.known)
    dist.barrier()  # enter the first exchange together (a rank waits at most SMG_XCHG_TIMEOUT_MS for a peer)
    z, r_his, ok = s.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 40)
    ora = Oracle(pr.P).precompute(pr.A, pr.known)
    z_ref, r_ref, _ = ora.solve(pr.rhs, pr.z0, pr.known_val, 1e-10, 40)
    err = float(np.linalg.norm(z - z_ref) / np.linalg.norm(z_ref))
    zs = [None] * world
    dist.all_gather_object(zs, z.tobytes())
    same = all(b == zs[0] for b in zs)
    print(f"rank {rank}/{world} dev {dev}: ok={ok} cycles={len(r_his)} (oracle {len(r_ref)}) "
          f"rel_err={err:.2e} identical_on_all_ranks={same} device_loop={s.solved_on_device} info={s.dist_info()}",
          flush=True)
    assert ok and err < 1e-7 and same
    dist.barrier()
    s.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
