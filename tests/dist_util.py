"""Ranks that share a device also share its context (hardware queues, copy engines,
allocator, module loader); the library detects that layout and synchronises their exchanges on
the host instead of spinning on the device (csrc/smg.cu::HostGroup).  Test bodies still call
``s.barrier()`` between precompute and the first collective operator to keep the ranks close.

Drive N ranks of the row-partitioned solver from one process: one thread per rank
(ctypes releases the GIL inside libsmg calls, so the ranks really run concurrently, which
the flag-based halo exchange needs).  Ranks may share a device (rank r -> devices[r])."""
import threading

from surface_multigrid_code_b200.solver import Solver


def run_ranks(world, fn, devices=None, smoother="multicolour", exact=False, dist_levels=-1,
              min_rows=0, timeout=300.0, comm_bytes=32 << 20, **solver_kw):
    """fn(rank, solver) -> result, called on `world` connected solvers; returns the list of
    results in rank order.  Raises the first exception of any rank (or on timeout)."""
    devices = devices or [0] * world
    blobs = [None] * world
    results = [None] * world
    errors = [None] * world
    barrier = threading.Barrier(world, timeout=timeout)

    def body(r):
        s = None
        try:
            s = Solver(smoother=smoother, device=devices[r], **solver_kw)
            s.dist_init(r, world, comm_bytes).dist_options(exact, dist_levels, min_rows)
            blobs[r] = s.dist_handle()
            barrier.wait()
            s.dist_connect(blobs)
            s.barrier = barrier.wait  # ranks sharing a device: call after precompute (see below)
            results[r] = fn(r, s)
            barrier.wait()  # nobody frees its comm buffer while a peer may still write to it
        except BaseException as e:  # noqa: BLE001
            errors[r] = e
            try:
                barrier.abort()
            except Exception:
                pass
        finally:
            if s is not None:
                s.close()

    threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout)
    if any(t.is_alive() for t in threads):
        raise TimeoutError("a rank did not finish")
    real = [e for e in errors if e is not None and not isinstance(e, threading.BrokenBarrierError)]
    if real:
        raise real[0]
    if any(errors):
        raise next(e for e in errors if e is not None)
    return results
