"""The production layout of the row partition: ONE PROCESS PER RANK, staging buffers mapped
across processes with CUDA IPC, halo exchanges that wait ON THE DEVICE inside the V-cycle graph
(halo_exchange_kernel; no host rendezvous).  tests/test_gpu_dist.py runs its ranks as threads of
one process, which takes the host-synchronised path; here every rank is a real process
(tests/dist_worker.py under torch.distributed.run):

* two processes on ONE GPU (what a single-GPU test box can run: the GPU time-slices the two
  contexts, so a kernel that waits for its peer is preempted until the peer's stores arrive);
* one process per GPU when the box has several.

Every rank checks its solution against the CPU checker (rel 1e-7) and that all ranks return
the same bits."""
import os
import re
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(nproc, sub, halo, env_extra=None, timeout=420):
    env = dict(os.environ)
    # ranks that share a GPU with each other (and with this pytest process's own context) are
    # time-sliced: give a waiting exchange plenty of time before it declares its peer dead
    env.setdefault("SMG_XCHG_TIMEOUT_MS", "120000")
    env.update(env_extra or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_worker.py"), str(sub), str(halo), "1"]
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    # one report per rank (the ranks share stdout: two reports may land on one line)
    lines = re.findall(r"rank \d+/\d+ dev \d+: .*?'exact': \d+\}", p.stdout, flags=re.S)
    assert len(lines) == nproc, p.stdout[-2000:]
    for ln in lines:
        assert "ok=True" in ln and "identical_on_all_ranks=True" in ln, ln
        assert "device_loop=True" in ln, ln  # one graph launch per solve on partitioned handles too
    return lines


@pytest.mark.parametrize("halo", [0, 1])
def test_two_processes_share_one_gpu_and_wait_on_the_device(halo):
    lines = _run(2, 5, halo)
    # the exchanges really ran (and through the device-waiting kernel: one launch per exchange)
    assert all("'exchanges': 0" not in ln for ln in lines)


def test_one_process_per_gpu():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    _run(min(n, 8), 7, 0)
    _run(2, 7, 1)
