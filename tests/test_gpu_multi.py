"""Multi-GPU parity (north star (5)): point-sharded BASELINE configs[4] and a slab-sharded grid on
every GPU of the box, one process per GPU, outputs all-gathered through dind_allgather_out.  Skipped with fewer
than 2 devices.  DIND_MULTI_FULL=1 runs the named size (4*10^9 points; needs 8 GPUs for the 64 GB gathered output)."""
import os
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


def test_point_sharded_c5_and_slab_grid_with_nccl_allgather():
    import dind_b200
    n_dev = dind_b200._lib.load().dind_device_count()
    if n_dev < 2:
        pytest.skip("needs at least 2 CUDA devices")
    world = 8 if n_dev >= 8 else (4 if n_dev >= 4 else 2)
    full = os.environ.get("DIND_MULTI_FULL") == "1"
    n_total = 4_000_000_000 if full else 20_000_000 * world + 3      # + 3: uneven shards (broadcast path)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multi_gpu_worker.py"), "--n-total", str(n_total),
           "--report", os.path.join(ROOT, "gpurun_out", f"multi_gpu_report_n{world}{'_full' if full else ''}.json")]
    if full and world < 8:
        cmd.append("--no-gather")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"multi_gpu_worker_n{world}.log"), "w") as f:
        f.write(p.stdout + "\n---- stderr ----\n" + p.stderr)
    tail = (p.stdout + p.stderr)[-4000:]
    assert p.returncode == 0, tail
    assert "MULTI_GPU_REPORT" in p.stdout, tail
