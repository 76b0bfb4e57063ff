"""The C ABI used from plain C (no Python objects, no torch): compiles tests/c_abi_smoke.c against
include/dind.h + libdind.so with gcc and runs it.  On the CPU box the library must refuse with
DIND_ERR_NO_DEVICE (exit 77); on the GPU box the client evaluates and checks the results itself."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "datainterpolationsnd.jl_b200")


def _build(tmp_path):
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.check_call(["/usr/bin/gcc", "-O1", "-std=c99", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", exe,
                           "-L", PKG, "-ldind", "-lm", f"-Wl,-rpath,{PKG}"])
    return exe


def test_c_client_refuses_without_device(tmp_path):
    import dind_b200
    if dind_b200._lib.load().dind_device_count() > 0:
        pytest.skip("a CUDA device is present")
    p = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert p.returncode == 77, p.stdout + p.stderr
    assert "no CPU fallback" in p.stdout


@pytest.mark.gpu
def test_c_client_on_gpu(tmp_path):
    p = subprocess.run([_build(tmp_path)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr


def _build_allgather(tmp_path):
    exe = str(tmp_path / "c_abi_allgather")
    subprocess.check_call(["/usr/bin/gcc", "-O1", "-std=gnu99", "-I", os.path.join(ROOT, "include"),
                           "-I", "/usr/local/cuda/include", os.path.join(ROOT, "tests", "c_abi_allgather.c"), "-o", exe,
                           "-L", PKG, "-ldind", "-L", "/usr/local/cuda/lib64", "-lcudart", "-lm",
                           f"-Wl,-rpath,{PKG}", "-Wl,-rpath,/usr/local/cuda/lib64"])
    return exe


def test_c_allgather_client_builds(tmp_path):
    """The multi-GPU client compiles and links against the header + libdind.so on the CPU box."""
    assert os.path.exists(_build_allgather(tmp_path))


@pytest.mark.gpu
def test_c_allgather_client_on_two_gpus(tmp_path):
    """dind_comm_init + dind_allgather_out from plain C, one process per GPU (NCCL inside libdind.so)."""
    import dind_b200
    if dind_b200._lib.load().dind_device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    exe = _build_allgather(tmp_path)
    idfile = str(tmp_path / "nccl_id")
    procs = [subprocess.Popen([exe, str(r), "2", idfile], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            p.kill()
            o, _ = p.communicate()
            o += "\nTIMEOUT"
        outs.append(o)
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
