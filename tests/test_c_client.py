"""The C ABI from plain C (examples/c_client.c, C99, gcc — no C++, torch or Python on that side): it compiles against
include/ssd_b200.h, links libssd_b200.so, fails loudly without a CUDA device (there is no CPU fallback) and, on a GPU, steps
two handles — device-resident `ssd_step` and the pipelined host-buffer `ssd_step_host_async` — to identical observations,
rewards and dones."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build(tmp_path):
    lib_dir = os.path.join(ROOT, "contracts_b200")
    if not os.path.exists(os.path.join(lib_dir, "libssd_b200.so")):
        pytest.fail("contracts_b200/libssd_b200.so is missing: run __graft_entry__.build()")
    gcc = shutil.which("gcc")
    if gcc is None or not os.path.exists(os.path.join(CUDA, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc or the CUDA headers are not available")
    exe = str(tmp_path / "c_client")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           os.path.join(ROOT, "examples", "c_client.c"), "-o", exe, "-L", lib_dir, "-lssd_b200", "-L", os.path.join(CUDA, "lib64"),
           "-lcudart", "-Wl,-rpath,%s:%s" % (lib_dir, os.path.join(CUDA, "lib64"))]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    assert p.returncode == 0, p.stdout
    return exe


def _has_gpu():
    return os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl")


def test_c_client_compiles_as_c99_and_fails_loudly_without_gpu(tmp_path):
    exe = _build(tmp_path)
    if _has_gpu():
        pytest.skip("a GPU is present: covered by the gpu test")
    p = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert p.returncode == 3, (p.returncode, p.stdout, p.stderr)
    assert "ssd_create failed" in p.stderr and "no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_c_client_runs_on_gpu(tmp_path):
    exe = _build(tmp_path)
    p = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert p.returncode == 0, (p.returncode, p.stdout, p.stderr)
    assert p.stdout.startswith("c_client ok: 1024 envs x 4 agents x 40 steps"), p.stdout
