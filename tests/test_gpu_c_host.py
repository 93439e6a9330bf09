"""GPU: the C ABI driven from a plain-C host program (no Python / torch in the process) — tests/c_host/abi_host.c —
checked against the plain-C oracle.  Demonstrates the boundary INTEGRATION.md §2 describes for compiled hosts."""
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_host_through_the_abi(tmp_path):
    from consolver_b200 import _lib
    _lib.load()                                                     # builds libconsolver.so if needed
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    cuda = os.environ.get("CUDA_HOME") or os.path.dirname(os.path.dirname(shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"))
    exe = str(tmp_path / "abi_host")
    libdir, odir = os.path.join(ROOT, "consolver_b200"), os.path.join(ROOT, "oracle", "_build")
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", os.path.join(ROOT, "tests", "c_host", "abi_host.c"),
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           "-L", libdir, "-lconsolver", "-L", odir, "-loracle", "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm",
           f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{odir}", f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c host ok" in r.stdout
