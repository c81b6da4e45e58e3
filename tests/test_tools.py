"""Development tools that carry logic of their own are checked on the CPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_microbench_lines_host_check(tmp_path):
    """tools/microbench_lines.cu (the line-blocked form of the dictionary sweep): its per-thread
    function is __host__ __device__; run on the CPU it must reproduce the one-row-per-thread result bit for bit for
    7- and 27-point stencils, R = 2, 4, 8, on grids whose line counts R does and does not divide."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "microbench_lines")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-fmad=false", "-ccbin", "/usr/bin/g++",
                           "-o", exe, os.path.join(ROOT, "tools", "microbench_lines.cu")])
    out = subprocess.run([exe, "--host-check"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:]
    assert out.stdout.count("bit-identical") == 24 and "MISMATCH" not in out.stdout


def test_microbench_transfer_host_check(tmp_path):
    """tools/microbench_transfer.cu (line-blocked restriction and prolongation of the geometric hierarchy): the
    per-thread functions on the CPU reproduce the dictionary walk of one row per thread bit for bit, R = 1, 2, 4."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "microbench_transfer")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-fmad=false", "-ccbin", "/usr/bin/g++",
                           "-o", exe, os.path.join(ROOT, "tools", "microbench_transfer.cu")])
    out = subprocess.run([exe, "--host-check"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:]
    assert out.stdout.count("restriction bit-identical, prolongation bit-identical") == 12 and "MISMATCH" not in out.stdout
