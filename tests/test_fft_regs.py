"""Host-side check of the in-register butterflies (csrc/fft_regs.cuh are __host__ __device__): every radix against a
naive float64 DFT, the pruned 16-point transform against the full one, table twiddles against cmul / cmulc."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not available")
def test_fft_regs_host(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "fft_regs_host_test")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-x", "cu", "-I", os.path.join(ROOT, "fpm-opencv_b200", "csrc"),
                           "-o", exe, os.path.join(ROOT, "tests", "native", "fft_regs_host_test.cu")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not available")
def test_phased_kernel_shared_memory_layout(tmp_path):
    """PhasedShape<N>::layout (host code of csrc/fpm_update_phased.cuh): alignment, no overlaps, fits a B200 CTA."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "phased_layout_host_test")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-I",
                           os.path.join(ROOT, "fpm-opencv_b200", "csrc"), "-o", exe,
                           os.path.join(ROOT, "tests", "native", "phased_layout_host_test.cu")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not available")
def test_cluster_kernel_shared_memory_layout(tmp_path):
    """ClusterLayout<N, C> (host code of csrc/fpm_update_cluster.cuh): alignment, no overlaps, window-slice buffers for the
    narrow instances only, enough touched-cell slots for any rectangle position, fits a B200 CTA."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "cluster_layout_host_test")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-I",
                           os.path.join(ROOT, "fpm-opencv_b200", "csrc"), "-o", exe,
                           os.path.join(ROOT, "tests", "native", "cluster_layout_host_test.cu")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "61088 bytes" in out.stdout          # the figure the library reports for the bench geometry on four CTAs
