"""The C++ host adapter (reference-named classes over the C-ABI) must compile with plain g++ and link against libeslam_b200.so."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_adapter_compiles_links_and_runs(tmp_path):
    exe = str(tmp_path / "host_adapter_check")
    libdir = os.path.join(ROOT, "object-oriented-slam_b200")
    cuda_lib = "/usr/local/cuda/lib64"
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "host_adapter_check.cpp"), "-L" + libdir, "-leslam_b200",
                           "-L" + cuda_lib, "-Wl,-rpath," + libdir, "-Wl,-rpath," + cuda_lib])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "adapter ok" in out.stdout


def test_reference_signature_overloads_compile_against_mock_eigen_opencv(tmp_path):
    """The ESLAM_HAVE_EIGEN_OPENCV overloads (the reference's exact prototypes: EllipsoidExtractor.h:57, Optimizer.h:19-21,
    Initializer.h:46-47, Frame.h:18-20, PlaneExtractor.h) are instantiated against mock Eigen / OpenCV headers and mock reference types:
    the image has neither library, so this is the compile check that can run here."""
    exe = str(tmp_path / "host_adapter_eigen_check")
    libdir = os.path.join(ROOT, "object-oriented-slam_b200")
    cuda_lib = "/usr/local/cuda/lib64"
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "tests", "mock_eigen_opencv"), "-o", exe,
                           os.path.join(ROOT, "tests", "host_adapter_eigen_check.cpp"), "-L" + libdir, "-leslam_b200", "-L" + cuda_lib,
                           "-Wl,-rpath," + libdir, "-Wl,-rpath," + cuda_lib])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "eigen adapter ok" in out.stdout
