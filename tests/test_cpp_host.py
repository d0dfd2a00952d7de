"""The C++ surface (include/vbdx.hpp over the C ABI): a host program restating the reference's integrator doctests is
compiled with g++ against libvbdx.so.  Without a GPU it must report the missing device (no fallback); on a GPU it must
pass."""
import os
import subprocess

import pytest

from physicsbasedanimationtoolkit_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "cube_doctest")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["/usr/bin/g++", "-std=c++20", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "cube_doctest.cpp"), "-L", libdir, "-l:libvbdx.so",
                    f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    return exe


def test_cpp_host_builds_and_refuses_to_run_without_a_device(tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; covered by the gpu test")
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3, r.stdout + r.stderr
    assert "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_cpp_host_cube_doctests(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PASS" in r.stdout
