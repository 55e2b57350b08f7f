"""The C++ LucidRenderer facade (include/lucid_renderer.hpp) driven from a small C++ program, the
way the reference's application drives its renderer."""
import os
import subprocess

import pytest

from lucid_b200 import build

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BIN = os.path.join(HERE, "cpp", "test_renderer")


def _compile():
    build.build()
    src = os.path.join(HERE, "cpp", "test_renderer.cpp")
    if os.path.exists(BIN) and os.path.getmtime(BIN) > max(os.path.getmtime(src), os.path.getmtime(build.SO_PATH)):
        return
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    lib_dir = os.path.dirname(build.SO_PATH)
    subprocess.run([cxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", BIN,
                    build.SO_PATH, f"-Wl,-rpath,{lib_dir}"], check=True)


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_facade_fails_loudly_without_a_device():
    """No CPU fallback: construction reports the CUDA error (skipped where a GPU is present)."""
    if _has_gpu():
        pytest.skip("a CUDA device is present")
    _compile()
    res = subprocess.run([BIN, "--no-device"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "failed as expected" in res.stdout


@pytest.mark.gpu
def test_facade_planes_known_answer():
    _compile()
    res = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "OK" in res.stdout
